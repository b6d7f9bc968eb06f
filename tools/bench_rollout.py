"""Rollout-only timing sweep over (groups, steps_per_graph): CP tokens/s of the recurrent rollout
(256 songs, full-size actor) on one GPU.   python tools/bench_rollout.py [--songs 256] [--steps 256]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cpmusic

VOCAB = [56, 135, 18, 87, 18, 25]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--songs", type=int, default=256)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--configs", nargs="+", default=["1x1"], help="groups x steps_per_graph")
    ap.add_argument("--modes", nargs="+", default=["unfused", "tc", "tc+pdl"])
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    actor = cpmusic.LinearTransformer(VOCAB, dropout=0.1).to(dev)
    init = torch.stack([torch.randint(0, n, (args.songs,)) for n in VOCAB], -1).to(dev)
    for cfg in [(c, md) for c in args.configs for md in args.modes]:
        cfg, md = cfg
        g, spg = (int(x) for x in cfg.split("x"))
        kw = dict(mode=md.split("+")[0], pdl=md.endswith("+pdl"), lazy_state=md.endswith("+lazy"), split_state=md.endswith("+split"),
                  prefetch_state=int(md[-1]) if md[-4:-1] == "+pf" else 0)
        if g == 1:
            eng = cpmusic.RolloutEngine(actor, args.songs, args.steps, greedy=False, seed=1, **kw)
        else:
            eng = cpmusic.GroupedRolloutEngine(actor, args.songs, args.steps, groups=g, steps_per_graph=spg, greedy=False, seed=1, **kw)
        eng.generate(init)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.generate(init)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(json.dumps({"mode": md, "groups": g, "steps_per_graph": spg, "songs": args.songs, "us_per_token_step": round(ms * 1e3 / args.steps, 1),
                          "tokens_per_s": round(args.songs * args.steps / ms * 1e3)}), flush=True)
        del eng


if __name__ == "__main__":
    main()
