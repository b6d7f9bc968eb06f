"""Static evidence per kernel, no GPU needed: registers / spills / shared memory from ``cuobjdump -res-usage`` and the
SASS mnemonics that prove what a kernel uses (``/opt/skills/guides/B200_PROFILING.md``): ``UTC*MMA`` (tcgen05.mma),
``UTCBAR`` (tcgen05.commit), ``LDTM`` / ``STTM`` (tcgen05.ld / st), ``UTMALDG`` / ``UTMASTG`` / ``UBLKCP`` / ``UBLKPF``
(TMA tensor / bulk copies and prefetch), ``SYNCS`` (mbarrier), ``HMMA`` (legacy mma.sync), ``MUFU``, 128- and 256-bit
global accesses, ``SHFL``, ``REDG`` / ``ATOMG``.

    python tools/sass_report.py > profiles/r01_sass_report.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "reinforcement-learning-in-music-generation_b200", "libcpmusic.so")
CUOBJDUMP = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")
CUFILT = os.environ.get("CUFILT", "/usr/local/cuda/bin/cu++filt")

PATTERNS = collections.OrderedDict([
    ("UTCMMA", r"\bUTC[A-Z]*MMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
    ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("UBLKPF", r"\bUBLKPF|\bUTMAPF"),
    ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("MUFU", r"\bMUFU"), ("LDG128", r"\bLDG\.E[.A-Z0-9]*\.128"),
    ("LDG256", r"\bLDG\.E[.A-Z0-9]*\.256"), ("STG128", r"\bSTG\.E[.A-Z0-9]*\.128"), ("STG256", r"\bSTG\.E[.A-Z0-9]*\.256"),
    ("SHFL", r"\bSHFL"), ("RED/ATOM", r"\bREDG|\bATOMG|\bRED\.|\bATOM\."), ("LDL/STL", r"\bLDL|\bSTL"),
])


def demangle(names):
    out = subprocess.run([CUFILT], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = {}
    for n, d in zip(names, out):
        d = re.sub(r"^void ", "", d)
        depth, cut = 0, len(d)                            # drop the trailing parameter list (template args may hold parentheses)
        for i in range(len(d) - 1, -1, -1):
            depth += d[i] == ")"
            depth -= d[i] == "("
            if depth == 0 and d[i] == "(":
                cut = i
                break
        d = d[:cut]
        short[n] = d.replace("cpm::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    return short


def main():
    if not os.path.exists(LIB):
        sys.exit(f"{LIB} missing: run python __graft_entry__.py first")
    res = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for line in res.splitlines():
        m = re.search(r"Function ([^:]+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            f = dict(re.findall(r"(\w+):(\d+)", line))
            usage[cur] = (int(f.get("REG", 0)), int(f.get("STACK", 0)), int(f.get("SHARED", 0)), int(f.get("LOCAL", 0)))
            cur = None
    sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True).stdout
    counts, n_instr, cur = {}, {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            n_instr[cur] = 0
            continue
        if cur is None or "/*" not in line:
            continue
        body = line.split("*/", 1)[-1] if line.lstrip().startswith("/*0") else line
        if re.search(r"^\s+/\*[0-9a-f]{4}\*/", line):
            n_instr[cur] += 1
            for key, pat in PATTERNS.items():
                if re.search(pat, body):
                    counts[cur][key] += 1
    names = sorted(counts)
    short = demangle(names)
    keys = list(PATTERNS)
    print("# libcpmusic.so, sm_100a: static per-kernel resources and SASS mnemonic counts (tools/sass_report.py)")
    print("# columns: registers/thread, stack bytes (spills), static shared bytes, SASS instructions, then mnemonic counts (blank = 0)")
    print("kernel | regs | stack | smem | instr | " + " | ".join(keys))
    for n in sorted(names, key=lambda k: short[k]):
        r = usage.get(n, (0, 0, 0, 0))
        cells = [str(counts[n][k]) if counts[n][k] else "" for k in keys]
        print(f"{short[n]} | {r[0]} | {r[1]} | {r[2]} | {n_instr[n]} | " + " | ".join(cells))
    tc = [short[n] for n in names if counts[n]["UTCMMA"]]
    tma = [short[n] for n in names if counts[n]["UTMALDG"] or counts[n]["UBLKCP"]]
    spill = [short[n] for n in names if usage.get(n, (0, 0))[1] > 0]
    print(f"\n# {len(names)} kernels; tcgen05.mma in {len(tc)}; TMA / bulk copies in {len(tma)}; kernels with a stack frame: {len(spill)}")
    print("# tcgen05 kernels: " + ", ".join(sorted(set(re.sub(r"<.*", "", t) for t in tc))))
    print("# kernels with a stack frame: " + ", ".join(sorted(set(re.sub(r"<.*", "", t) for t in spill))))


if __name__ == "__main__":
    main()
