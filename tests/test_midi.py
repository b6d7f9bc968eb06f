"""CP words -> events -> Standard MIDI File and the generation driver, host side.
``inference_from_scratch`` is held to words the REFERENCE's own function produced (lifted from
dqn_policy/testing-no-type-cp.py by tests/golden/make_ref_golden.py).  ``write_midi`` cannot be run from the reference
(it needs miditoolkit, absent), so its event walk is checked against hand-derived expectations of
testing-no-type-cp.py:57-122 and by a write -> read round trip."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import ref_weights  # noqa: E402
from oracle import model_oracle as mo, sampling_oracle as so  # noqa: E402

VOCAB = [56, 135, 18, 87, 18, 25]
ATTRS = ("tempo", "chord", "barbeat", "pitch", "duration", "velocity")


class _ReferenceShapedOracle(torch.nn.Module):
    """The oracle behind the two methods the driver calls (the product model offers both natively)."""

    def __init__(self):
        super().__init__()
        self.m = mo.OracleCPModel(VOCAB, is_training=False, d_model=128, n_layer=2, n_head=2, d_inner=2048).eval()
        ref_weights.fill_(self.m, seed=11)

    def forward_hidden(self, x, memory=None, is_training=True):
        return self.m.forward_hidden(x, memory, is_training=is_training)

    def forward_output_sampling(self, h):
        return so.forward_output_sampling({a: lg.squeeze().numpy() for a, lg in zip(ATTRS, self.m.forward_output(h))})


def test_inference_from_scratch_matches_reference_run(cpm, golden):
    g = golden("ref_rl")
    _, w2e = ref_weights.synthetic_dictionary()
    np.random.seed(int(g["gen_np_seed"]))
    words = cpm.midi.inference_from_scratch(_ReferenceShapedOracle(), w2e, int(g["gen_bar_cond"]))
    assert words.dtype == np.int64 and np.array_equal(words, g["gen_words"])
    # the batched driver trims a longer roll-out at the same word
    longer = np.concatenate([g["gen_words"], g["gen_words"][1:]])
    assert np.array_equal(cpm.midi.trim_at_bar(longer, w2e, int(g["gen_bar_cond"])), g["gen_words"])
    assert np.array_equal(cpm.midi.trim_at_bar(g["gen_words"][:10], w2e, 50), g["gen_words"][:10])     # never reached


def test_decode_words_follows_the_reference_walk(cpm):
    e2w, w2e = ref_weights.synthetic_dictionary()
    W = lambda tempo=0, chord=0, bb=0, pitch=0, dur=0, vel=0: [e2w["tempo"][tempo], e2w["chord"][chord], e2w["bar-beat"][bb],  # noqa: E731
                                                               e2w["pitch"][pitch], e2w["duration"][dur], e2w["velocity"][vel]]
    words = np.array([
        W(bb="Bar"),                                                                    # bar 1
        W(tempo="Tempo_110", chord="C_M", bb="Beat_0"),                                 # tick 1920: tempo + marker
        W(pitch="Note_Pitch_60", dur="Note_Duration_480", vel="Note_Velocity_64"),
        W(pitch="Note_Pitch_64", dur="Note_Duration_0", vel="Note_Velocity_70"),        # duration 0 -> 60 ticks
        W(tempo="CONTI", chord="CONTI", bb="Beat_8"),                                   # tick 1920 + 8*120, nothing attached
        W(pitch="Note_Pitch_67", dur="Note_Duration_240", vel="Note_Velocity_100"),
        W(bb="Bar"),                                                                    # bar 2
        W(tempo="CONTI", chord="G_Mm7", bb="Beat_4"),
        W(pitch="Note_Pitch_55", dur="Note_Duration_960", vel="Note_Velocity_52"),
        W(),                                                                            # all-padding word: ignored
    ])
    song = cpm.midi.decode_words(words, w2e)
    N = cpm.midi.Note
    assert song.n_bars == 2
    assert song.tempos == [(1920, 110)]
    assert song.markers == [(1920, "C_M"), (2 * 1920 + 4 * 120, "G_Mm7")]
    assert song.notes == [N(60, 1920, 2400, 64), N(64, 1920, 1980, 70), N(67, 2880, 3120, 100), N(55, 4320, 5280, 52)]


def test_midi_file_round_trip(cpm, golden, tmp_path):
    _, w2e = ref_weights.synthetic_dictionary()
    words = golden("ref_rl")["gen_words"]
    path = str(tmp_path / "song.mid")
    song = cpm.midi.write_midi(words, path, w2e)                 # reference signature (words, path_outfile, word2event)
    # a random-weight model mixes note and metrical fields in one word; write_midi then reads the word as a NOTE and its
    # 'Bar' does not count (testing-no-type-cp.py:70-76), unlike the generation loop's bar counter (:171)
    metrical_bars = sum(1 for w in words if w[2] == 1 and not (w[3] and w[4] and w[5]))
    assert song.n_bars == metrical_bars and len(song.notes) == sum(1 for w in words if w[3] and w[4] and w[5])
    raw = open(path, "rb").read()
    assert raw[:4] == b"MThd" and raw[8:14] == bytes([0, 1, 0, 2, 0x01, 0xE0])      # format 1, 2 tracks, 480 ticks per beat
    tpb, back = cpm.midi.read_midi(path)
    assert tpb == 480
    # overlapping notes of one pitch cannot be paired uniquely in MIDI: compare the on and the off events as multisets
    assert sorted((n.pitch, n.start, min(max(n.velocity, 1), 127)) for n in song.notes) == sorted((n.pitch, n.start, n.velocity) for n in back.notes)
    assert sorted((n.pitch, n.end) for n in song.notes) == sorted((n.pitch, n.end) for n in back.notes)
    assert back.markers == sorted(song.markers, key=lambda m: m[0])
    in_time = sorted(song.tempos, key=lambda t: t[0])            # a random model's beat positions are not monotonic; the file is
    assert [t for t, _ in back.tempos] == [t for t, _ in in_time]
    assert all(abs(a[1] - b[1]) <= 1 for a, b in zip(back.tempos, in_time))         # bpm -> us per beat -> bpm rounding
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.mid"
        bad.write_bytes(b"RIFFxxxx")
        cpm.midi.read_midi(str(bad))


def test_load_dictionary_drops_type_and_orders_classes(cpm, tmp_path):
    e2w, w2e = ref_weights.synthetic_dictionary()
    full_e2w = {"tempo": e2w["tempo"], "chord": e2w["chord"], "bar-beat": e2w["bar-beat"], "type": {"EOS": 0, "Metrical": 1, "Note": 2},
                "pitch": e2w["pitch"], "duration": e2w["duration"], "velocity": e2w["velocity"]}
    full_w2e = {k: {w: e for e, w in v.items()} for k, v in full_e2w.items()}
    p = tmp_path / "dictionary.pkl"
    p.write_bytes(pickle.dumps((full_e2w, full_w2e)))
    a, b, n_class = cpm.midi.load_dictionary(str(p))
    assert list(a) == list(b) == ["tempo", "chord", "bar-beat", "pitch", "duration", "velocity"] and n_class == VOCAB
    assert cpm.midi.load_dictionary(str(p), drop_type=False)[2] == [56, 135, 18, 3, 87, 18, 25]
