"""CP words <-> events <-> Standard MIDI File, and the reference's generation driver.

* ``load_dictionary``          ``dictionary.pkl`` = ``(event2word, word2event)``; the generation and RL scripts drop the
                               ``type`` class before use (testing-no-type-cp.py:229-234, IRL_dqn_train.py:425-430).
* ``decode_words``             the event walk of ``write_midi`` (testing-no-type-cp.py:57-122): bar counter, beat position,
                               tempo changes, chord markers, notes — returned as plain tuples.
* ``write_midi``               same signature as the reference; writes a format-1 Standard MIDI File directly (the reference
                               goes through ``miditoolkit``, which is not a dependency here): conductor track with tempo and
                               marker meta events, one piano track.
* ``read_midi``                the inverse, for round-trip checks of files this module wrote.
* ``inference_from_scratch``   drop-in for testing-no-type-cp.py:126-179: prime with the bar token, sample until
                               ``bar_cond`` bars, return ``np (T, 6)``.  ``batched_generate`` is the device-resident form
                               (many songs at once through ``RolloutEngine``, trimmed at the bar condition on the host).
Host-side code only: nothing here touches the GPU except through the model it is handed.
"""
from __future__ import annotations

import pickle
import struct
from typing import Dict, List, NamedTuple, Sequence, Tuple

import numpy as np
import torch

BEAT_RESOL = 480                    # ticks per quarter note (testing-no-type-cp.py:52-54)
BAR_RESOL = BEAT_RESOL * 4
TICK_RESOL = BEAT_RESOL // 4
BAR_TOKEN = (0, 0, 1, 0, 0, 0)      # the priming word: bar-beat index 1 == 'Bar' (testing-no-type-cp.py:135-137)


class Note(NamedTuple):
    pitch: int
    start: int
    end: int
    velocity: int


class Song(NamedTuple):
    notes: List[Note]
    tempos: List[Tuple[int, int]]        # (tick, bpm)
    markers: List[Tuple[int, str]]       # (tick, chord text)
    n_bars: int


# ------------------------------------------------------------------------------------------------ dictionary
def load_dictionary(path: str, drop_type: bool = True):
    """-> (event2word, word2event, n_class).  ``n_class`` lists the class sizes in dictionary order, which is the model's
    ``n_token`` argument (testing-no-type-cp.py:239-242)."""
    with open(path, "rb") as f:
        event2word, word2event = pickle.load(f)
    if drop_type:
        event2word = {k: v for k, v in event2word.items() if k != "type"}
        word2event = {k: v for k, v in word2event.items() if k != "type"}
    return event2word, word2event, [len(event2word[k]) for k in event2word]


# ------------------------------------------------------------------------------------------------ words -> events
def decode_words(words, word2event: Dict[str, Dict[int, object]]) -> Song:
    """Walks ``words`` (T, n_class) exactly as the reference's ``write_midi`` does.  A word whose pitch, duration and
    velocity events are all strings is a note at the current position; anything else is metrical: 'Bar' advances the bar
    counter (so the first bar starts at tick ``BAR_RESOL``, one empty bar in, as in the reference), 'Beat_k' sets the
    position and may carry a tempo change and a chord marker ('CONTI' and the padding event 0 carry nothing).  Notes whose
    fields do not parse are skipped (the reference's bare ``except: continue``); duration 0 becomes 60 ticks."""
    keys = list(word2event.keys())
    notes, tempos, markers = [], [], []
    bar_cnt = cur_pos = 0
    for w in np.asarray(words):
        vals = [word2event[k][int(w[i])] for i, k in enumerate(keys)]
        if all(isinstance(v, str) for v in vals[3:6]):
            try:
                pitch, dur, vel = (int(v.split("_")[-1]) for v in vals[3:6])
            except ValueError:
                continue
            notes.append(Note(pitch, cur_pos, cur_pos + (dur if dur != 0 else 60), vel))
            continue
        pos = vals[2]
        if pos == "Bar":
            bar_cnt += 1
        elif isinstance(pos, str) and "Beat" in pos:
            cur_pos = bar_cnt * BAR_RESOL + int(pos.split("_")[1]) * TICK_RESOL
            if vals[1] != "CONTI" and vals[1] != 0:
                markers.append((cur_pos, str(vals[1])))
            if vals[0] != "CONTI" and vals[0] != 0:
                tempos.append((cur_pos, int(vals[0].split("_")[-1])))
    return Song(notes, tempos, markers, bar_cnt)


# ------------------------------------------------------------------------------------------------ Standard MIDI File
def _vlq(n: int) -> bytes:
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append(0x80 | (n & 0x7F))
        n >>= 7
    return bytes(reversed(out))


def _track(events: Sequence[Tuple[int, int, bytes]]) -> bytes:
    """events: (tick, order, payload) -> an MTrk chunk with delta times, closed by End Of Track."""
    body, last = bytearray(), 0
    for tick, _, payload in sorted(events, key=lambda e: (e[0], e[1])):
        body += _vlq(tick - last) + payload
        last = tick
    body += b"\x00\xff\x2f\x00"
    return b"MTrk" + struct.pack(">I", len(body)) + bytes(body)


def _meta(kind: int, data: bytes) -> bytes:
    return bytes([0xFF, kind]) + _vlq(len(data)) + data


def song_to_bytes(song: Song, ticks_per_beat: int = BEAT_RESOL) -> bytes:
    conductor = [(t, 0, _meta(0x51, struct.pack(">I", int(round(60_000_000 / max(bpm, 1))))[1:])) for t, bpm in song.tempos]
    conductor += [(t, 1, _meta(0x06, text.encode("utf-8"))) for t, text in song.markers]
    piano = [(0, 0, _meta(0x03, b"piano")), (0, 1, bytes([0xC0, 0]))]
    for n in song.notes:
        pitch, vel = min(max(n.pitch, 0), 127), min(max(n.velocity, 1), 127)
        piano.append((n.start, 3, bytes([0x90, pitch, vel])))
        piano.append((max(n.end, n.start), 2, bytes([0x80, pitch, 0])))     # offs sort before ons at the same tick
    header = b"MThd" + struct.pack(">IHHH", 6, 1, 2, ticks_per_beat)
    return header + _track(conductor) + _track(piano)


def write_midi(words, path_outfile: str, word2event) -> Song:
    """Reference signature (testing-no-type-cp.py:57).  Returns the decoded ``Song`` as well."""
    song = decode_words(words, word2event)
    with open(path_outfile, "wb") as f:
        f.write(song_to_bytes(song))
    return song


def read_midi(path: str):
    """-> (ticks_per_beat, Song).  Understands what ``song_to_bytes`` writes plus running status; n_bars is 0."""
    data = open(path, "rb").read()
    if data[:4] != b"MThd":
        raise ValueError("not a Standard MIDI File")
    _, _fmt, ntrk, tpb = struct.unpack(">IHHH", data[4:14])
    pos, notes, tempos, markers = 14, [], [], []
    for _ in range(ntrk):
        if data[pos:pos + 4] != b"MTrk":
            raise ValueError("bad track chunk")
        end = pos + 8 + struct.unpack(">I", data[pos + 4:pos + 8])[0]
        pos += 8
        tick, status, open_notes = 0, 0, {}
        while pos < end:
            delta = 0
            while True:
                b = data[pos]
                pos += 1
                delta = (delta << 7) | (b & 0x7F)
                if not b & 0x80:
                    break
            tick += delta
            if data[pos] & 0x80:
                status = data[pos]
                pos += 1
            if status == 0xFF:
                kind = data[pos]
                pos += 1
                ln = 0
                while True:
                    b = data[pos]
                    pos += 1
                    ln = (ln << 7) | (b & 0x7F)
                    if not b & 0x80:
                        break
                payload = data[pos:pos + ln]
                pos += ln
                if kind == 0x51:
                    tempos.append((tick, int(round(60_000_000 / int.from_bytes(payload, "big")))))
                elif kind == 0x06:
                    markers.append((tick, payload.decode("utf-8")))
            elif status & 0xF0 in (0x80, 0x90):
                pitch, vel = data[pos], data[pos + 1]
                pos += 2
                if status & 0xF0 == 0x90 and vel > 0:
                    open_notes.setdefault(pitch, []).append((tick, vel))
                elif open_notes.get(pitch):
                    start, v = open_notes[pitch].pop(0)
                    notes.append(Note(pitch, start, tick, v))
            elif status & 0xF0 in (0xC0, 0xD0):
                pos += 1
            else:
                pos += 2
        pos = end
    return tpb, Song(sorted(notes, key=lambda n: (n.start, n.pitch)), tempos, markers, 0)


# ------------------------------------------------------------------------------------------------ generation drivers
def inference_from_scratch(model, word2event, bar_cond: int, max_tokens: int = 20000, verbose: bool = False) -> np.ndarray:
    """testing-no-type-cp.py:126-179 on any model with the reference's recurrent surface
    (``forward_hidden(x (1,1,6), memory, is_training=False) -> (h, memory)``, ``forward_output_sampling(h) -> (6,) ints``):
    prime with the bar token, then sample / feed back until the bar counter (which starts at 1) reaches ``bar_cond``.
    ``max_tokens`` bounds a model that never emits 'Bar' (the reference loops forever there)."""
    bar_events = word2event["bar-beat"]
    device = next(model.parameters()).device
    words = [np.asarray(BAR_TOKEN, dtype=np.int64)]
    cnt_bar = 1
    with torch.no_grad():
        memory = None
        h, memory = model.forward_hidden(torch.as_tensor(words[0], device=device).view(1, 1, -1), memory, is_training=False)
        while len(words) < max_tokens:
            nxt = np.asarray(model.forward_output_sampling(h), dtype=np.int64)
            words.append(nxt)
            if verbose:
                print("bar:", cnt_bar, " ==", [word2event[k][int(nxt[i])] for i, k in enumerate(word2event)])
            h, memory = model.forward_hidden(torch.as_tensor(nxt, device=device).view(1, 1, -1), memory, is_training=False)
            if bar_events[int(nxt[2])] == "Bar":
                cnt_bar += 1
            if cnt_bar == bar_cond:
                break
    return np.stack(words)


def trim_at_bar(tokens: np.ndarray, word2event, bar_cond: int) -> np.ndarray:
    """The prefix of a generated sequence (T,6) that ``inference_from_scratch`` would have returned: up to and including
    the word that brings the bar counter (1 after the priming bar token at index 0) to ``bar_cond``."""
    bar_events = word2event["bar-beat"]
    cnt_bar = 1
    for t in range(1, len(tokens)):
        if bar_events[int(tokens[t][2])] == "Bar":
            cnt_bar += 1
        if cnt_bar == bar_cond:
            return tokens[:t + 1]
    return tokens


def batched_generate(model, word2event, bar_cond: int, n_songs: int, max_tokens: int = 4096, seed: int = 0, **kw) -> List[np.ndarray]:
    """Device-resident form: ``n_songs`` songs advance together through ``model.inference`` (``RolloutEngine``: one CUDA
    graph per token step, per-song Philox streams), one D2H copy at the end, each song trimmed at its own bar condition."""
    init = torch.tensor([BAR_TOKEN] * n_songs, dtype=torch.int64, device=next(model.parameters()).device)
    res = model.inference(init, max_tokens - 1, seed=seed, **kw)
    toks = res["tokens"].cpu().numpy()
    return [trim_at_bar(toks[i], word2event, bar_cond) for i in range(n_songs)]
