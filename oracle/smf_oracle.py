"""TEST INFRASTRUCTURE (checker, never shipped or measured): Python restatement of how mido 1.3.2 reads a Standard MIDI File --
``for msg in mido.MidiFile(path)`` as the reference calls it (MMGAN_MIDI_DES/datasets.py:18,34).  mido is a third-party dependency that is not in
the reference tree (requirements.txt: mido 1.3.2); the algorithm restated here is its published one: ``MidiFile._load`` / ``read_track``
(running status, meta and sysex events), ``merge_tracks`` (absolute ticks, stable sort, end_of_track metas dropped and one re-appended at
the last tick) and ``MidiFile.__iter__`` (``tick2second`` with the running tempo).  Pinned by tests/golden/midi_streams.npz (the reference's 30
shipped .mid files) and tests/golden/simlog_cases.npz (streams the unmodified reference produced through a mido build-side shim).
The product's reader is ``mmg_smf_parse`` (csrc/smf.cu); tests/test_smf_native.py holds it to this file bit for bit."""
import struct

import numpy as np

KIND_OTHER, KIND_ON, KIND_OFF = 0, 1, 2


def _vlq(buf, i):
    v = 0
    while True:
        b = buf[i]
        i += 1
        v = (v << 7) | (b & 0x7F)
        if not b & 0x80:
            return v, i


def parse_track(buf):
    """-> list of (abs_tick, kind, pitch, velocity, tempo_or_None, is_end_of_track)"""
    i, t, status, out = 0, 0, 0, []
    n = len(buf)
    while i < n:
        d, i = _vlq(buf, i)
        t += d
        b = buf[i]
        if b == 0xFF:                                   # meta
            typ = buf[i + 1]
            ln, j = _vlq(buf, i + 2)
            data = buf[j:j + ln]
            i = j + ln
            tempo = int.from_bytes(data, "big") if typ == 0x51 and ln == 3 else None
            out.append((t, KIND_OTHER, 0, 0, tempo, typ == 0x2F))
        elif b in (0xF0, 0xF7):                         # sysex
            ln, j = _vlq(buf, i + 1)
            i = j + ln
            out.append((t, KIND_OTHER, 0, 0, None, False))
        else:
            if b & 0x80:
                status = b
                i += 1
            hi = status & 0xF0
            nbytes = 1 if hi in (0xC0, 0xD0) else 2
            if status >= 0xF0:                           # system common / realtime
                nbytes = {0xF1: 1, 0xF2: 2, 0xF3: 1}.get(status, 0)
            d1 = buf[i] if nbytes >= 1 else 0
            d2 = buf[i + 1] if nbytes >= 2 else 0
            i += nbytes
            kind = KIND_ON if hi == 0x90 else KIND_OFF if hi == 0x80 else KIND_OTHER
            out.append((t, kind, d1 if kind else 0, d2 if kind else 0, None, False))
    return out


def read_smf_bytes(raw):
    """-> dt (float64 seconds), meta (uint32 kind | pitch << 8 | velocity << 16), abs ticks (int64), ticks_per_beat, [(tick, tempo)] of the set_tempo messages"""
    if raw[:4] != b"MThd":
        raise ValueError("not a Standard MIDI File")
    hlen, fmt, ntrk, div = struct.unpack(">IHHH", raw[4:14])
    if div & 0x8000:
        raise ValueError("SMPTE time division is not supported")
    if fmt == 2:
        raise TypeError("can't merge tracks in type 2 (asynchronous) file")
    pos, events = 8 + hlen, []
    for _ in range(ntrk):
        while raw[pos:pos + 4] != b"MTrk":
            pos += 8 + struct.unpack(">I", raw[pos + 4:pos + 8])[0]
        ln = struct.unpack(">I", raw[pos + 4:pos + 8])[0]
        events.extend(parse_track(raw[pos + 8:pos + 8 + ln]))
        pos += 8 + ln
    events.sort(key=lambda e: e[0])                      # stable, like mido.merge_tracks
    end_tick = max([e[0] for e in events], default=0)
    events = [e for e in events if not e[5]] + [(end_tick, KIND_OTHER, 0, 0, None, True)]
    ticks = np.array([e[0] for e in events], dtype=np.int64)
    dticks = np.diff(ticks, prepend=0)
    tempo, dt = 500000, np.zeros(len(events), dtype=np.float64)
    for i, e in enumerate(events):
        if dticks[i] > 0:
            dt[i] = int(dticks[i]) * (tempo * 1e-6 / div)
        if e[4] is not None:
            tempo = e[4]
    kind = np.array([e[1] for e in events], dtype=np.uint32)
    meta = kind | (np.array([e[2] for e in events], dtype=np.uint32) << 8) | (np.array([e[3] for e in events], dtype=np.uint32) << 16)
    return dt, meta, ticks, div, [(e[0], e[4]) for e in events if e[4] is not None]


def read_smf(path):
    with open(path, "rb") as f:
        return read_smf_bytes(f.read())


def beat_grid(changes, div, last_tick):
    """Quarter-note beats (seconds) along the tempo map from tick 0 to last_tick: the host-side stand-in for pretty_midi.get_beats
    (datasets.py:57; parity with pretty_midi itself is unpinned).  ``changes`` = [(0, 500000)] + the (tick, microseconds per beat) of every
    set_tempo message in stream order.  Checker of mmg_smf_beat_grid."""
    beats, t_sec, tick, k = [], 0.0, 0, 0
    tempo = changes[0][1]
    while tick <= last_tick:
        beats.append(t_sec)
        nxt = tick + div
        while k + 1 < len(changes) and changes[k + 1][0] < nxt:      # integrate across tempo changes
            k += 1
            c_tick = max(changes[k][0], tick)
            t_sec += (c_tick - tick) * (tempo * 1e-6 / div)
            tick, tempo = c_tick, changes[k][1]
        t_sec += (nxt - tick) * (tempo * 1e-6 / div)
        tick = nxt
    return np.array(beats, dtype=np.float64)


