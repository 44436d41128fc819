"""Host-side logic that needs no GPU: the SMF reader / EventStream front end and dataset plumbing."""
import os
import struct

import numpy as np
import pytest

from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds


def _vlq(n):
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    return bytes(reversed(out))


def _write_smf(path, tracks, tpb=480, fmt=1):
    body = b""
    for ev in tracks:
        t = b"".join(_vlq(d) + data for d, data in ev) + _vlq(0) + b"\xff\x2f\x00"
        body += b"MTrk" + struct.pack(">I", len(t)) + t
    with open(path, "wb") as f:
        f.write(b"MThd" + struct.pack(">IHHH", 6, fmt, len(tracks), tpb) + body)


def test_smf_reader_merges_tracks_and_follows_tempo(tmp_path):
    p = str(tmp_path / "a.mid")
    tempo = lambda us: b"\xff\x51\x03" + us.to_bytes(3, "big")
    _write_smf(p, [
        [(0, tempo(500000)), (960, tempo(250000))],
        [(0, b"\x90\x3c\x50"), (480, b"\x3c\x00"), (480, b"\x80\x40\x00"), (480, b"\x90\x41\x7f")],   # running status, vel-0 note_on
    ])
    s = ds.read_smf(p)
    kinds = s.meta & 0xFF
    # merged order by absolute tick (stable): tempo, on60 | on60(v0)@480 | tempo@960, off64@960 | on65@1440 | end_of_track
    assert list(kinds) == [0, 1, 1, 0, 2, 1, 0]
    assert list((s.meta >> 8) & 0xFF) == [0, 60, 60, 0, 64, 65, 0]
    assert list((s.meta >> 16) & 0xFF) == [0, 80, 0, 0, 0, 127, 0]
    # 480 ticks at 500000 us/beat = 0.5 s; after the set_tempo at tick 960, 480 ticks = 0.25 s
    assert np.allclose(s.dt, [0, 0, 0.5, 0.5, 0, 0.25, 0])
    assert s.dt[2] == 480 * (500000 * 1e-6 / 480)
    assert len(s.beats) >= 3 and s.beats[0] == 0.0 and abs(s.beats[1] - 0.5) < 1e-12


def test_event_stream_from_messages():
    class M:
        def __init__(self, type, time, note=0, velocity=0):
            self.type, self.time, self.note, self.velocity = type, time, note, velocity
    s = ds.EventStream.from_messages([M("note_on", 0.4, 60, 80), M("control_change", 1.0), M("note_off", 2.0, 60, 0)], beats=[0.5])
    assert list(s.meta) == [1 | 60 << 8 | 80 << 16, 0, 2 | 60 << 8]
    assert list(s.dt) == [0.4, 1.0, 2.0]
    with pytest.raises(ValueError):
        ds.EventStream([0.1, 0.2], [1])


def test_generate_piano_roll_rejects_bad_input():
    with pytest.raises(ValueError, match="midi_input must be a file path or a mido.MidiFile object"):
        ds.generate_piano_roll(123)


def test_pickle_dataset(tmp_path, monkeypatch):
    import pickle
    import torch
    monkeypatch.chdir(tmp_path)
    os.makedirs("data")
    items = [(torch.zeros(128, 50), torch.ones(128, 50), torch.arange(50.0)) for _ in range(3)]
    with open(os.path.join("data", "p.pkl"), "wb") as f:
        pickle.dump(items, f)
    d = ds.MaestroDatasetPickle("p.pkl")
    assert len(d) == 3 and d[1][2][7] == 7.0
    assert len(ds.MaestroDatasetMidi("nowhere")) == 0


def test_total_time_steps_matches_sequential_loop():
    """total_time of the notebook's preprocessing loop (cell 10): sequential float64 sum, round-half-even, the breaking message included."""
    import numpy as np
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
    rng = np.random.default_rng(2)
    for n, scale in ((0, 1.0), (7, 0.5), (400, 0.9), (400, 0.3), (50, 10.0)):
        dt = rng.exponential(scale, size=n)
        dt[rng.random(n) < 0.3] = 0.5
        s = ds.EventStream(dt, np.zeros(n, dtype=np.uint32))
        t, total = 0.0, 0
        for d in dt:
            t += float(d)
            total = int(round(t))
            if total >= 300:
                break
        assert ds.total_time_steps(s, 300) == total
