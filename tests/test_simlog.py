"""the sim-log -> note-event step (MMGAN_MIDI_DES/sim_log_to_midi.py mirror) against vectors frozen from the UNMODIFIED
reference (oracle/make_golden.py simlog: the reference's process_adjsim_log / MidiGenerator run on synthetic simulator logs through the
mido build-side shim).  CPU tests: the post-mido message stream (delta seconds bit-exact, kinds / pitches / velocities) and, through the
raster oracle, the piano rolls.  GPU test: the same logs through process_adjsim_log and through the batched device path."""
import os

import numpy as np
import pytest
import torch

import raster_oracle as ro


def _cases(golden_dir):
    c = np.load(os.path.join(golden_dir, "simlog_cases.npz"))
    for name in c["names"]:
        generate, start, end = (int(v) for v in c[name + ".args"])
        yield name, c, [str(x) for x in c[name + ".lines"]], bool(generate), start, end


def test_message_stream_matches_reference(golden_dir):
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import sim_log_to_midi as sl
    n_with_notes = 0
    for name, c, lines, generate, start, end in _cases(golden_dir):
        stream, gen = sl.sim_log_to_event_stream(lines, c[name + ".instruments"], c[name + ".note_levels"], c[name + ".gen2"], generate)
        assert len(stream) == len(c[name + ".dt"]), name
        assert np.array_equal(stream.dt, c[name + ".dt"]), name                    # float64 delta seconds, bit-exact
        assert np.array_equal(stream.meta, c[name + ".meta"]), name
        n_with_notes += int((stream.meta != 0).any())
        # the CPU restatement of the rasteriser on that stream gives the reference's rolls (W = end - start, sequence_length 100)
        kind, pitch, vel = ro.unpack_meta(stream.meta)
        roll, dur = ro.raster_events(stream.dt, kind, pitch, vel, 100, start, end)
        assert np.array_equal(roll, c[name + ".roll"]) and np.array_equal(dur, c[name + ".dur"]), name
    assert n_with_notes >= 8


def test_quirks_kept(golden_dir, tmp_path):
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import sim_log_to_midi as sl
    c = np.load(os.path.join(golden_dir, "simlog_cases.npz"))
    # a log whose line count is not a multiple of 100 never reaches the file: one end_of_track, empty roll (sim_log_to_midi.py:268-272)
    assert len(c["not_multiple_of_100.dt"]) == 1 and not c["not_multiple_of_100.roll"].any()
    with pytest.raises(TypeError):                 # `range` parameter shadows the builtin (:14,52)
        sl.MidiGenerator(n=10, instruments=None, note_levels=None, gen2_output=np.full(10, 0.5))
    with pytest.raises(ValueError, match="Error in processing log file"):
        sl.sim_log_to_event_stream(["INFO:root:1.0 - 4 - 99 - arrival\n"], np.arange(16), np.arange(16), np.full(10, 0.2))     # unknown server -> KeyError
    lp = sl.LogLineProcessor(sl.LOG_REGEX)
    assert lp.process_line("INFO:root:12.5 - 3 - 7 - departure") == ("12.5", "3", "7", "departure")
    assert lp.process_line("INFO:root:12.5 - 3 - 7 - processing") is None and lp.process_line("garbage") is None
    # the .mid written on request plays back (through the SMF reader of datasets.py) to the same stream, minus mido's trailing end_of_track bookkeeping
    name = "generate"
    p = str(tmp_path / "generation.mid")
    stream, _ = sl.sim_log_to_event_stream([str(x) for x in c[name + ".lines"]], c[name + ".instruments"], c[name + ".note_levels"], c[name + ".gen2"], True, p)
    back = ds.read_smf(p)
    assert np.array_equal(back.meta, stream.meta) and np.allclose(back.dt, stream.dt, rtol=0, atol=0)


def test_native_converter_matches_reference_vectors(golden_dir):
    """csrc/simlog.cu (mmg_simlog_to_events) on the vectors frozen from the unmodified reference: float64 delta seconds and records, bit-exact."""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import sim_log_to_midi as sl
    n = 0
    for name, c, lines, generate, _, _ in _cases(golden_dir):
        stream = sl.sim_log_to_event_stream_native(lines, c[name + ".instruments"], c[name + ".note_levels"], c[name + ".gen2"], generate)
        assert np.array_equal(stream.dt, c[name + ".dt"]) and np.array_equal(stream.meta, c[name + ".meta"]), name
        n += 1
    assert n == 11


def _fuzz_log(rng, n_lines, t_max):
    """synthetic logs that also visit the corners: integer and fractional times, times past 200, '.5'-style numbers, junk, huge queues"""
    sys_path_oracle()
    from make_golden import synth_sim_log
    lines = synth_sim_log(rng, n_lines, t_max, n_servers=int(rng.choice([3, 16])), n_customers=int(rng.choice([5, 40, 4000])), junk=0.15)
    for _ in range(int(rng.integers(0, 6))):
        i = int(rng.integers(0, len(lines)))
        lines[i] = str(rng.choice(["INFO:root:.5 - 4 - 2 - arrival\n", "INFO:root:7 - 4 - 2 - departure trailing\n", "INFO:root:12..5 - 4 - 2 - arrival\n",
                                   "INFO:root:3.25 - 8 - 1 - processing\n", " INFO:root:3 - 8 - 1 - arrival\n", "INFO:root:3 - 8 - 1 -arrival\n", "\n",
                                   "INFO:root:250.5 - 2 - 3 - arrival\n", "INFO:root:199.99 - 6 - 0 - departure"]))
    return lines


def sys_path_oracle():
    import sys
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    if d not in sys.path:
        sys.path.insert(0, d)


def test_native_converter_equals_python_mirror_on_random_logs():
    """300 random logs through the Python state machine (itself pinned to the reference by the vectors above and the live test) and through the
    native converter: identical streams; float32 and float64 gen2 rows (products in the row's own precision), generate on and off."""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import sim_log_to_midi as sl
    rng = np.random.default_rng(11)
    with_notes = 0
    for case in range(300):
        lines = _fuzz_log(rng, int(rng.choice([100, 200, 300, 137, 500, 1200])), float(rng.choice([20.0, 60.0, 150.0, 320.0])))
        g = np.concatenate([rng.random(6) * np.array([1, 1, 1.3, 1, float(rng.choice([1.0, 1.0, 12.0, 0.0])), float(rng.choice([1.0, 0.0]))]), rng.random(4)])
        g = g.astype(np.float32) if case % 2 == 0 else g
        ins, nl = rng.integers(0, 100, 16), rng.integers(0, 128, 16)
        generate = bool(case % 3 == 0)
        want, _ = sl.sim_log_to_event_stream(lines, ins, nl, g, generate)
        got = sl.sim_log_to_event_stream_native(lines, ins, nl, g, generate)
        assert np.array_equal(got.dt, want.dt) and np.array_equal(got.meta, want.meta), case
        with_notes += int((want.meta != 0).any())
    assert with_notes >= 100


def test_native_converter_refuses_what_the_reference_refuses():
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import sim_log_to_midi as sl
    g = np.full(10, 0.2, dtype=np.float32)
    for bad in (["INFO:root:1.0 - 4 - 99 - arrival\n"],                 # unknown server: KeyError
                ["INFO:root:1.0 - 4 - 03 - arrival\n"],                 # '03' is not a key either
                ["INFO:root:1.0 - 4.5 - 3 - arrival\n"],                # int('4.5')
                ):
        with pytest.raises(ValueError, match="Error in processing log file"):
            sl.sim_log_to_event_stream(bad, np.arange(16), np.arange(16), g)
        with pytest.raises(ValueError, match="Error in processing log file"):
            sl.sim_log_to_event_stream_native(bad, np.arange(16), np.arange(16), g)
    # the same lines are harmless where the reference never evaluates them: a departure of an unknown server, a customer the skips reject
    ok = ["INFO:root:1.0 - 4 - 99 - departure\n", "INFO:root:1.0 - 7 - 99 - arrival\n", "INFO:root:300 - 4.5 - 3 - arrival\n"]
    a, _ = sl.sim_log_to_event_stream(ok, np.arange(16), np.arange(16), g, True)
    b = sl.sim_log_to_event_stream_native(ok, np.arange(16), np.arange(16), g, True)
    assert np.array_equal(a.dt, b.dt) and np.array_equal(a.meta, b.meta)
    with pytest.raises(ValueError, match="Error in processing log file"):       # one bad song fails the batch, naming it
        sl.sim_logs_to_event_batch([ok, ["INFO:root:1.0 - 4 - 99 - arrival\n"]], [np.arange(16)] * 2, [np.arange(16)] * 2, [g, g], True)


def test_native_batch_equals_per_song_and_is_thread_count_independent():
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import sim_log_to_midi as sl
    rng = np.random.default_rng(3)
    S = 37
    logs = [_fuzz_log(rng, int(rng.choice([100, 200, 300])), 60.0) for _ in range(S)]
    ins, nl = rng.integers(0, 100, (S, 16)), rng.integers(0, 128, (S, 16))
    g = rng.random((S, 10)).astype(np.float32)
    per_song = [sl.sim_log_to_event_stream(lg, ins[i], nl[i], g[i], False)[0] for i, lg in enumerate(logs)]
    for threads in (1, 4, 0):
        dt, meta, off = sl.sim_logs_to_event_batch(logs, ins, nl, g, False, threads=threads)
        assert off.numel() == S + 1 and int(off[-1]) == sum(len(s) for s in per_song) == dt.numel() == meta.numel()
        for i, s in enumerate(per_song):
            a, b = int(off[i]), int(off[i + 1])
            assert np.array_equal(dt[a:b].numpy(), s.dt) and np.array_equal(meta[a:b].numpy().view(np.uint32), s.meta), (threads, i)
    dt, meta, off = sl.sim_logs_to_event_batch([], [], [], np.zeros((0, 10), dtype=np.float32))
    assert dt.numel() == 0 and off.tolist() == [0]


def test_process_adjsim_log_hands_the_reference_stream_to_the_rasteriser(golden_dir, monkeypatch, tmp_path):
    """host half of process_adjsim_log on the reference vectors (the device rasterisation is replaced by a recorder): the native conversion when
    nothing is written to disk, the Python state machine when a .mid is asked for -- the same stream either way; argument errors as in the reference"""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import sim_log_to_midi as sl
    seen = {}

    def recorder(stream, start=0, end=50, **kw):
        seen["stream"], seen["window"] = stream, (start, end)
        return "roll", "dur", "beats"

    monkeypatch.setattr(ds, "generate_piano_roll", recorder)
    calls = {"native": 0, "python": 0}
    real_native, real_py = sl.sim_log_to_event_stream_native, sl.sim_log_to_event_stream
    monkeypatch.setattr(sl, "sim_log_to_event_stream_native", lambda *a, **k: (calls.__setitem__("native", calls["native"] + 1), real_native(*a, **k))[1])
    monkeypatch.setattr(sl, "sim_log_to_event_stream", lambda *a, **k: (calls.__setitem__("python", calls["python"] + 1), real_py(*a, **k))[1])
    for name, c, lines, generate, start, end in _cases(golden_dir):
        for midi_path in (None, str(tmp_path / "x.mid")):
            out = sl.process_adjsim_log(instruments=c[name + ".instruments"], note_levels=c[name + ".note_levels"], gen2_output=c[name + ".gen2"],
                                        start=start, end=end, generate=generate, log_lines=lines, midi_path=midi_path)
            assert out == ("roll", "dur", "beats") and seen["window"] == (start, end)
            assert np.array_equal(seen["stream"].dt, c[name + ".dt"]) and np.array_equal(seen["stream"].meta, c[name + ".meta"]), (name, midi_path)
    assert calls == {"native": 11, "python": 11}
    with pytest.raises(TypeError):
        sl.process_adjsim_log(gen2_output=None, log_lines=[])
    with pytest.raises(TypeError):                                      # MidiGenerator's `range` quirk (:14,52) is still what instruments=None meets
        sl.process_adjsim_log(instruments=None, note_levels=np.arange(16), gen2_output=np.full(10, 0.5, dtype=np.float32), log_lines=[])
    with pytest.raises(ValueError, match="Error in processing log file"):
        sl.process_adjsim_log(note_levels=np.arange(16), gen2_output=np.full(10, 0.2, dtype=np.float32), log_lines=["INFO:root:1.0 - 4 - 99 - arrival\n"])
    with pytest.raises(ValueError, match="Error in processing log file"):
        sl.process_adjsim_log(note_levels=np.arange(16), gen2_output=np.full(10, 0.2, dtype=np.float32), log_path=str(tmp_path / "missing.log"))


def test_event_batch_packing(golden_dir):
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import sim_log_to_midi as sl
    cs = list(_cases(golden_dir))[:4]
    dt, meta, off = sl.sim_logs_to_event_batch([x[2] for x in cs], [x[1][x[0] + ".instruments"] for x in cs], [x[1][x[0] + ".note_levels"] for x in cs],
                                               [x[1][x[0] + ".gen2"] for x in cs])
    assert dt.dtype == torch.float64 and meta.dtype == torch.int32 and off.dtype == torch.int64 and off.numel() == 5
    for i, (name, c, _, generate, _, _) in enumerate(cs):
        if generate:
            continue
        a, b = int(off[i]), int(off[i + 1])
        assert np.array_equal(dt[a:b].numpy(), c[name + ".dt"]) and np.array_equal(meta[a:b].numpy().view(np.uint32), c[name + ".meta"])


@pytest.mark.gpu
def test_process_adjsim_log_device_rolls_match_reference(golden_dir):
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import sim_log_to_midi as sl
    batch = []
    for name, c, lines, generate, start, end in _cases(golden_dir):
        roll, dur, beats = sl.process_adjsim_log(instruments=c[name + ".instruments"], note_levels=c[name + ".note_levels"], gen2_output=c[name + ".gen2"],
                                                 start=start, end=end, generate=generate, log_lines=lines)
        assert roll.dtype == np.float64 and np.array_equal(roll, c[name + ".roll"]) and np.array_equal(dur, c[name + ".dur"]), name
        assert beats.shape == (50,)
        if (start, end) == (0, 50) and not generate:
            batch.append(name)
    # the batched path of the training loop: all songs in one H2D + one device rasterisation, uint8 rolls
    c = np.load(os.path.join(golden_dir, "simlog_cases.npz"))
    ev = sl.sim_logs_to_event_batch([[str(x) for x in c[n + ".lines"]] for n in batch], [c[n + ".instruments"] for n in batch],
                                    [c[n + ".note_levels"] for n in batch], [c[n + ".gen2"] for n in batch])
    out = ds.rasterize_events(*(t.cuda() for t in ev), 100, 0, 50, torch.uint8).cpu().numpy()
    for i, n in enumerate(batch):
        assert np.array_equal(out[i, 0], c[n + ".roll"]) and np.array_equal(out[i, 1], c[n + ".dur"]), n
