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
