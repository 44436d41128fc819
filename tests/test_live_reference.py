"""Live checks against the UNMODIFIED reference, run only where /root/reference exists (the build container; skipped on the GPU box):
many more random cases than the frozen golden files hold, for the two host-checkable pieces -- the raster oracle (numpy + C) against the
reference's generate_piano_roll, and the sim-log mirror against the reference's MidiGenerator / process_adjsim_log (mido build-side shim of
oracle/_refimport.py).  CPU only: the CUDA path is compared with the same oracle / golden files in the -m gpu tests."""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import pytest

import raster_oracle as ro

REF = os.environ.get("MMG_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "MMGAN_MIDI_DES")), reason="reference tree not present")

NAMES = {0: "control_change", 1: "note_on", 2: "note_off"}


@pytest.fixture(scope="module")
def ref():
    import _refimport as R
    sl = R.import_simlog()
    _, ds = R.import_mmgan()
    return R, ds, sl


def test_raster_oracle_vs_reference_random_streams(ref):
    R, ds, _ = ref
    rng = np.random.default_rng(77)
    for case in range(120):
        n = int(rng.integers(0, 400))
        sl = [100, 60, 300, None][case % 4]
        start, end = [(0, 50), (0, 50), (0, 300), (0, 50), (3, 33), (10, 140), (100, 150)][case % 7]
        scale = float(rng.choice([0.05, 0.3, 1.5]))
        dt = rng.exponential(scale, size=n)
        dt[rng.random(n) < 0.25] = 0.5                                  # exact .5 boundaries: round-half-even territory
        kind = rng.integers(0, 3, n)
        pitch = rng.integers(0, 128, n)
        vel = rng.integers(0, 128, n)
        ev = [R.ShimMessage(NAMES[int(k)], float(t), int(p), int(v)) for t, k, p, v in zip(dt, kind, pitch, vel)]
        with contextlib.redirect_stdout(io.StringIO()):
            roll, dur, _ = ds.generate_piano_roll(R.ShimMidiFile(ev), sequence_length=sl, beats_length=50, start=start, end=end)
        a, b = ro.raster_events(dt, kind, pitch, vel, sl, start, end)
        assert a.shape == roll.shape and np.array_equal(a, roll) and np.array_equal(b, dur), case
        c, _ = ro.raster_batch_c(dt, ro.pack_meta(kind, pitch, vel), np.array([0, n]), sl, start, end)
        assert np.array_equal(c[0, 0], roll) and np.array_equal(c[0, 1], dur), case


def test_simlog_mirror_vs_reference_random_logs(ref):
    R, _, m = ref
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from make_golden import synth_sim_log
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import sim_log_to_midi as sl
    rng = np.random.default_rng(5)
    cwd = os.getcwd()
    checked = 0
    for case in range(40):
        n_lines = int(rng.choice([100, 200, 300, 137, 500, 1200]))
        lines = synth_sim_log(rng, n_lines, float(rng.choice([20.0, 60.0, 150.0, 320.0])))
        gen2 = np.concatenate([rng.random(6) * np.array([1, 1, 1, 1, float(rng.choice([1.0, 1.0, 12.0])), 1]), rng.random(4)]).astype(np.float32)
        instruments, note_levels = rng.integers(0, 100, 16), rng.integers(30, 100, 16)
        generate = bool(case % 3 == 0)
        seen = {}
        real = m.generate_piano_roll

        def spy(midi, **kw):
            seen["msgs"] = [(x.type, float(x.time), int(getattr(x, "note", 0)), int(getattr(x, "velocity", 0))) for x in midi]
            return real(midi, **kw)

        m.generate_piano_roll = spy
        with tempfile.TemporaryDirectory() as td:
            os.makedirs(os.path.join(td, "logs"))
            os.makedirs(os.path.join(td, "adj_sim_outputs", "midi"))
            open(os.path.join(td, "logs", "simulation.log"), "w").writelines(lines)
            os.chdir(td)
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    roll, dur, _ = m.process_adjsim_log(instruments=instruments, note_levels=note_levels, gen2_output=gen2, start=0, end=50, generate=generate)
            finally:
                os.chdir(cwd)
                m.generate_piano_roll = real
        stream, _ = sl.sim_log_to_event_stream(lines, instruments, note_levels, gen2, generate)
        kinds = {"note_on": 1, "note_off": 2}
        want_dt = np.array([x[1] for x in seen["msgs"]], dtype=np.float64)
        want_meta = np.array([(kinds[x[0]] | (x[2] << 8) | (x[3] << 16)) if x[0] in kinds else 0 for x in seen["msgs"]], dtype=np.uint32)
        assert np.array_equal(stream.dt, want_dt) and np.array_equal(stream.meta, want_meta), case
        native = sl.sim_log_to_event_stream_native(lines, instruments, note_levels, gen2, generate)      # csrc/simlog.cu against the unmodified reference
        assert np.array_equal(native.dt, want_dt) and np.array_equal(native.meta, want_meta), case
        k, p, v = ro.unpack_meta(stream.meta)
        a, b = ro.raster_events(stream.dt, k, p, v, 100, 0, 50)
        assert np.array_equal(a, roll) and np.array_equal(b, dur), case
        checked += int((stream.meta != 0).any())
    assert checked >= 15
