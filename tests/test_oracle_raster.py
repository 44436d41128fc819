"""The rasteriser oracle (numpy + C twins) against the vectors frozen from the UNMODIFIED
reference generate_piano_roll (tests/golden/raster_cases.npz, made by oracle/make_golden.py)."""
import os

import numpy as np
import pytest

import raster_oracle as ro


@pytest.fixture(scope="module")
def cases(golden_dir):
    return np.load(os.path.join(golden_dir, "raster_cases.npz"))


def _args(c, name):
    sl, start, end = (int(v) for v in c[name + ".args"])
    return (None if sl < 0 else sl), start, end


def test_all_cases_numpy(cases):
    for name in cases["names"]:
        sl, start, end = _args(cases, name)
        kind, pitch, vel = ro.unpack_meta(cases[name + ".meta"])
        roll, dur = ro.raster_events(cases[name + ".dt"], kind, pitch, vel, sl, start, end)
        assert roll.shape == cases[name + ".roll"].shape, name
        assert np.array_equal(roll, cases[name + ".roll"]), name
        assert np.array_equal(dur, cases[name + ".dur"]), name


def test_all_cases_c(cases):
    for name in cases["names"]:
        sl, start, end = _args(cases, name)
        dt, meta = cases[name + ".dt"], cases[name + ".meta"]
        out, notes = ro.raster_batch_c(dt, meta, np.array([0, len(dt)]), sl, start, end)
        assert out.shape[1:] == (2,) + cases[name + ".roll"].shape, name
        assert np.array_equal(out[0, 0], cases[name + ".roll"]), name
        assert np.array_equal(out[0, 1], cases[name + ".dur"]), name


def test_appendix_a_known_answers(cases):
    """SURVEY.md Appendix A spelled out (independent of the npz contents)."""
    r, d = cases["K1.roll"], cases["K1.dur"]
    assert r[60, 0] == 80 and r[64, 0] == 70 and r[72, 49] == 100 and np.count_nonzero(r) == 3
    assert (d[60, :2] == 2).all() and (d[64, :4] == 4).all() and np.count_nonzero(d) == 6
    assert cases["K4.roll"].shape == (128, 48) and cases["K4.roll"][72, 47] == 100
    assert np.array_equal(cases["K3.roll"], cases["K1.roll"])
    d5 = cases["K5.dur"]
    assert (d5[60, 3:10] == 9).all() and (d5[61, 8:10] == 12).all() and (d5[62, :10] == 25).all() and not d5[63].any()
    assert np.count_nonzero(cases["K6.dur"]) == 0 and cases["K6.roll"][60, 3] == 90
    assert (cases["K7.dur"][60, 3:10] == 26).all()
    assert cases["K8.roll"][60, 2] == 70 and (cases["K8.dur"][60, 4:7] == 3).all()


def test_c_matches_numpy_ragged():
    rng = np.random.default_rng(3)
    lens = [0, 1, 17, 0, 900, 33, 2500]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    E = int(off[-1])
    dt = rng.exponential(0.05, size=E)
    meta = ro.pack_meta(rng.integers(0, 3, E), rng.integers(0, 128, E), rng.integers(0, 128, E))
    for (sl, start, end) in [(100, 0, 50), (300, 0, 300), (40, 3, 33), (200, 10, 140)]:
        a = ro.raster_batch(dt, meta, off, sl, start, end)
        b, _ = ro.raster_batch_c(dt, meta, off, sl, start, end, n_threads=3)
        assert np.array_equal(a, b)


def test_beats_padding():
    assert np.array_equal(ro.pad_beats([0.5, 1.0], 5), [0.5, 1.0, 0, 0, 0])
    assert np.array_equal(ro.pad_beats(np.arange(9.0), 4), [0, 1, 2, 3])
