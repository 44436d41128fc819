"""GPU parity: CUDA rasteriser (through the C ABI) vs the golden vectors frozen from the unmodified
reference and vs the oracle; bit-exact (integer work)."""
import os

import numpy as np
import pytest
import torch

import raster_oracle as ro

pytestmark = pytest.mark.gpu


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.cuda()


PATH = {"v": None}


@pytest.fixture(autouse=True, params=["stream", "sort"])
def raster_path(request):
    """every test of this file runs through both kernel paths of csrc/raster.cu (warp-specialised stream kernel; chain + sort-by-pitch kernels)"""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
    PATH["v"], old = request.param, ds.RASTER_PATH
    ds.RASTER_PATH = request.param                 # the API-level tests (generate_piano_roll, preprocess_maestro) follow it too
    yield request.param
    PATH["v"], ds.RASTER_PATH = None, old


def _run(dt, meta, off, sl, start, end, out_dtype=torch.float32, status=False):
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
    return ds.rasterize_events(_dev(dt.astype(np.float64)), _dev(meta.astype(np.uint32).view(np.int32)), _dev(off.astype(np.int64)),
                               sl, start, end, out_dtype, status, path=PATH["v"])


def test_golden_cases(golden_dir):
    c = np.load(os.path.join(golden_dir, "raster_cases.npz"))
    for name in c["names"]:
        sl, start, end = (int(v) for v in c[name + ".args"])
        sl = None if sl < 0 else sl
        dt, meta = c[name + ".dt"], c[name + ".meta"]
        out = _run(dt, meta, np.array([0, len(dt)]), sl, start, end).cpu().numpy()
        assert out.shape[2:] == c[name + ".roll"].shape, name
        assert np.array_equal(out[0, 0], c[name + ".roll"]), name
        assert np.array_equal(out[0, 1], c[name + ".dur"]), name


def test_golden_cases_as_one_ragged_batch(golden_dir):
    """all default-argument cases in ONE launch (ragged offsets, an empty song included)."""
    c = np.load(os.path.join(golden_dir, "raster_cases.npz"))
    names = [n for n in c["names"] if tuple(c[n + ".args"]) == (100, 0, 50)]
    assert "empty" in names and len(names) >= 4
    dt = np.concatenate([c[n + ".dt"] for n in names])
    meta = np.concatenate([c[n + ".meta"] for n in names])
    off = np.concatenate([[0], np.cumsum([len(c[n + ".dt"]) for n in names])])
    out = _run(dt, meta, off, 100, 0, 50).cpu().numpy()
    for i, n in enumerate(names):
        assert np.array_equal(out[i, 0], c[n + ".roll"]) and np.array_equal(out[i, 1], c[n + ".dur"]), n


@pytest.mark.parametrize("args", [(100, 0, 50), (300, 0, 300), (40, 3, 33), (200, 10, 140), (None, 0, 17), (700, 0, 600)])
def test_random_ragged_vs_oracle(args):
    rng = np.random.default_rng(11)
    lens = [0, 1, 17, 0, 900, 33, 2500, 64, 4096]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    E = int(off[-1])
    dt = rng.exponential(0.07, size=E)
    dt[rng.random(E) < 0.2] = 0.5            # exact .5 boundaries
    meta = ro.pack_meta(rng.integers(0, 3, E), rng.integers(0, 128, E), rng.integers(0, 128, E))
    want, _ = ro.raster_batch_c(dt, meta, off, *args)
    got, st = _run(dt, meta, off, *args, status=True)
    assert np.array_equal(got.cpu().numpy(), want)
    assert not st.any()


def test_maestro_scale_full_size_vs_c_oracle():
    """BASELINE config 4 at full size: 1276 songs x 15000 messages, 300-step window, bit-exact."""
    dt, meta, off = ro.synth_songs(1276, 15000, 300.0, seed=0)
    want, notes = ro.raster_batch_c(dt, meta, off, 300, 0, 300, n_threads=os.cpu_count() or 1)
    got = _run(dt, meta, off, 300, 0, 300)
    g = got.cpu().numpy()
    assert np.array_equal(g, want)
    # size-independent properties: idempotence, and the uint8 output equals the saturated f32 output
    again = _run(dt, meta, off, 300, 0, 300)
    assert torch.equal(got, again)
    u8 = _run(dt, meta, off, 300, 0, 300, torch.uint8).cpu().numpy()
    assert np.array_equal(u8, np.minimum(want, 255).astype(np.uint8))
    assert notes > 3_000_000


def test_out_of_contract_inputs_are_flagged():
    dt = np.array([1.0, -3.0, 1.0])
    meta = ro.pack_meta([1, 1, 1], [60, 61, 62], [9, 9, 9])
    out, st = _run(dt, meta, np.array([0, 3]), 100, 0, 50, status=True)
    assert int(st[0]) & 1
    assert out[0, 0, 60, 1] == 9 and out[0, 0, 61].sum() == 0


def test_generate_piano_roll_api(golden_dir):
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
    c = np.load(os.path.join(golden_dir, "raster_cases.npz"))
    kind, pitch, vel = ro.unpack_meta(c["K1.meta"])
    s = ds.EventStream.from_arrays(c["K1.dt"], kind, pitch, vel, beats=[0.5, 1.0])
    roll, dur, beats = ds.generate_piano_roll(s)
    assert roll.dtype == np.float64 and roll.shape == (128, 50)
    assert np.array_equal(roll, c["K1.roll"]) and np.array_equal(dur, c["K1.dur"])
    assert beats.shape == (50,) and beats[1] == 1.0 and not beats[2:].any()       # K10
    roll4, _, _ = ds.generate_piano_roll(s, start=2, end=52)
    assert np.array_equal(roll4, c["K4.roll"])
    item = ds.MaestroDatasetMidi([s], 100, 50, device="cuda")[0]
    assert item[0].shape == (128, 50) and item[0].dtype == torch.float32 and item[0].is_cuda


def test_preprocess_maestro_matches_notebook_loop(tmp_path):
    """the MAESTRO pickling pipeline (notebook cells 10-11) with one batched device rasterisation, against the plain-Python
    restatement; includes a song shorter than one slice, one shorter than the window and ones that overrun it."""
    import pickle
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
    rng = np.random.default_rng(5)
    streams, ref_in = [], []
    for n_msg, total in ((0, 0.0), (40, 30.0), (900, 170.0), (3000, 420.0), (5000, 299.6), (2500, 1000.0)):
        dt = rng.exponential(total / max(n_msg, 1), size=n_msg)
        dt[rng.random(n_msg) < 0.1] = 0.5
        meta = ro.pack_meta(rng.integers(0, 3, n_msg), rng.integers(21, 109, n_msg), rng.integers(0, 128, n_msg))
        beats = np.sort(rng.random(rng.integers(0, 80))) * 60
        streams.append(ds.EventStream(dt, meta, beats))
        ref_in.append((dt, meta, beats))
    got = ds.preprocess_maestro(streams)
    want = ro.preprocess_reference(ref_in)
    assert len(got) == len(want) and len(got) >= 8
    for (gr, gd, gb), (wr, wd, wb) in zip(got, want):
        assert gr.dtype == torch.float32 and tuple(gr.shape) == (128, 50) and tuple(gb.shape) == (50,)
        assert np.array_equal(gr.numpy(), wr) and np.array_equal(gd.numpy(), wd) and np.array_equal(gb.numpy(), wb)
    p = tmp_path / "preprocessed_data_50.pkl"
    ds.save_preprocessed(got, p)
    back = pickle.load(open(p, "rb"))
    assert len(back) == len(got) and torch.equal(back[3][1], got[3][1])


def test_speculative_steps_match_sequential_chain_and_fall_back_when_needed():
    """csrc/raster.cu K1: time steps by a parallel prefix sum that is VERIFIED against an error bound, the exact sequential float64 chain only
    for the songs that fail the test.  (a) both modes give the C oracle's rolls on continuous dt, with (almost) no fall-backs; (b) streams whose
    prefix sums sit exactly on k + 1/2 (round-half-even territory) and streams that sit a few ulps next to k + 1/2 -- where a different
    summation order WOULD round differently -- must take the fall-back and still be bit-exact; (c) negative dt (outside the contract) falls back."""
    from gan_des_midi_music_gen_b200 import _native as N
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
    if PATH["v"] != "sort":
        pytest.skip("the speculative step kernel belongs to the workspace ('sort') path")
    rng = np.random.default_rng(123)

    def run(dt, meta, off, S, a, b, mode):
        N.lib().mmg_raster_set_mode(mode)
        try:
            ws = torch.zeros(ds.raster_workspace_bytes(len(off) - 1, len(dt)), dtype=torch.uint8, device="cuda")
            out = ds.rasterize_events(_dev(dt), _dev(meta.view(np.int32)), _dev(off), S, a, b, path="sort", workspace=ws)
            torch.cuda.synchronize()
            return out.cpu().numpy(), int(ws[-8:].view(torch.int64).item())
        finally:
            N.lib().mmg_raster_set_mode(0)

    # (a) continuous dt, long songs (several 512-message chunks, carries between chunks)
    dt, meta, off = ro.synth_songs(96, 6000, 300.0, seed=9)
    want, _ = ro.raster_batch_c(dt, meta, off, 300, 0, 300)
    got0, fb0 = run(dt, meta, off, 300, 0, 300, 0)
    got1, fb1 = run(dt, meta, off, 300, 0, 300, 1)
    assert np.array_equal(got0, want) and np.array_equal(got1, want)
    assert fb1 == 0 and fb0 <= 1, (fb0, fb1)                       # expected fall-back rate ~1e-5 per song
    # long ragged songs (several chunks, carries between chunks, a cut-off in the middle of a song, an empty song); every other song is made of
    # exact quarter / tenth steps, so its prefix sums keep landing on or next to k + 1/2: those songs must fall back, and everything stays bit-exact
    lens = [5000, 2049, 7777, 12, 0, 3000, 4097, 2600]
    parts = [rng.choice([0.5, 0.25, 0.1, 0.2], n) if k % 2 == 0 else rng.exponential(0.02, n) for k, n in enumerate(lens)]
    dtl = np.concatenate(parts)
    offl = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    kindl = rng.integers(0, 3, len(dtl)).astype(np.uint32)
    metal = (kindl | (rng.integers(21, 109, len(dtl)).astype(np.uint32) << 8) | (rng.integers(0, 128, len(dtl)).astype(np.uint32) << 16)).astype(np.uint32)
    metal[kindl == 0] = 0
    wantl, _ = ro.raster_batch_c(dtl, metal, offl, 300, 0, 300)
    g0, f0 = run(dtl, metal, offl, 300, 0, 300, 0)
    g1, _ = run(dtl, metal, offl, 300, 0, 300, 1)
    assert np.array_equal(g0, wantl) and np.array_equal(g1, wantl) and f0 >= 3, f0
    # (b) exact half-integers, and half-integers perturbed by a few ulps (0.1 + 0.2 + 0.2 style sums)
    songs = []
    for k in range(24):
        n = int(rng.integers(40, 1400))
        if k % 3 == 0:
            d = rng.integers(0, 4, n) * 0.5
        elif k % 3 == 1:
            d = rng.choice([0.1, 0.2, 0.3, 0.7, 0.15, 0.35], n)     # decimal fractions: prefix sums land within ulps of k + 1/2 again and again
        else:
            d = rng.choice([0.5, 0.25, 1.0 / 3.0, 1.0 / 6.0], n)
        songs.append(d.astype(np.float64))
    dt = np.concatenate(songs)
    off = np.concatenate([[0], np.cumsum([len(s) for s in songs])]).astype(np.int64)
    kind = rng.integers(0, 3, len(dt)).astype(np.uint32)
    meta = (kind | (rng.integers(21, 109, len(dt)).astype(np.uint32) << 8) | (rng.integers(0, 128, len(dt)).astype(np.uint32) << 16)).astype(np.uint32)
    meta[kind == 0] = 0
    want, _ = ro.raster_batch_c(dt, meta, off, 100, 0, 50)
    got0, fb0 = run(dt, meta, off, 100, 0, 50, 0)
    got1, _ = run(dt, meta, off, 100, 0, 50, 1)
    assert np.array_equal(got0, want) and np.array_equal(got1, want)
    assert fb0 >= 8, fb0                                           # at least the exact-half-integer songs went through the chain
    # (c) a negative dt: the bound does not cover it
    dt2, meta2, off2 = ro.synth_songs(4, 700, 60.0, seed=2)
    dt2[1000] = -0.25
    got0, fb0 = run(dt2, meta2, off2, 100, 0, 50, 0)
    got1, _ = run(dt2, meta2, off2, 100, 0, 50, 1)
    assert np.array_equal(got0, got1) and fb0 == 1                 # (decreasing time is outside the contract of the sort-by-pitch replay: only the step kernels are compared)
