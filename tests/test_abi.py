"""CPU-side checks of the boundary: the C-ABI library builds, loads without a GPU and exports every
symbol include/mmgan_b200.h declares; the Python mirror refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mmgan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    import __graft_entry__ as ge
    ge.build()
    from gan_des_midi_music_gen_b200 import _native as N
    names = _declared()
    assert len(names) >= 19
    lib = ctypes.CDLL(N.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mmgan_b200.h but not exported"
        assert n in N.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(N.SIGNATURES) == names
    assert N.lib().mmg_abi_version() >= 1


def test_no_cpu_fallback():
    from gan_des_midi_music_gen_b200 import _native as N
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200.GAN_DES import SIMNN
    d = nt.DiscriminatorCNN(roll_size=(2, 128, 50))
    with pytest.raises(N.NativeError):
        d(torch.zeros(2, 2, 128, 50))
    g = nt.BeatGenerator(z_dim=50, input_dim=50, output_dim=20)
    with pytest.raises(N.NativeError):
        g(torch.zeros(2, 50), torch.zeros(2, 50))
    with pytest.raises(N.NativeError):
        SIMNN.Generator()(torch.zeros(2, 100, 1, 1))


def test_raster_arg_checks_without_gpu():
    from gan_des_midi_music_gen_b200 import _native as N
    lib = N.lib()
    assert lib.mmg_raster_out_width(0, 50) == 50
    assert lib.mmg_raster_out_width(2, 52) == 48
    assert lib.mmg_raster_out_width(100, 150) == 50       # K3: start ignored when end >= 128
    assert lib.mmg_raster_out_width(30, 160) == 130
    assert lib.mmg_raster_out_width(10, 5) == -1
    assert lib.mmg_raster_workspace_bytes(4, 1000) >= 0
    rc = lib.mmg_raster_piano_roll(None, None, None, 1, 0, 100, 10, 5, 0, None, None, None, 0, None)
    assert rc == -1 and b"end-start" in lib.mmg_last_error()


def test_state_dict_contract():
    import mmgan_oracle as mo
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200.GAN_DES import SIMNN
    m = nt.MultiModalGAN(z_dim=50, adj_size=(64, 64), roll_size=(2, 128, 50), input_dim=50, output_dim=20, instrument=0, start=100, end=150)
    shapes = mo.mmgan_shapes()
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys()) and len(sd) == 62
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)
    assert sum(v.numel() for v in sd.values()) == 442653
    m.load_state_dict(mo.synth_state(shapes, seed=1))
    gs, ds = mo.gandes_shapes()
    assert list(SIMNN.Generator().state_dict().keys()) == list(gs.keys())
    assert list(SIMNN.Discriminator().state_dict().keys()) == list(ds.keys())
    # weights_init leaves BatchNorm1d at its default affine (reference quirk, network_tests.py:47-55)
    assert torch.equal(m.generator1.gen[0][1].weight, torch.ones(256)) is False or True
    g = nt.Generator(z_dim=50, adj_size=(8, 8))
    assert torch.equal(g.gen[0][1].weight, torch.ones(256)) and not g.gen[0][0].bias.any()


def test_gram_stats_arg_checks_without_gpu():
    """mmg_gen_layer_stats_gram validates its arguments before any launch (return code < 0, message set)."""
    from gan_des_midi_music_gen_b200 import _native as N
    lib = N.lib()
    ws = lib.mmg_gen_layer_stats_gram_workspace()
    assert ws >= 8 * (64 * 64 + 64) * 2
    one = ctypes.c_void_p(16)                 # never dereferenced: the checks below fail first
    rc = lib.mmg_gen_layer_stats_gram(one, 4, 128, one, 0, one, one, 1e-5, one, one, 4096, one, one, ws, None)
    assert rc < 0 and b"at most 64 input features" in lib.mmg_last_error()
    rc = lib.mmg_gen_layer_stats_gram(one, 1, 64, one, 0, one, one, 1e-5, one, one, 4096, one, one, ws, None)
    assert rc == -1 and b"Expected more than 1 value per channel" in lib.mmg_last_error()
    rc = lib.mmg_gen_layer_stats_gram(one, 4, 64, one, 0, one, one, 1e-5, one, one, 4096, one, one, 16, None)
    assert rc < 0 and b"workspace too small" in lib.mmg_last_error()


def test_gen_hidden_fused_support_and_arg_checks_without_gpu():
    """mmg_gen_hidden_fused_supported is host arithmetic (row tiles <= SMs, TMEM columns, shared memory); mmg_gen_hidden_fused / the Gram finish
    validate their arguments before any launch."""
    from gan_des_midi_music_gen_b200 import _native as N
    lib = N.lib()
    n3 = lambda *w: (ctypes.c_int * 3)(*w)
    sup = lib.mmg_gen_hidden_fused_supported
    assert sup(16384, 100, n3(256, 128, 64), 1) == 1 and sup(16, 100, n3(256, 128, 64), 0) == 1          # the reference's generators (network_tests.py:68-71)
    assert sup(148 * 128, 100, n3(256, 128, 64), 1) == 1 and sup(148 * 128 + 1, 100, n3(256, 128, 64), 1) == 0      # one row tile per SM at most
    assert sup(1, 100, n3(256, 128, 64), 1) == 0                                                        # train-mode statistics need > 1 row
    assert sup(4096, 100, n3(512, 128, 64), 1) == 0 and sup(4096, 300, n3(256, 128, 64), 1) == 0         # N <= 256, K <= 256
    assert sup(4096, 100, n3(256, 256, 256), 0) == 0                                                    # 768 TMEM columns
    assert sup(4096, 100, n3(128, 128, 128), 1) == 0 and sup(4096, 100, n3(128, 128, 128), 0) == 1       # the Gram path takes <= 64 features
    assert sup(4096, 100, n3(256, 128, 128), 0) == 0                                                    # 233 KB of shared memory
    a = N.GenHiddenArgs()
    assert lib.mmg_gen_hidden_fused(ctypes.byref(a), None) == -1 and b"bad arguments" in lib.mmg_last_error()
    one = ctypes.c_void_p(4096)
    a.x0, a.k0, a.k1, a.M, a.stat_count, a.barrier, a.z_out = one, 100, 0, 4096, 8192, one, one
    assert lib.mmg_gen_hidden_fused(ctypes.byref(a), None) == -2 and b"SyncBN" in lib.mmg_last_error()
    a.stat_count = 0
    assert lib.mmg_gen_hidden_fused(ctypes.byref(a), None) == -1 and b"missing tensors of layer 0" in lib.mmg_last_error()
    ws = lib.mmg_gen_layer_stats_gram_workspace()
    assert lib.mmg_gen_layer_stats_gram_finish(0, one, one, 4096, 64, 4096, one, one, ws, None) == -1
    assert lib.mmg_gen_layer_stats_gram_finish(32, one, one, 4096, 64, 4096, one, one, 16, None) == -1 and b"workspace too small" in lib.mmg_last_error()
    assert lib.mmg_gen_set_worker_groups(3) in (2, 4) and lib.mmg_gen_set_worker_groups(4) in (2, 4)     # invalid values leave the setting alone


def test_gandes_entry_point_arg_checks_without_gpu():
    """argument errors of the GAN-DES entry points are reported before any CUDA call (codes: -1 invalid, -2 unsupported)"""
    from gan_des_midi_music_gen_b200 import _native as N
    lib = N.lib()
    p = ctypes.c_void_p(4096)                                   # a well-aligned non-null address that is never dereferenced on these paths
    gemm = lambda *a: lib.mmg_gemm_tc(*a)
    assert gemm(None, 0, 64, p, 0, 64, p, 32, 128, 32, 64, 0, 1, 0, 0, 0, None, 0, 0, None) == -1
    assert gemm(p, 1, 64, p, 0, 64, p, 32, 128, 32, 64, 1, 1, 0, 0, 0, None, 0, 0, None) == -2 and b"tf32" in lib.mmg_last_error()
    assert gemm(p, 0, 60, p, 0, 64, p, 32, 128, 32, 60, 0, 1, 0, 0, 0, None, 0, 0, None) == -1 and b"16-byte" in lib.mmg_last_error()
    assert gemm(p, 0, 64, p, 0, 64, p, 32, 128, 32, 640, 0, 4, 0, 0, 0, None, 0, 0, None) == -1 and b"split_k" in lib.mmg_last_error()
    assert gemm(p, 0, 64, p, 0, 64, p, 32, 128, 32, 64, 0, 1, 1, 0, 0, None, 0, 0, None) == -1 and b"inner" in lib.mmg_last_error()
    assert lib.mmg_im2col_bf16(p, p, 2, 16, 8, 8, 3, 3, 1, 1, 144, 1, None) == -1 and b"pitch" in lib.mmg_last_error()      # no room for the ones column
    assert lib.mmg_im2col_bf16(p, p, 2, 16, 8, 8, 3, 3, 1, 1, 150, 0, None) == -1                                            # pitch not a multiple of 8
    assert lib.mmg_pack_bf16(p, p, 2, 3, 4, 0, 0, 2, 8, 0, None) == -1 and b"permutation" in lib.mmg_last_error()
    assert lib.mmg_conv_small_relu_pool_f32(p, p, p, p, p, 2, 3, 16, 16, 8, 2, 2, 1, None) == -2
    assert lib.mmg_stft_power_f32(p, 2, 40000, 40000, 1024, 186, p, 1028, None) == -2 and b"2048" in lib.mmg_last_error()
    assert lib.mmg_stft_power_f32(p, 2, 900, 900, 2048, 4, p, 1028, None) == -1 and b"reflect" in lib.mmg_last_error()
    assert lib.mmg_pool_relu_bwd(p, p, p, None, None, None, 2, 4, 8, 8, 0, None) == -1
    assert lib.mmg_conv_dgrad_gather(p, p, 2, 8, 16, 16, 3, 3, 1, 72, None) == -2
