"""GPU parity of the bf16 tensor-core generator blocks (csrc/gen_tc.cu through the C ABI) against a plain PyTorch fp32
restatement of the reference blocks (network_tests.py:75-80: Linear -> BatchNorm1d -> Sigmoid) and against the golden
vectors frozen from the unmodified reference.  Tolerances: outputs abs 2e-2 (bf16 operands, fp32 accumulate),
running statistics rel 2e-2."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _torch_ref(sd, prefix, x, training, momentum=0.1, eps=1e-5):
    """fp32 CPU restatement; returns the output and the updated running stats"""
    stats = []
    for i in range(4):
        p = f"{prefix}.gen.{i}"
        z = F.linear(x, sd[p + ".0.weight"], sd[p + ".0.bias"])
        rm, rv = sd[p + ".1.running_mean"].clone(), sd[p + ".1.running_var"].clone()
        x = torch.sigmoid(F.batch_norm(z, rm, rv, sd[p + ".1.weight"], sd[p + ".1.bias"], training, momentum, eps))
        stats.append((rm, rv))
    return x, stats


@pytest.mark.parametrize("B,which", [(16, "generator1"), (300, "generator1"), (16, "generator2"), (1000, "generator2"), (129, "generator2")])
@pytest.mark.parametrize("training", [True, False])
def test_gen_tc_vs_torch_fp32(B, which, training):
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200.gen_tc import GenTC
    torch.manual_seed(B)
    m = nt.MultiModalGAN(z_dim=50, adj_size=(64, 64), roll_size=(2, 128, 50), input_dim=50, output_dim=20, instrument=0, start=100, end=150, device="cuda")
    g = getattr(m, which)
    with torch.no_grad():                    # non-trivial affine / running stats
        for blk in g.gen:
            blk[1].weight.uniform_(0.5, 1.5); blk[1].bias.uniform_(-0.5, 0.5)
            blk[1].running_mean.uniform_(-0.3, 0.3); blk[1].running_var.uniform_(0.5, 2.0)
            blk[0].bias.uniform_(-0.2, 0.2)
    g.train(training)
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    noise = torch.randn(B, 50)
    inp = torch.randn(B, 50) if which == "generator1" else 25 * torch.rand(B, 50)       # beats are raw seconds
    want, stats = _torch_ref(sd, which, torch.cat((noise, inp), 1), training)
    tc = GenTC(g, max_batch=B)
    got = tc.forward(noise.cuda(), inp.cuda())
    torch.cuda.synchronize()
    assert got.shape == want.shape
    err = (got.cpu() - want).abs().max().item()
    assert err < 2e-2, err
    assert (got.cpu() - want).abs().mean().item() < 3e-3
    for i, (rm, rv) in enumerate(stats):
        bn = g.gen[i][1]
        assert torch.allclose(bn.running_mean.cpu(), rm, rtol=2e-2, atol=2e-3), i
        assert torch.allclose(bn.running_var.cpu(), rv, rtol=2e-2, atol=2e-3), i
        assert int(bn.num_batches_tracked) == (1 if training else 0)


@pytest.mark.parametrize("training", [True, False])
def test_gen_tc_many_items_per_cta_and_worker_groups(training):
    """B = 2500: the output layer has 20 row tiles x 16 column groups = 320 items over 148 CTAs (contiguous ranges that cross a row tile, both
    accumulators, the cached column constants, a partial last row tile); 2 and 4 builder / epilogue warp groups give identical bits."""
    from gan_des_midi_music_gen_b200 import _native as N
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200.gen_tc import GenTC
    B = 2500
    torch.manual_seed(7)
    m = nt.MultiModalGAN(z_dim=50, adj_size=(64, 64), roll_size=(2, 128, 50), input_dim=50, output_dim=20, instrument=0, start=100, end=150, device="cuda")
    g = m.generator1
    with torch.no_grad():
        for blk in g.gen:
            blk[1].weight.uniform_(0.5, 1.5); blk[1].bias.uniform_(-0.5, 0.5)
            blk[1].running_mean.uniform_(-0.3, 0.3); blk[1].running_var.uniform_(0.5, 2.0)
            blk[0].bias.uniform_(-0.2, 0.2)
    g.train(training)
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    noise, inp = torch.randn(B, 50), torch.randn(B, 50)
    want, stats = _torch_ref({k: v.cpu() for k, v in sd0.items()}, "generator1", torch.cat((noise, inp), 1), training)
    outs = []
    prev = N.lib().mmg_gen_set_worker_groups(4)
    try:
        for groups in (4, 2):
            N.lib().mmg_gen_set_worker_groups(groups)
            m.load_state_dict(sd0)
            got = GenTC(g, max_batch=B).forward(noise.cuda(), inp.cuda())
            torch.cuda.synchronize()
            assert (got.cpu() - want).abs().max().item() < 2e-2 and (got.cpu() - want).abs().mean().item() < 3e-3
            for i, (rm, rv) in enumerate(stats):
                bn = g.gen[i][1]
                assert torch.allclose(bn.running_mean.cpu(), rm, rtol=2e-2, atol=2e-3), (groups, i)
                assert torch.allclose(bn.running_var.cpu(), rv, rtol=2e-2, atol=2e-3), (groups, i)
            outs.append(got.clone())
    finally:
        N.lib().mmg_gen_set_worker_groups(prev)
    if not training:                         # (train mode: the order of the fp64 atomics of the column sums is not fixed)
        assert torch.equal(outs[0], outs[1])
    else:
        assert (outs[0] - outs[1]).abs().max().item() < 1e-5


@pytest.mark.parametrize("B,which", [(16, "generator1"), (300, "generator2"), (2500, "generator1"), (2500, "generator2"), (16384, "generator1")])
def test_gen_tc_fused_hidden_matches_per_layer_launches(B, which):
    """Train mode: the three hidden blocks as ONE cooperative launch (mmg_gen_hidden_fused: pre-activations kept in TMEM, grid-wide barrier per
    layer, Gram partials of the output block's statistics) against the per-layer launches: same arithmetic per element, so outputs and running
    statistics agree to the order of the fp64 column-sum atomics; and against the torch fp32 restatement at the stated bf16 tolerance."""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200.gen_tc import GenTC
    torch.manual_seed(B + 1)
    m = nt.MultiModalGAN(z_dim=50, adj_size=(64, 64), roll_size=(2, 128, 50), input_dim=50, output_dim=20, instrument=0, start=100, end=150, device="cuda")
    g = getattr(m, which).train()
    with torch.no_grad():
        for blk in g.gen:
            blk[1].weight.uniform_(0.5, 1.5); blk[1].bias.uniform_(-0.5, 0.5)
            blk[1].running_mean.uniform_(-0.3, 0.3); blk[1].running_var.uniform_(0.5, 2.0)
            blk[0].bias.uniform_(-0.2, 0.2)
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    noise = torch.randn(B, 50)
    inp = torch.randn(B, 50) if which == "generator1" else 25 * torch.rand(B, 50)
    res = {}
    for fused in (False, True):
        m.load_state_dict(sd0)
        tc = GenTC(g, max_batch=B, fused_hidden=fused)
        assert tc.fused_hidden == fused and (not fused or tc._fused_supported(B))
        for _ in range(2):                                   # twice: the barrier word and the sums are re-zeroed per call
            got = tc.forward(noise.cuda(), inp.cuda())
        torch.cuda.synchronize()
        res[fused] = (got.clone(), [(blk[1].running_mean.clone(), blk[1].running_var.clone(), int(blk[1].num_batches_tracked)) for blk in g.gen])
    (y0, st0), (y1, st1) = res[False], res[True]
    assert torch.isfinite(y1).all() and (y0 - y1).abs().max().item() < 1e-5, (y0 - y1).abs().max().item()
    for (m0, v0, n0), (m1, v1, n1) in zip(st0, st1):
        assert n0 == n1 == 2 and torch.allclose(m0, m1, rtol=1e-5, atol=1e-6) and torch.allclose(v0, v1, rtol=1e-5, atol=1e-6)
    want, _ = _torch_ref({k: v.cpu() for k, v in sd0.items()}, which, torch.cat((noise, inp), 1), True)
    assert (y1.cpu() - want).abs().max().item() < 2e-2


def test_gen_tc_vs_reference_golden(golden_dir):
    """Generator outputs of the unmodified reference (mmgan_b16.npz, first D-step forward) within the stated bf16 tolerance."""
    import mmgan_oracle as mo
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200.gen_tc import GenTC
    g = np.load(os.path.join(golden_dir, "mmgan_b16.npz"))
    B, adj, out_dim, seed, iters = (int(v) for v in g["meta"])
    m = nt.MultiModalGAN(z_dim=50, adj_size=(adj, adj), roll_size=(2, 128, 50), input_dim=50, output_dim=out_dim, instrument=0, start=100, end=150, device="cuda")
    m.load_state_dict(mo.synth_state(mo.mmgan_shapes(adj_size=(adj, adj), output_dim=out_dim), seed=seed, d_scale=0.25))
    m.train()
    inp = {k: v.cuda() for k, v in mo.synth_inputs(B, seed=seed * 1000).items()}
    o1 = GenTC(m.generator1, B).forward(inp["noise1"], torch.from_numpy(g["it0.inner_d"]).cuda()).view(B, 1, adj, adj)
    o2 = GenTC(m.generator2, B).forward(inp["noise2"], inp["beats"])
    assert np.abs(o1.cpu().numpy()[:, :, ::4, ::4] - g["it0.g1_d.sub"]).max() < 2e-2
    assert np.abs(o2.cpu().numpy() - g["it0.g2_d"]).max() < 2e-2


def test_gen_tc_rejects_single_sample_in_training():
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200.gen_tc import GenTC
    g = nt.Generator(z_dim=50, adj_size=(64, 64), device="cuda").cuda().train()
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        GenTC(g, 4).forward(torch.randn(1, 50).cuda(), torch.randn(1, 50).cuda())


@pytest.mark.parametrize("B", [16, 333, 4096])
def test_gram_statistics_match_the_gemm_statistics_pass(B):
    """The wide output layer's batch sums from the 64 x 64 Gram matrix of its input (mmg_gen_layer_stats_gram) vs the GEMM statistics pass they
    replace: same bf16 operands, so sums, outputs and running statistics agree to accumulation order (fp32 tensor-core sums vs fp64)."""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200.gen_tc import GenTC
    out = {}
    for gram in (True, False):
        torch.manual_seed(7)
        m = nt.MultiModalGAN(z_dim=50, adj_size=(64, 64), roll_size=(2, 128, 50), input_dim=50, output_dim=20, instrument=0, start=100, end=150, device="cuda")
        g = m.generator1.train()
        with torch.no_grad():
            for blk in g.gen:
                blk[1].weight.uniform_(0.5, 1.5); blk[1].bias.uniform_(-0.5, 0.5); blk[0].bias.uniform_(-0.2, 0.2)
        torch.manual_seed(8)
        noise, inner = torch.randn(B, 50, device="cuda"), torch.randn(B, 50, device="cuda")
        tc = GenTC(g, max_batch=B, gram_stats=gram)
        assert tc.gram_stats == gram
        y = tc.forward(noise, inner).clone()
        torch.cuda.synchronize()
        out[gram] = (y, tc.sum_views[3].clone(), g.gen[3][1].running_mean.clone(), g.gen[3][1].running_var.clone(), g.gen[2][1].running_var.clone())
    a, b = out[True], out[False]
    n = 4096
    assert torch.allclose(a[1][:n], b[1][:n], rtol=1e-5, atol=1e-4 * B), "column sums"
    assert torch.allclose(a[1][n:], b[1][n:], rtol=1e-5, atol=1e-4 * B), "column sums of squares"
    assert (a[0] - b[0]).abs().max().item() < 1e-4
    assert torch.allclose(a[2], b[2], rtol=1e-5, atol=1e-6) and torch.allclose(a[3], b[3], rtol=1e-4, atol=1e-7)
    assert torch.equal(a[4], b[4]), "the previous layer's running statistics are updated exactly once either way"
