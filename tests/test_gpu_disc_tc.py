"""GPU parity of the bf16 tensor-core discriminator path (csrc/disc_tc*.cu) stage by stage against a torch
fp32 restatement with bf16-rounded operands (same rounding points: weights, conv1/conv2 activations,
gradient tensors), and end to end against the fp32 oracle within the bf16 tolerances of SURVEY 8(d):
logits 0.5 % of scale, loss rel 1e-3, gradients rel-L2 1e-2."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import mmgan_oracle as mo

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rb(x):
    """round to bf16, straight-through gradient"""
    return x + (x.bfloat16().float() - x).detach()


def _unpack_p(p, B, C=16):
    """(B*429, 4*C) S2D rows -> (B, C, 64, 25) NCHW"""
    t = p.float().view(B, 33, 13, 2, 2, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, 66, 26)
    return t[:, :, 1:65, 1:26]


def _rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("B,dtype", [(5, torch.uint8), (3, torch.float32), (37, torch.uint8), (300, torch.uint8)])
def test_disc_tc_stages(B, dtype, fused):
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200.disc_tc import DiscTC
    sd = mo.synth_state(mo.mmgan_shapes(), seed=31, d_scale=0.25)
    D = nt.DiscriminatorCNN(roll_size=(2, 128, 50)).to(DEV)
    D.load_state_dict({k[len("discriminator."):]: v for k, v in sd.items() if k.startswith("discriminator.")})
    tc = DiscTC(D, max_batch=B + 2, fused_backward=fused, fused_forward=fused)
    x8 = torch.from_numpy(mo.synth_rolls(B, 50, seed=32, p=0.05)).to(DEV)
    x = x8 if dtype == torch.uint8 else x8.float()
    logits = tc.forward(x).clone()
    torch.cuda.synchronize()

    # ---- torch restatement with the same rounding points
    w1, b1, w2, b2, wf, bf = (p.detach().clone().requires_grad_(True) for p in (D.conv1.weight, D.conv1.bias, D.conv2.weight, D.conv2.bias, D.fc.weight, D.fc.bias))
    a1 = _rb(F.leaky_relu(F.conv2d(x8.float(), _rb(w1), b1, stride=2, padding=1), 0.2))
    a2 = _rb(F.leaky_relu(F.conv2d(a1, _rb(w2), b2, stride=2, padding=1), 0.2))
    ref_logits = (a2.reshape(B, -1) @ wf.t() + bf).squeeze(1)

    got_a1 = _unpack_p(tc.p1[:B * 429], B)
    assert _rel_l2(got_a1, a1.detach()) < 2e-3, "conv1 activations"
    p1 = tc.p1[:B * 429].float().view(B, 33, 13, 2, 2, 16)
    assert not p1[:, 0, :, 0].any() and not p1[:, 32, :, 1].any() and not p1[:, :, 0, :, 0].any(), "P1 pad cells must stay zero"
    got_a2 = tc.a2[:B * 429].float().view(B, 33, 13, 32)
    assert not got_a2[:, 32].any() and not got_a2[:, :, 12].any(), "A2 junk rows must be zero"
    assert _rel_l2(got_a2[:, :32, :12].permute(0, 3, 1, 2), a2.detach()) < 3e-3, "conv2 activations"
    assert (logits - ref_logits.detach()).abs().max().item() <= 5e-3 * ref_logits.detach().abs().max().item() + 1e-4, "logits"

    # ---- backward with an injected dlogit
    dl = torch.from_numpy(np.random.default_rng(33).standard_normal(B).astype(np.float32)).to(DEV) / B
    for p in D.parameters():
        p.grad = None
    tc.backward(dl)
    torch.cuda.synchronize()
    ref = torch.autograd.grad(ref_logits, [w1, b1, w2, b2, wf, bf], dl)
    names = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "fc.weight", "fc.bias"]
    got = dict(D.named_parameters())
    for n, r in zip(names, ref):
        assert _rel_l2(got[n].grad, r) < 1e-2, (n, _rel_l2(got[n].grad, r))
    # backward accumulates (reference: the G-step backward adds onto the D-step grads)
    before = {n: got[n].grad.clone() for n in names}
    tc.backward(dl)
    torch.cuda.synchronize()
    for n in names:
        assert _rel_l2(got[n].grad, 2 * before[n]) < 1e-3, n


def test_disc_tc_matches_fp32_module():
    """same weights, same rolls: tensor-core logits vs the fp32 drop-in module (itself pinned to the reference)."""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200.disc_tc import DiscTC
    B = 16
    sd = mo.synth_state(mo.mmgan_shapes(), seed=3, d_scale=0.25)
    D = nt.DiscriminatorCNN(roll_size=(2, 128, 50)).to(DEV)
    D.load_state_dict({k[len("discriminator."):]: v for k, v in sd.items() if k.startswith("discriminator.")})
    x8 = torch.from_numpy(mo.synth_rolls(B, 50, seed=8)).to(DEV)
    with torch.no_grad():
        want = D(x8.float()).squeeze(1)
    got = DiscTC(D, max_batch=B).forward(x8)
    assert (got - want).abs().max().item() <= 5e-3 * want.abs().max().item() + 1e-4


def test_full_size_batch_properties():
    """BASELINE config 3 at its full per-GPU batch (16 384 rolls) through the fused kernels, checked with size-independent properties:
    samples are independent (the logits of the full batch equal the logits of its four quarters run separately, whatever CTA a sample lands
    on, up to the order in which a sample's eight per-warp fp32 partial dots are added), and the mean-loss gradient is additive over shards (full-batch gradient = sum of the quarter gradients, fp32 summation
    order only) -- the property the data-parallel all-reduce relies on."""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200.disc_tc import DiscTC
    B, Q = 16384, 4096
    sd = mo.synth_state(mo.mmgan_shapes(), seed=31, d_scale=0.25)
    D = nt.DiscriminatorCNN(roll_size=(2, 128, 50)).to(DEV)
    D.load_state_dict({k[len("discriminator."):]: v for k, v in sd.items() if k.startswith("discriminator.")})
    g = torch.Generator(device=DEV).manual_seed(5)
    x = ((torch.rand(B, 2, 128, 50, device=DEV, generator=g) < 0.02) * torch.randint(1, 128, (B, 2, 128, 50), device=DEV, generator=g)).to(torch.uint8)
    dl = torch.randn(B, device=DEV, generator=g) / B
    names = [n for n, _ in D.named_parameters()]

    def grads():
        return {n: p.grad.clone() for n, p in D.named_parameters()}

    tc = DiscTC(D, max_batch=B)
    for p in D.parameters():
        p.grad = None
    full_logits = tc.forward(x).clone()
    tc.backward(dl)
    torch.cuda.synchronize()
    full = grads()
    assert torch.isfinite(full_logits).all()
    scale = full_logits.abs().max().item()
    tcq = DiscTC(D, max_batch=Q)
    for p in D.parameters():
        p.grad = None
    for i in range(B // Q):
        lq = tcq.forward(x[i * Q:(i + 1) * Q])
        assert (lq - full_logits[i * Q:(i + 1) * Q]).abs().max().item() <= 1e-5 * scale + 1e-6, f"quarter {i}: logits depend on the batch"
        tcq.backward(dl[i * Q:(i + 1) * Q].contiguous())            # accumulates
    torch.cuda.synchronize()
    parts = grads()
    for n in names:
        assert _rel_l2(parts[n], full[n]) < 5e-4, (n, _rel_l2(parts[n], full[n]))      # fp32 accumulation order under heavy cancellation (zero-mean dlogits)
    # idempotence of the forward on the same buffers
    assert (tc.forward(x) - full_logits).abs().max().item() <= 1e-5 * scale + 1e-6


def test_forward_gathers_rows_by_index():
    """mmg_disc_fwd_fused_gather: the pass on rows x[index] of a resident set (duplicates and arbitrary order included) == the pass on the
    gathered copy, activations for the backward included."""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200.disc_tc import DiscTC
    sd = mo.synth_state(mo.mmgan_shapes(), seed=31, d_scale=0.25)
    D = nt.DiscriminatorCNN(roll_size=(2, 128, 50)).to(DEV)
    D.load_state_dict({k[len("discriminator."):]: v for k, v in sd.items() if k.startswith("discriminator.")})
    NDS, B = 500, 333
    for dtype in (torch.uint8, torch.float32):
        pool = torch.from_numpy(mo.synth_rolls(NDS, 50, seed=2, p=0.05)).to(DEV).to(dtype)
        idx = torch.randint(0, NDS, (B,), device=DEV)
        idx[:7] = 3                                                   # duplicates
        a = DiscTC(D, max_batch=B)
        want = a.forward(pool[idx].contiguous()).clone()
        want_p1, want_xs = a.p1[:B * 429].clone(), a.xs[:B * 1690].clone()
        b = DiscTC(D, max_batch=B)
        got = b.forward(pool, idx)
        assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item() + 1e-6
        assert torch.equal(b.p1[:B * 429], want_p1) and torch.equal(b.xs[:B * 1690], want_xs)
    with pytest.raises(ValueError):
        b.forward(pool, idx.int())
