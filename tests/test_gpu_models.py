"""GPU parity of the fp32 path: the drop-in modules (CUDA kernels through the C ABI) run the
reference's loop body and are compared with the vectors frozen from the UNMODIFIED reference
(tests/golden/mmgan_*.npz, gandes_b3.npz).  Tolerances (SURVEY 8d, fp32 path): logits / losses /
generator outputs rel 2e-5 of scale; gradients and post-Adam weights rel-L2 1e-4."""
import os

import numpy as np
import pytest
import torch

import mmgan_oracle as mo

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(a, b, rtol=2e-5, atol=1e-6, what=""):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b).max() if a.size else 0.0
    assert err <= atol + rtol * max(np.abs(b).max(), 1e-30), (what, err, np.abs(b).max())


def _rel_l2(a, b, tol=1e-4, what=""):
    a = a.detach().cpu().double().numpy().ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    assert np.linalg.norm(a - b) <= tol * max(np.linalg.norm(b), 1e-12), (what, np.linalg.norm(a - b), np.linalg.norm(b))


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("fname", ["mmgan_b4_small.npz", "mmgan_b16.npz"])
def test_mmgan_loop_body(golden_dir, fname, fused):
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200 import optim as fo
    g = np.load(os.path.join(golden_dir, fname))
    B, adj, out_dim, seed, iters = (int(v) for v in g["meta"])
    sd0 = mo.synth_state(mo.mmgan_shapes(adj_size=(adj, adj), output_dim=out_dim), seed=seed, d_scale=0.25)
    mmgan = nt.MultiModalGAN(z_dim=50, adj_size=(adj, adj), roll_size=(2, 128, 50), input_dim=50, output_dim=out_dim, instrument=0,
                             start=100, end=150, device=DEV)
    mmgan.load_state_dict(sd0)
    criterion = fo.BCEWithLogitsLoss() if fused else torch.nn.BCEWithLogitsLoss()
    Adam = fo.FusedAdam if fused else torch.optim.Adam
    gen_opt = Adam(list(mmgan.generator1.parameters()) + list(mmgan.generator2.parameters()), lr=0.01)
    disc_opt = Adam(mmgan.discriminator.parameters(), lr=0.01)
    mmgan.train()
    D = mmgan.discriminator
    for it in range(iters):
        inp = {k: v.to(DEV) for k, v in mo.synth_inputs(B, seed=seed * 1000 + it).items()}
        pre = f"it{it}."
        real, fake_label = torch.ones(B, device=DEV), torch.zeros(B, device=DEV)
        # non-contiguous (B,2,128,W) view of a (2,B,128,W) buffer, as network_tests.py:290 builds it
        real_data = torch.stack([inp["real"][:, 0], inp["real"][:, 1]]).permute(1, 0, 2, 3)
        assert not real_data.is_contiguous()
        disc_opt.zero_grad()
        g1 = mmgan.generator1(inp["noise1"], torch.from_numpy(g[pre + "inner_d"]).to(DEV))
        g2 = mmgan.generator2(inp["noise2"], inp["beats"])
        _close(g2, g[pre + "g2_d"], what="g2_d")
        if pre + "g1_d" in g.files:
            _close(g1, g[pre + "g1_d"], what="g1_d")
        else:
            _close(g1[:, :, ::4, ::4], g[pre + "g1_d.sub"], what="g1_d")
        fake_output = D(inp["fake_d"])
        logit_real = D(real_data)
        disc_loss = criterion(fake_output.squeeze(), fake_label) + criterion(logit_real.squeeze(), real)
        disc_loss.backward()
        _close(fake_output, g[pre + "logit_fake_d"], atol=2e-6, what="logit_fake_d")
        _close(logit_real, g[pre + "logit_real"], atol=2e-6, what="logit_real")
        _close(disc_loss.reshape(1), g[pre + "disc_loss"].reshape(1), what="disc_loss")
        for k, p in D.named_parameters():
            _rel_l2(p.grad, g[pre + "grad_d.discriminator." + k], what="grad_d." + k)
        disc_opt.step()
        for k, p in D.named_parameters():
            _rel_l2(p, g[pre + "param_d.discriminator." + k], what="param_d." + k)
        gen_opt.zero_grad()
        g1 = mmgan.generator1(inp["noise1"], torch.from_numpy(g[pre + "inner_g"]).to(DEV))
        g2 = mmgan.generator2(inp["noise2"], inp["beats"])
        _close(g2, g[pre + "g2_g"], what="g2_g")
        fake_output = D(inp["fake_g"])
        gen_loss = criterion(fake_output.squeeze(), real)
        gen_loss.backward()
        gen_opt.step()
        assert all(p.grad is None for p in mmgan.generator1.parameters()) and len(gen_opt.state) == 0
        _close(fake_output, g[pre + "logit_fake_g"], atol=2e-6, what="logit_fake_g")
        _close(gen_loss.reshape(1), g[pre + "gen_loss"].reshape(1), what="gen_loss")
        for k, p in D.named_parameters():
            _rel_l2(p.grad, g[pre + "grad_g.discriminator." + k], what="grad_g." + k)
    sd = mmgan.state_dict()
    for k in g.files:
        if k.startswith("final."):
            _close(sd[k[6:]].float(), g[k].astype(np.float64), what=k)
    mmgan.generator1.eval(); mmgan.generator2.eval()
    inp = {k: v.to(DEV) for k, v in mo.synth_inputs(B, seed=seed * 1000 + 77).items()}
    with torch.no_grad():
        g1 = mmgan.generator1(inp["noise1"], torch.from_numpy(g["eval.inner"]).to(DEV))
        g2 = mmgan.generator2(inp["noise2"], inp["beats"])
    _close(g2, g["eval.g2"], what="eval.g2")
    _close(g1[:, :, ::4, ::4], g["eval.g1.sub"], what="eval.g1")


def test_multimodal_forward_plumbing():
    """MultiModalGAN.forward through an injected host bridge: (logits (B,1), failed count); the inner randn of
    generator1 consumes the global CPU RNG stream exactly like the reference (network_tests.py:83-84)."""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    B = 6
    sd0 = mo.synth_state(mo.mmgan_shapes(adj_size=(16, 16), output_dim=16), seed=5, d_scale=0.25)
    inp = mo.synth_inputs(B, seed=9)
    seen = {}

    def bridge(g1, g2, adj_size, instrument, start, end, count=0, generate=False):
        seen.update(g1=g1.shape, g2=g2.shape, dev=g1.device.type, count=count, se=(start, end), grad=g1.requires_grad)
        return [a.double().numpy() for a in inp["fake_d"]], 2

    m = nt.MultiModalGAN(z_dim=50, adj_size=(16, 16), roll_size=(2, 128, 50), input_dim=50, output_dim=16, instrument=0, start=100,
                         end=150, device=DEV, bridge=bridge)
    m.load_state_dict(sd0)
    m.train()
    torch.manual_seed(123)
    logits, failed = m(inp["noise1"].to(DEV), inp["noise2"].to(DEV), inp["beats"].to(DEV), 7, False)
    after = torch.randn(3)
    torch.manual_seed(123)
    inner = torch.randn(B, 50)
    assert torch.equal(after, torch.randn(3))
    assert failed == 2 and logits.shape == (B, 1) and seen == dict(g1=(B, 1, 16, 16), g2=(B, 16), dev="cuda", count=7, se=(100, 150), grad=False)
    inp["inner_d"] = inner
    ref = mo.mmgan_iteration(dict(sd0), {}, inp)
    _close(logits, ref["logit_fake_d"].numpy(), atol=2e-6)
    m2 = nt.MultiModalGAN(z_dim=50, adj_size=(16, 16), output_dim=16, device=DEV)
    with pytest.raises(RuntimeError, match="bridge is not set"):
        m2(inp["noise1"].to(DEV), inp["noise2"].to(DEV), inp["beats"].to(DEV), 1)
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        m.generator2(inp["noise2"][:1].to(DEV), inp["beats"][:1].to(DEV))


def test_generator_backward_injected_grad():
    """A13: the generator backward (never run by the reference loop) against torch autograd on the oracle."""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    B = 8
    shapes = mo.mmgan_shapes(adj_size=(16, 16), output_dim=16)
    sd0 = mo.synth_state(shapes, seed=21)
    inp = mo.synth_inputs(B, seed=22)
    gout = torch.from_numpy(np.random.default_rng(23).standard_normal((B, 16)).astype(np.float32))
    keys = [k for k in shapes if k.startswith("generator2") and k.endswith((".0.weight", ".0.bias", ".1.weight", ".1.bias"))]
    ref_sd = {k: (v.clone().requires_grad_(True) if k in keys else v.clone()) for k, v in sd0.items()}
    x_ref = inp["beats"].clone().requires_grad_(True)
    y = mo.gen_forward(ref_sd, "generator2", inp["noise2"], x_ref, training=True)
    grads = torch.autograd.grad(y, [ref_sd[k] for k in keys] + [x_ref], gout)
    m = nt.MultiModalGAN(z_dim=50, adj_size=(16, 16), output_dim=16, device=DEV)
    m.load_state_dict(sd0)
    m.train()
    x = inp["beats"].to(DEV).requires_grad_(True)
    m.generator2(inp["noise2"].to(DEV), x).backward(gout.to(DEV))
    named = dict(m.named_parameters())
    for k, gr in zip(keys, grads[:-1]):
        if k.endswith(".0.bias"):       # a bias in front of BatchNorm has an analytically zero gradient: both sides are rounding noise
            assert named[k].grad.abs().max().item() < 1e-5 and np.abs(gr.numpy()).max() < 1e-5, k
        else:
            _rel_l2(named[k].grad, gr.numpy(), tol=2e-4, what=k)
    _rel_l2(x.grad, grads[-1].numpy(), tol=2e-4, what="dx")


@pytest.mark.parametrize("fused", [False, True])
def test_gandes_loop_body(golden_dir, fused):
    from gan_des_midi_music_gen_b200.GAN_DES import SIMNN
    from gan_des_midi_music_gen_b200 import optim as fo
    g = np.load(os.path.join(golden_dir, "gandes_b3.npz"))
    B = int(g["meta"][0])
    gshapes, dshapes = mo.gandes_shapes()
    gen, disc = SIMNN.Generator().to(DEV), SIMNN.Discriminator().to(DEV)
    gen.load_state_dict(mo.synth_state(gshapes, seed=11)); disc.load_state_dict(mo.synth_state(dshapes, seed=12))
    criterion = fo.BCEWithLogitsLoss() if fused else torch.nn.BCEWithLogitsLoss()
    Adam = fo.FusedAdam if fused else torch.optim.Adam
    gen_opt = Adam(gen.parameters(), lr=2e-5, betas=(0.5, 0.999))
    disc_opt = Adam(disc.parameters(), lr=2e-5, betas=(0.5, 0.999))
    real, fake, noise = (torch.from_numpy(g[k]).to(DEV) for k in ("real", "fake", "noise"))
    disc_opt.zero_grad()
    p_real = disc(real).reshape(-1)
    l_real = criterion(p_real, torch.ones(B, device=DEV) * 0.9)
    gen_out = gen(noise)
    p_fake = disc(fake.detach()).reshape(-1)
    l_fake = criterion(p_fake, torch.ones(B, device=DEV) * 0.1)
    d_loss = l_fake + l_real
    d_loss.backward()
    _close(gen_out, g["gen_out"], what="gen_out")
    _close(p_real, g["p_real"], what="p_real"); _close(p_fake, g["p_fake"], what="p_fake")
    _close(d_loss.reshape(1), g["disc_loss"].reshape(1), what="disc_loss")
    sl = (slice(0, 128, 16), slice(0, None, 97))
    for k, p in disc.named_parameters():
        a = p.grad.double()
        want = g["grad_d." + k + ".sum"]
        assert abs(a.sum().item() - want[0]) <= 1e-4 * max(abs(want[0]), np.sqrt(want[1])), k
        assert abs((a * a).sum().item() - want[1]) <= 2e-4 * want[1], k
        _rel_l2(p.grad[sl] if k == "fc1.weight" else p.grad, g["grad_d." + k], tol=2e-4, what=k)
    disc_opt.step()
    for k, p in disc.named_parameters():
        _rel_l2(p[sl] if k == "fc1.weight" else p, g["param_d." + k], what=k)
    gen_opt.zero_grad()
    p_g = disc(fake).squeeze()
    g_loss = criterion(p_g, torch.ones(B, device=DEV))
    g_loss.backward()
    gen_opt.step()
    assert all(p.grad is None for p in gen.parameters())
    _close(p_g, g["p_fake_g"], what="p_fake_g"); _close(g_loss.reshape(1), g["gen_loss"].reshape(1), what="gen_loss")
    sd = gen.state_dict()
    for k in g.files:
        if k.startswith("final.gen."):
            _close(sd[k[10:]].float(), g[k].astype(np.float64), what=k)
    gen.eval()
    with torch.no_grad():
        _close(gen(noise), g["eval.gen_out"], what="eval.gen_out")


def test_mlp_discriminator_forward_backward_vs_oracle():
    """R8: the MLP `Discriminator` (network_tests.py:126-144; unused by MultiModalGAN, part of the module's API): forward on a flattened roll and the
    autograd backward of a BCE loss through it, against the oracle restatement (mmgan_oracle.disc_mlp_forward) with the same weights."""
    import torch.nn.functional as F
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    B, roll = 6, (2, 32, 10)
    torch.manual_seed(12)
    D = nt.Discriminator(im_chan=1, hidden_dim=16, roll_size=roll, device=DEV).to(DEV)
    assert [k for k in D.state_dict()] == [f"disc.{i}.0.{n}" for i in range(3) for n in ("weight", "bias")]
    sd = {k: v.detach().cpu().clone() for k, v in D.state_dict().items()}
    x = torch.from_numpy(mo.synth_rolls(B, roll[2], seed=3, p=0.2)[:, :, :roll[1], :]).float().reshape(B, -1)
    assert x.shape[1] == roll[0] * roll[1] * roll[2]
    # oracle, CPU fp32
    ps = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    want = mo.disc_mlp_forward(ps, x)
    loss_w = F.binary_cross_entropy_with_logits(want.squeeze(1), torch.ones(B))
    gw = torch.autograd.grad(loss_w, list(ps.values()))
    # this library's kernels
    got = D(x.to(DEV))
    assert got.shape == (B, 1)
    loss = F.binary_cross_entropy_with_logits(got.squeeze(1), torch.ones(B, device=DEV))
    loss.backward()
    torch.cuda.synchronize()
    scale = want.abs().max().item()
    assert (got.detach().cpu() - want.detach()).abs().max().item() <= 2e-5 * scale + 1e-6
    assert abs(loss.item() - loss_w.item()) <= 2e-5 * abs(loss_w.item())
    for (k, p), g in zip(D.named_parameters(), gw):
        err = ((p.grad.cpu() - g).norm() / g.norm().clamp_min(1e-30)).item()
        assert err <= 1e-4, (k, err)
    with pytest.raises(Exception):
        D(x)                                   # CPU tensor: no fallback


def test_modules_can_route_to_tensor_cores():
    """the drop-in nn.Modules reach the tcgen05 kernels without the trainer: DiscriminatorCNN.enable_tensor_cores (forward +
    autograd backward through the fused bf16 kernels) and Generator.enable_tensor_cores (under no_grad), against the fp32 modules (bf16 bars)."""
    import torch.nn.functional as F
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    B = 24
    sd = mo.synth_state(mo.mmgan_shapes(), seed=3, d_scale=0.25)
    m = nt.MultiModalGAN(z_dim=50, adj_size=(64, 64), roll_size=(2, 128, 50), input_dim=50, output_dim=20, instrument=0, start=100, end=150, device=DEV)
    m.load_state_dict(sd)
    m.train()
    inp = {k: v.to(DEV) for k, v in mo.synth_inputs(B, seed=5).items()}
    D = m.discriminator
    ones = torch.ones(B, device=DEV)
    want = D(inp["real"])
    F.binary_cross_entropy_with_logits(want.squeeze(1), ones).backward()
    gw = {k: p.grad.clone() for k, p in D.named_parameters()}
    for p in D.parameters():
        p.grad = None
    D.enable_tensor_cores(B)
    got = D(inp["real"])
    assert got.shape == (B, 1) and got.requires_grad
    other = D(inp["fake_d"])                       # a second forward before the backward (the reference's D step does this): own activation buffers
    F.binary_cross_entropy_with_logits(got.squeeze(1), ones).backward()
    assert torch.isfinite(other).all()
    torch.cuda.synchronize()
    assert (got - want).abs().max().item() <= 5e-3 * want.abs().max().item() + 1e-4
    for k, p in D.named_parameters():
        err = ((p.grad - gw[k]).norm() / gw[k].norm()).item()
        assert err <= 5e-2, (k, err)               # bf16 operands vs fp32 on this synthetic state (tests/_emul.py has the tight comparison)
    g2 = D(inp["real"].to(torch.uint8))            # uint8 rolls are accepted on this path
    assert (g2 - got).abs().max().item() <= 1e-5 * got.abs().max().item() + 1e-6
    D.enable_tensor_cores(B, enabled=False)
    assert (D(inp["real"]) - want).abs().max().item() <= 1e-6 * want.abs().max().item() + 1e-7
    # generators: eval-mode forward under no_grad
    m.eval()
    with torch.no_grad():
        w1 = m.generator1(inp["noise1"], inp["inner_d"])
        w2 = m.generator2(inp["noise2"], inp["beats"])
        m.generator1.enable_tensor_cores(B)
        m.generator2.enable_tensor_cores(B)
        y1 = m.generator1(inp["noise1"], inp["inner_d"])
        y2 = m.generator2(inp["noise2"], inp["beats"])
    assert y1.shape == w1.shape and (y1 - w1).abs().max().item() <= 2e-2 and (y2 - w2).abs().max().item() <= 2e-2
    y3 = m.generator2(inp["noise2"], inp["beats"])     # autograd on: the fp32 differentiable kernels
    assert y3.requires_grad and (y3 - w2).abs().max().item() <= 1e-5
