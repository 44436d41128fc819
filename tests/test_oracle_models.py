"""The model/training-step oracle (oracle/mmgan_oracle.py) against the vectors frozen from the
UNMODIFIED reference classes and loop body (tests/golden/mmgan_*.npz, gandes_b3.npz)."""
import os

import numpy as np
import pytest
import torch

import mmgan_oracle as mo

RTOL = 2e-5      # fp32 CPU vs fp32 CPU, different op order only


def _close(a, b, rtol=RTOL, atol=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b).max() if a.size else 0.0
    assert err <= atol + rtol * max(np.abs(b).max(), 1e-30), (err, np.abs(b).max())


@pytest.mark.parametrize("fname", ["mmgan_b4_small.npz", "mmgan_b16.npz"])
def test_mmgan_iterations(golden_dir, fname):
    torch.set_num_threads(1)
    g = np.load(os.path.join(golden_dir, fname))
    B, adj, out_dim, seed, iters = (int(v) for v in g["meta"])
    sd = mo.synth_state(mo.mmgan_shapes(adj_size=(adj, adj), output_dim=out_dim), seed=seed, d_scale=0.25)
    adam = {}
    for it in range(iters):
        inp = mo.synth_inputs(B, seed=seed * 1000 + it)
        inp["inner_d"] = torch.from_numpy(g[f"it{it}.inner_d"])
        inp["inner_g"] = torch.from_numpy(g[f"it{it}.inner_g"])
        out = mo.mmgan_iteration(sd, adam, inp, lr=0.01)
        pre = f"it{it}."
        for k in ("logit_fake_d", "logit_real", "logit_fake_g", "g2_d", "g2_g"):
            _close(out[k], g[pre + k])
        _close(out["disc_loss"], g[pre + "disc_loss"].reshape(()))
        _close(out["gen_loss"], g[pre + "gen_loss"].reshape(()))
        for k in mo.D_KEYS:
            _close(out["grad_d." + k], g[pre + "grad_d." + k], rtol=1e-4)
            _close(out["grad_g." + k], g[pre + "grad_g." + k], rtol=1e-4)
            _close(sd[k], g[pre + "param_d." + k], rtol=1e-4)
        for nm in ("g1_d", "g1_g"):
            if pre + nm in g.files:
                _close(out[nm], g[pre + nm])
            else:
                _close(out[nm][:, :, ::4, ::4], g[pre + nm + ".sub"])
                a = out[nm].double()
                _close([a.sum().item(), (a * a).sum().item()], g[pre + nm + ".sum"], rtol=1e-6)
    for k in g.files:
        if k.startswith("final."):
            _close(sd[k[6:]], g[k])
    # eval-mode generators (generate_midi path)
    inp = mo.synth_inputs(B, seed=seed * 1000 + 77)
    g1 = mo.gen_forward(sd, "generator1", inp["noise1"], torch.from_numpy(g["eval.inner"]), training=False).view(B, 1, adj, adj)
    g2 = mo.gen_forward(sd, "generator2", inp["noise2"], inp["beats"], training=False)
    _close(g2, g["eval.g2"])
    _close(g1[:, :, ::4, ::4], g["eval.g1.sub"])


def test_gandes_iteration(golden_dir):
    torch.set_num_threads(1)
    g = np.load(os.path.join(golden_dir, "gandes_b3.npz"))
    gshapes, dshapes = mo.gandes_shapes()
    gsd, dsd = mo.synth_state(gshapes, seed=11), mo.synth_state(dshapes, seed=12)
    out = mo.gandes_iteration(gsd, dsd, {}, torch.from_numpy(g["noise"]), torch.from_numpy(g["real"]), torch.from_numpy(g["fake"]))
    _close(out["gen_out"], g["gen_out"])
    _close(out["p_real"].reshape(-1), g["p_real"])
    _close(out["p_fake"].reshape(-1), g["p_fake"])
    _close(out["p_fake_g"].reshape(-1), g["p_fake_g"])
    _close(out["disc_loss"], g["disc_loss"].reshape(()))
    _close(out["gen_loss"], g["gen_loss"].reshape(()))
    sl = (slice(0, 128, 16), slice(0, None, 97))
    for k in mo.GD_KEYS:
        gr, pa = out["grad_d." + k], dsd[k]
        a = gr.double()
        _close([a.sum().item(), (a * a).sum().item()], g["grad_d." + k + ".sum"], rtol=1e-4)
        if k == "fc1.weight":
            gr, pa = gr[sl], pa[sl]
        _close(gr, g["grad_d." + k], rtol=1e-4)
        _close(pa, g["param_d." + k], rtol=1e-4)
    for k in g.files:
        if k.startswith("final.gen."):
            _close(gsd[k[10:]], g[k])
    _close(mo.gandes_gen_forward(gsd, torch.from_numpy(g["noise"]), training=False), g["eval.gen_out"])


def test_gandes_d_step_one_batch_of_2b_equals_two_passes():
    """The identity GANDESTrainer.d_step relies on (batch_d_passes): the discriminator (SIMNN.py:123-142) has no BatchNorm, so one pass over
    cat(real, fake) with targets [0.9 ... | 0.1 ...] and 2 x the mean over 2B gives the logits, the loss and the gradients of the reference's two
    passes (SIMNN.py:279-313).  Checked on the oracle in float64, where the only difference is the summation order."""
    _, dshapes = mo.gandes_shapes()
    dsd = {k: v.double() for k, v in mo.synth_state(dshapes, seed=12).items()}
    g = torch.Generator().manual_seed(5)
    B = 2
    real, fake = torch.randn(B, 128, 216, generator=g).double(), torch.randn(B, 128, 216, generator=g).double()
    res = []
    for batched in (False, True):
        params = {k: dsd[k].detach().clone().requires_grad_(True) for k in mo.GD_KEYS}
        if batched:
            p = mo.gandes_disc_forward(params, torch.cat([real, fake]))
            tgt = torch.cat([torch.full((B, 1), 0.9, dtype=torch.float64), torch.full((B, 1), 0.1, dtype=torch.float64)])
            loss = 2.0 * mo.bce_with_logits(p, tgt)
        else:
            p_real, p_fake = mo.gandes_disc_forward(params, real), mo.gandes_disc_forward(params, fake)
            loss = mo.bce_with_logits(p_fake, torch.full_like(p_fake, 0.1)) + mo.bce_with_logits(p_real, torch.full_like(p_real, 0.9))
            p = torch.cat([p_real, p_fake])
        grads = torch.autograd.grad(loss, [params[k] for k in mo.GD_KEYS])
        res.append((p.detach(), loss.detach(), grads))
    (p0, l0, g0), (p1, l1, g1) = res
    assert torch.allclose(p0, p1, rtol=0, atol=1e-13) and abs(l0.item() - l1.item()) < 1e-13
    for k, a, b in zip(mo.GD_KEYS, g0, g1):
        assert (a - b).abs().max().item() <= 1e-12 * max(1.0, a.abs().max().item()), k


def test_bn_rejects_single_sample():
    sd = mo.synth_state(mo.mmgan_shapes(adj_size=(16, 16)), seed=0)
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        mo.gen_forward(sd, "generator2", torch.zeros(1, 50), torch.zeros(1, 50), training=True)
