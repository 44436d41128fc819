"""ORACLE (test infrastructure, not product code) -- MM-GAN / GAN-DES models and training step.

A CPU fp32, functional restatement (plain torch ops on a state dict; no nn.Module
classes) of the reference hot path:

  * MM-GAN models   /root/reference/MMGAN_MIDI_DES/network_tests.py:43-206
  * MM-GAN loop body /root/reference/MMGAN_MIDI_DES/network_tests.py:281-321
  * GAN-DES models  /root/reference/GAN_DES/SIMNN.py:37-142
  * GAN-DES loop body /root/reference/GAN_DES/SIMNN.py:275-334
  * Adam            torch/optim/adam.py (torch 2.11, single-tensor path)

Only ``tests/``, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` /
``--impl reference`` legs may import it.  It is pinned against the UNMODIFIED
reference classes (imported from /root/reference by ``oracle/make_golden.py`` in
the build container) through ``tests/golden/mmgan_*.npz`` / ``gandes_*.npz``;
``tests/test_oracle_models.py`` re-checks the oracle against those files.

The host DES (matrix_to_midi) is NOT part of the path: the fake piano rolls it
would return are inputs (``fake_d`` for the D step, ``fake_g`` for the G step).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ------------------------------------------------------------------------------------------
# deterministic synthetic state / inputs (numpy Generator, so both sides can rebuild them)
# ------------------------------------------------------------------------------------------
def mmgan_shapes(z_dim=50, hidden_dim=64, adj_size=(64, 64), roll_size=(2, 128, 50), input_dim=50, output_dim=20):
    """state-dict key -> shape, in the reference's registration order."""
    shapes = {}
    for g, in_extra, last in (("generator1", z_dim, adj_size[0] * adj_size[1]), ("generator2", input_dim, output_dim)):
        widths = [z_dim + in_extra, hidden_dim * 4, hidden_dim * 2, hidden_dim, last]
        for i in range(4):
            p = f"{g}.gen.{i}"
            shapes[f"{p}.0.weight"] = (widths[i + 1], widths[i])
            shapes[f"{p}.0.bias"] = (widths[i + 1],)
            shapes[f"{p}.1.weight"] = (widths[i + 1],)
            shapes[f"{p}.1.bias"] = (widths[i + 1],)
            shapes[f"{p}.1.running_mean"] = (widths[i + 1],)
            shapes[f"{p}.1.running_var"] = (widths[i + 1],)
            shapes[f"{p}.1.num_batches_tracked"] = ()
    hd = 16
    shapes["discriminator.conv1.weight"] = (hd, roll_size[0], 4, 4)
    shapes["discriminator.conv1.bias"] = (hd,)
    shapes["discriminator.conv2.weight"] = (2 * hd, hd, 4, 4)
    shapes["discriminator.conv2.bias"] = (2 * hd,)
    final = 2 * hd * (roll_size[1] // 4) * (roll_size[2] // 4)
    shapes["discriminator.fc.weight"] = (1, final)
    shapes["discriminator.fc.bias"] = (1,)
    return shapes


def synth_state(shapes, seed=0, d_scale=1.0):
    """Deterministic weights: N(0, 2/(fan_in+fan_out)) for matrices/filters, small random
    biases, BN affine around (1, 0), running stats perturbed so eval mode is exercised."""
    rng = np.random.default_rng(seed)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(0, dtype=torch.int64)
            continue
        if k.endswith("running_mean"):
            a = 0.1 * rng.standard_normal(shp)
        elif k.endswith("running_var"):
            a = 1.0 + 0.2 * rng.random(shp)
        elif ".1.weight" in k or "batch_norm" in k and k.endswith("weight"):
            a = 1.0 + 0.1 * rng.standard_normal(shp)
        elif len(shp) == 1:
            a = 0.05 * rng.standard_normal(shp)
        else:
            fan_out = shp[0] * int(np.prod(shp[2:])) if len(shp) > 2 else shp[0]
            fan_in = shp[1] * int(np.prod(shp[2:])) if len(shp) > 2 else shp[1]
            a = math.sqrt(2.0 / (fan_in + fan_out)) * rng.standard_normal(shp)
            if k.startswith("discriminator") or k.startswith("conv") or k.startswith("fc"):
                a = a * d_scale
        sd[k] = torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shp)).clone()
    return sd


def synth_rolls(B, W=50, seed=1, p=0.02):
    """(B,2,128,W) integer-valued rolls, each cell non-zero with
    probability p; channel 0 = velocity randint(1,128), channel 1 = duration randint(1,W+1)."""
    rng = np.random.default_rng(seed)
    mask = rng.random((B, 2, 128, W)) < p
    vel = rng.integers(1, 128, size=(B, 128, W))
    dur = rng.integers(1, W + 1, size=(B, 128, W))
    x = np.stack([vel, dur], axis=1) * mask
    return x.astype(np.uint8)


def synth_inputs(B, z_dim=50, input_dim=50, W=50, seed=2):
    """noise1, noise2, inner noise for the D-step and G-step forwards of generator1, beats, rolls."""
    rng = np.random.default_rng(seed)
    f = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
    return {
        "noise1": f(B, z_dim), "noise2": f(B, z_dim),
        "inner_d": f(B, z_dim), "inner_g": f(B, z_dim),
        "beats": torch.from_numpy((25.0 * rng.random((B, input_dim))).astype(np.float32)),
        "real": torch.from_numpy(synth_rolls(B, W, seed + 100)).float(),
        "fake_d": torch.from_numpy(synth_rolls(B, W, seed + 200)).float(),
        "fake_g": torch.from_numpy(synth_rolls(B, W, seed + 300)).float(),
    }


# ------------------------------------------------------------------------------------------
# layers, written out (not nn.Module calls) so the arithmetic being matched is explicit
# ------------------------------------------------------------------------------------------
def batch_norm(z, sd, p, training, dims):
    """nn.BatchNorm{1,2}d semantics (network_tests.py:78, SIMNN.py:85-87): training mode
    normalises with the batch mean and BIASED variance and updates running stats with
    momentum 0.1 using the UNBIASED variance; eval mode uses the running stats."""
    shape = [1, -1] + [1] * (z.dim() - 2)
    if training:
        n = z.numel() // z.shape[1]
        if n <= 1:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(z.shape)}")
        mean = z.mean(dim=dims)
        var = z.var(dim=dims, unbiased=False)
        with torch.no_grad():
            sd[p + ".running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach())
            sd[p + ".running_var"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var.detach() * (n / (n - 1)))
            sd[p + ".num_batches_tracked"] += 1
    else:
        mean, var = sd[p + ".running_mean"], sd[p + ".running_var"]
    inv = torch.rsqrt(var + BN_EPS)
    return (z - mean.view(shape)) * (inv * sd[p + ".weight"]).view(shape) + sd[p + ".bias"].view(shape)


def gen_forward(sd, g, noise, extra, training=True):
    """Generator / BeatGenerator (network_tests.py:58-123): cat -> 4 x [Linear, BN1d, Sigmoid]."""
    x = torch.cat((noise, extra), dim=1)                               # :86 / :122
    for i in range(4):
        p = f"{g}.gen.{i}"
        z = F.linear(x, sd[p + ".0.weight"], sd[p + ".0.bias"])        # :77
        x = torch.sigmoid(batch_norm(z, sd, p + ".1", training, (0,)))  # :78-79
    return x


def disc_forward(sd, image, prefix="discriminator"):
    """DiscriminatorCNN (network_tests.py:147-160)."""
    x = F.leaky_relu(F.conv2d(image, sd[prefix + ".conv1.weight"], sd[prefix + ".conv1.bias"], stride=2, padding=1), 0.2)
    x = F.leaky_relu(F.conv2d(x, sd[prefix + ".conv2.weight"], sd[prefix + ".conv2.bias"], stride=2, padding=1), 0.2)
    return F.linear(x.reshape(len(x), -1), sd[prefix + ".fc.weight"], sd[prefix + ".fc.bias"])


def disc_mlp_forward(sd, image, prefix="disc"):
    """Discriminator MLP (network_tests.py:126-144): 3 x [Linear, LeakyReLU(0.2)]."""
    x = image
    for i in range(3):
        x = F.leaky_relu(F.linear(x, sd[f"{prefix}.{i}.0.weight"], sd[f"{prefix}.{i}.0.bias"]), 0.2)
    return x


def bce_with_logits(x, y):
    """nn.BCEWithLogitsLoss(), mean reduction (network_tests.py:248)."""
    return (torch.clamp(x, min=0) - x * y + torch.log1p(torch.exp(-x.abs()))).mean()


def adam_step(params, grads, state, lr, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.Adam single-tensor update; params with grad None are skipped and get no state."""
    b1, b2 = betas
    for k, p in params.items():
        g = grads.get(k)
        if g is None:
            continue
        st = state.setdefault(k, {"step": 0, "m": torch.zeros_like(p), "v": torch.zeros_like(p)})
        st["step"] += 1
        t = st["step"]
        st["m"].lerp_(g, 1 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        step_size = lr / (1 - b1 ** t)
        denom = (st["v"].sqrt() / math.sqrt(1 - b2 ** t)).add_(eps)
        p.addcdiv_(st["m"], denom, value=-step_size)


D_KEYS = ["discriminator.conv1.weight", "discriminator.conv1.bias", "discriminator.conv2.weight",
          "discriminator.conv2.bias", "discriminator.fc.weight", "discriminator.fc.bias"]


def mmgan_iteration(sd, adam_state, inp, lr=0.01, d_grads_in=None):
    """One iteration of the reference loop body (network_tests.py:292-315), DES replaced by the
    given fake rolls.  ``sd`` and ``adam_state`` are updated in place.  Returns every observable:
    G outputs of both forwards, logits, losses, D grads after the D step (what Adam consumed) and
    after the G step (accumulated: gen_opt.zero_grad() does not clear D grads)."""
    out = {}
    params = {k: sd[k].detach().requires_grad_(True) for k in D_KEYS}
    dsd = dict(sd)
    dsd.update(params)
    B = len(inp["noise1"])
    ones, zeros = torch.ones(B), torch.zeros(B)

    # ---- D step (:293-308)
    with torch.no_grad():                      # graph is cut at matrix_to_midi (:189-193)
        out["g1_d"] = gen_forward(sd, "generator1", inp["noise1"], inp["inner_d"]).view(B, 1, *_adj(sd))
        out["g2_d"] = gen_forward(sd, "generator2", inp["noise2"], inp["beats"])
    logit_fake = disc_forward(dsd, inp["fake_d"])
    logit_real = disc_forward(dsd, inp["real"])
    loss_fake = bce_with_logits(logit_fake.squeeze(), zeros)            # :304
    loss_real = bce_with_logits(logit_real.squeeze(), ones)             # :305
    disc_loss = loss_fake + loss_real                                   # :306
    grads = torch.autograd.grad(disc_loss, [params[k] for k in D_KEYS])
    gd = {k: g.clone() for k, g in zip(D_KEYS, grads)}
    out.update({"logit_fake_d": logit_fake.detach(), "logit_real": logit_real.detach(), "disc_loss": disc_loss.detach()})
    for k in D_KEYS:
        out["grad_d." + k] = gd[k]
    with torch.no_grad():
        adam_step({k: sd[k] for k in D_KEYS}, gd, adam_state, lr)      # :308

    # ---- G step (:311-315): D forward with the UPDATED weights, grads accumulate onto gd
    params = {k: sd[k].detach().requires_grad_(True) for k in D_KEYS}
    dsd.update(params)
    with torch.no_grad():
        out["g1_g"] = gen_forward(sd, "generator1", inp["noise1"], inp["inner_g"]).view(B, 1, *_adj(sd))
        out["g2_g"] = gen_forward(sd, "generator2", inp["noise2"], inp["beats"])
    logit_g = disc_forward(dsd, inp["fake_g"])
    gen_loss = bce_with_logits(logit_g.squeeze(), ones)                # :313
    grads = torch.autograd.grad(gen_loss, [params[k] for k in D_KEYS])
    out.update({"logit_fake_g": logit_g.detach(), "gen_loss": gen_loss.detach()})
    for k, g in zip(D_KEYS, grads):
        out["grad_g." + k] = gd[k] + g                                  # what .grad holds after :314
    # gen_opt.step() (:315) is a no-op: every generator grad is None
    return out


def _adj(sd):
    n = sd["generator1.gen.3.0.weight"].shape[0]
    s = int(round(math.sqrt(n)))
    return (s, s)


# ------------------------------------------------------------------------------------------
# GAN-DES (GAN_DES/SIMNN.py)
# ------------------------------------------------------------------------------------------
def gandes_shapes(noise_dim=100, gen_dim=32):
    g = {"conv1.weight": (noise_dim, gen_dim * 4, 4, 4), "conv2.weight": (gen_dim * 4, gen_dim * 2, 4, 4),
         "conv3.weight": (gen_dim * 2, gen_dim, 4, 4), "conv4.weight": (gen_dim, 1, 5, 5)}
    for i, c in ((1, gen_dim * 4), (2, gen_dim * 2), (3, gen_dim)):
        for nm in ("weight", "bias", "running_mean", "running_var"):
            g[f"batch_norm{i}.{nm}"] = (c,)
        g[f"batch_norm{i}.num_batches_tracked"] = ()
    d = {"conv1.weight": (16, 1, 2, 2), "conv1.bias": (16,), "conv2.weight": (32, 16, 3, 3), "conv2.bias": (32,),
         "fc1.weight": (128, 32 * 32 * 54), "fc1.bias": (128,), "fc2.weight": (1, 128), "fc2.bias": (1,)}
    return g, d


def gandes_gen_forward(sd, noise, training=True):
    """SIMNN.py:97-112: ConvT(100->128,k4,s1,p0)+BN+ReLU, ConvT(k4,s2,p1)+BN+ReLU x2, ConvT(32->1,k5) -> sigmoid."""
    x = F.conv_transpose2d(noise, sd["conv1.weight"], stride=1, padding=0)
    x = torch.relu(batch_norm(x, sd, "batch_norm1", training, (0, 2, 3)))
    x = F.conv_transpose2d(x, sd["conv2.weight"], stride=2, padding=1)
    x = torch.relu(batch_norm(x, sd, "batch_norm2", training, (0, 2, 3)))
    x = F.conv_transpose2d(x, sd["conv3.weight"], stride=2, padding=1)
    x = torch.relu(batch_norm(x, sd, "batch_norm3", training, (0, 2, 3)))
    return torch.sigmoid(F.conv_transpose2d(x, sd["conv4.weight"], stride=1, padding=0))


def gandes_disc_forward(sd, x):
    """SIMNN.py:130-142 (output is already sigmoided -- the loop then applies BCEWithLogits to it)."""
    x = x.unsqueeze(1)
    x = F.max_pool2d(torch.relu(F.conv2d(x, sd["conv1.weight"], sd["conv1.bias"], stride=1, padding=1)), 2, 2)
    x = F.max_pool2d(torch.relu(F.conv2d(x, sd["conv2.weight"], sd["conv2.bias"], stride=1, padding=1)), 2, 2)
    x = torch.relu(F.linear(x.reshape(-1, 32 * 32 * 54), sd["fc1.weight"], sd["fc1.bias"]))
    return torch.sigmoid(F.linear(x, sd["fc2.weight"], sd["fc2.bias"]))


GD_KEYS = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"]


def gandes_iteration(gsd, dsd, adam_state, noise, real, fake, lr=2e-5, betas=(0.5, 0.999)):
    """SIMNN.py:280-331 with matrix_to_wav replaced by the given ``fake`` spectrograms:
    D step = BCE(disc(real), 0.9) + BCE(disc(fake.detach()), 0.1), Adam(lr 2e-5, betas (0.5, 0.999));
    G step = BCE(disc(fake), 1.0) -- reaches only D's grads (the bridge cuts the graph)."""
    out = {}
    B = len(real)
    with torch.no_grad():
        out["gen_out"] = gandes_gen_forward(gsd, noise, training=True)
    params = {k: dsd[k].detach().requires_grad_(True) for k in GD_KEYS}
    p_real = gandes_disc_forward(params, real)
    p_fake = gandes_disc_forward(params, fake)
    l_real = bce_with_logits(p_real, torch.full_like(p_real, 0.9))
    l_fake = bce_with_logits(p_fake, torch.full_like(p_fake, 0.1))
    d_loss = l_fake + l_real                                            # :312
    grads = torch.autograd.grad(d_loss, [params[k] for k in GD_KEYS])
    gd = dict(zip(GD_KEYS, grads))
    out.update({"p_real": p_real.detach(), "p_fake": p_fake.detach(), "disc_loss": d_loss.detach()})
    for k in GD_KEYS:
        out["grad_d." + k] = gd[k].clone()
    with torch.no_grad():
        adam_step({k: dsd[k] for k in GD_KEYS}, gd, adam_state, lr, betas)
    params = {k: dsd[k].detach().requires_grad_(True) for k in GD_KEYS}
    p_g = gandes_disc_forward(params, fake)
    g_loss = bce_with_logits(p_g, torch.ones_like(p_g))
    out.update({"p_fake_g": p_g.detach(), "gen_loss": g_loss.detach()})
    return out
