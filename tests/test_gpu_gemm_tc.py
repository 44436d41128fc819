"""GPU parity of the GAN-DES tensor-core path (csrc/gemm_tc.cu, functional_tc.py):
  (a) mmg_gemm_tc against a float64 matmul of the same bf16 (or fp32) operands -- every operand-major combination, M / N / K tails,
      NCHW / transposed stores, split-K accumulation, the tf32 kind;
  (b) the three autograd Functions (Linear, Conv2d, ConvTranspose2d: forward, data gradient, weight gradient) against torch's own fp32
      operators applied to the same bf16-rounded operands (outputs 2e-3 of scale, gradients rel-L2 1e-2: the kernels additionally round the
      incoming gradient to bf16), and -- for the layers without a ReLU, whose mask a rounded pre-activation can flip -- against plain fp32;
  (c) the GAN-DES loop body (SIMNN.py:275-334) with enable_tensor_cores() against the reference's golden run (tests/golden/gandes_b3.npz):
      outputs / losses at fixed bars; gradients tight against the bf16-rounding-point restatement and, against the fp32 reference, never
      worse than that restatement (ReLU masks of near-zero pre-activations flip under ANY bf16-operand arithmetic; with B = 3 samples
      every flip shows: DESIGN.md section 2)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import mmgan_oracle as mo

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _r8(n):
    return (n + 7) // 8 * 8


class _RoundBF16(torch.autograd.Function):
    """bf16 rounding point with a straight-through gradient: what the tensor-core path does to every GEMM operand."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


rb = _RoundBF16.apply


@pytest.mark.parametrize("M,N,K", [(128, 32, 64), (200, 30, 100), (145, 32, 1000), (128, 256, 64), (70, 300, 130), (1, 128, 30), (513, 64, 8)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_gemm_tc_all_majors(M, N, K, a_mn, b_mn):
    from gan_des_midi_music_gen_b200 import _native as Nt
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + a_mn * 2 + b_mn)
    A = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    B = torch.randn(N, K, generator=g).to(DEV).bfloat16()
    want = A.double() @ B.double().T

    def store(X, mn):          # K-major: [rows][pitch(K)]; MN-major: [K][pitch(rows)]
        X = X.T.contiguous() if mn else X
        buf = torch.zeros(X.shape[0], _r8(X.shape[1]), device=DEV, dtype=torch.bfloat16)
        buf[:, :X.shape[1]] = X
        return buf
    Ab, Bb = store(A, a_mn), store(B, b_mn)
    for trans, inner, split in ((0, 0, 1), (1, M, 1), (0, 0, 3), (1, 0, 1)):
        if trans and inner == 0:
            inner = next(d for d in (7, 5, 1) if M % d == 0)      # "NCHW" with `inner` pixels per image
        C = torch.zeros(M * N + 8, device=DEV)
        Nt.call("mmg_gemm_tc", Nt.ptr(Ab), a_mn, Ab.stride(0), Nt.ptr(Bb), b_mn, Bb.stride(0), Nt.ptr(C), N, M, N, K, 0, split, trans, inner, int(split > 1),
                None, 0, 0, Nt.stream())
        torch.cuda.synchronize()
        assert C[M * N:].abs().max().item() == 0                  # nothing written past the matrix
        got = C[:M * N]
        if trans:
            got = got.view(M // inner, N, inner).permute(0, 2, 1).reshape(M, N)
        else:
            got = got.view(M, N)
        assert _rel(got, want) < 2e-6, (trans, inner, split, _rel(got, want))


def test_gemm_tc_bias_act_and_tf32():
    from gan_des_midi_music_gen_b200 import _native as Nt
    M, N, K = 150, 40, 264
    g = torch.Generator(device="cpu").manual_seed(5)
    A, B = torch.randn(M, K, generator=g).to(DEV), torch.randn(N, K, generator=g).to(DEV)
    bias_n, bias_m = torch.randn(N, generator=g).to(DEV), torch.randn(M, generator=g).to(DEV)
    C = torch.empty(M, N, device=DEV)
    Nt.call("mmg_gemm_tc", Nt.ptr(A), 0, K, Nt.ptr(B), 0, K, Nt.ptr(C), N, M, N, K, 1, 1, 0, 0, 0, Nt.ptr(bias_n), 0, 2, Nt.stream())
    want = torch.relu(A.double() @ B.double().T + bias_n.double())
    assert _rel(C, want) < 2e-3                                   # tf32: 10-bit mantissas
    Ab, Bb = A.bfloat16().contiguous(), B.bfloat16().contiguous()
    Nt.call("mmg_gemm_tc", Nt.ptr(Ab), 0, K, Nt.ptr(Bb), 0, K, Nt.ptr(C), N, M, N, K, 0, 1, 0, 0, 0, Nt.ptr(bias_m), 1, 3, Nt.stream())
    want = torch.sigmoid(Ab.double() @ Bb.double().T + bias_m.double()[:, None])
    assert _rel(C, want) < 1e-5
    with pytest.raises(ValueError):                           # tf32 operands must be K-major
        Nt.call("mmg_gemm_tc", Nt.ptr(A), 1, K, Nt.ptr(B), 0, K, Nt.ptr(C), N, M, N, K, 1, 1, 0, 0, 0, None, 0, 0, Nt.stream())
    with pytest.raises(ValueError):                           # split-K without atomic accumulation
        Nt.call("mmg_gemm_tc", Nt.ptr(Ab), 0, K, Nt.ptr(Bb), 0, K, Nt.ptr(C), N, M, N, K, 0, 2, 0, 0, 0, None, 0, 0, Nt.stream())


@pytest.mark.parametrize("B,K,Nf,act", [(30, 55296, 128, 2), (5, 128, 1, 3), (40, 100, 72, 0)])
def test_linear_tc_vs_torch(B, K, Nf, act):
    from gan_des_midi_music_gen_b200 import functional_tc as T
    g = torch.Generator(device="cpu").manual_seed(B + K)
    x = torch.randn(B, K, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(Nf, K, generator=g) / K ** 0.5).to(DEV).requires_grad_(True)
    b = torch.randn(Nf, generator=g).to(DEV).requires_grad_(True)
    gy = torch.randn(B, Nf, generator=g).to(DEV)
    f = {0: lambda z: z, 2: torch.relu, 3: torch.sigmoid}[act]
    want = f(F.linear(rb(x), rb(w), b))
    gw = torch.autograd.grad(want, (x, w, b), gy)
    got = T.linear(x, w, b, act)
    gg = torch.autograd.grad(got, (x, w, b), gy)
    assert (got - want).abs().max().item() <= 2e-3 * want.abs().max().item()
    for a, c, n in zip(gg, gw, ("dx", "dw", "db")):
        assert _rel(a, c) < 1e-2, (n, _rel(a, c))
    if act != 2:
        g32 = torch.autograd.grad(f(F.linear(x, w, b)), (x, w, b), gy)
        for a, c, n in zip(gg, g32, ("dx", "dw", "db")):
            assert _rel(a, c) < 1.5e-2, (n, _rel(a, c))


@pytest.mark.parametrize("B,Ci,H,W,Co,k,pad,act", [(3, 16, 64, 108, 32, 3, 1, 2), (2, 1, 128, 216, 16, 2, 1, 2), (2, 5, 9, 11, 7, 3, 0, 0)])
def test_conv2d_tc_vs_torch(B, Ci, H, W, Co, k, pad, act):
    from gan_des_midi_music_gen_b200 import functional_tc as T
    g = torch.Generator(device="cpu").manual_seed(Ci + H)
    x = torch.randn(B, Ci, H, W, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5).to(DEV).requires_grad_(True)
    b = torch.randn(Co, generator=g).to(DEV).requires_grad_(True)
    f = torch.relu if act == 2 else (lambda z: z)
    want = f(F.conv2d(rb(x), rb(w), rb(b), 1, pad))             # the bias rides the GEMM as a bf16 weight column
    gy = torch.randn(want.shape, generator=g).to(DEV)
    gw = torch.autograd.grad(want, (x, w, b), gy)
    got = T.conv2d(x, w, b, 1, pad, act)
    assert got.shape == want.shape
    gg = torch.autograd.grad(got, (x, w, b), gy)
    assert (got - want).abs().max().item() <= 2e-3 * want.abs().max().item()
    for a, c, n in zip(gg, gw, ("dx", "dw", "db")):
        assert _rel(a, c) < 1e-2, (n, _rel(a, c))
    if act != 2:
        g32 = torch.autograd.grad(F.conv2d(x, w, b, 1, pad), (x, w, b), gy)
        for a, c, n in zip(gg, g32, ("dx", "dw", "db")):
            assert _rel(a, c) < 1.5e-2, (n, _rel(a, c))


@pytest.mark.parametrize("B,Ci,H,W,Co,k,pad", [(2, 1, 128, 216, 16, 2, 1), (3, 16, 64, 108, 32, 3, 1), (2, 3, 9, 7, 5, 3, 1)])
def test_conv2d_relu_pool_tc_vs_torch(B, Ci, H, W, Co, k, pad):
    """conv + ReLU + maxpool as one node (forward through the GEMM, backward through the fused pool / ReLU kernel), odd pre-pool sizes included"""
    from gan_des_midi_music_gen_b200 import functional_tc as T
    g = torch.Generator(device="cpu").manual_seed(Ci + H + W)
    x = torch.randn(B, Ci, H, W, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(Co, Ci, k, k, generator=g) / (Ci * k * k) ** 0.5).to(DEV).requires_grad_(True)
    b = torch.randn(Co, generator=g).to(DEV).requires_grad_(True)
    stencil = (Ci, k) == (1, 2)          # K = 4: the fused fp32 stencil kernel computes the forward (no operand rounding); its weight gradient is a GEMM again
    r = (lambda t: t) if stencil else rb
    want = F.max_pool2d(torch.relu(F.conv2d(r(x), r(w), r(b), 1, pad)), 2)
    gy = torch.randn(want.shape, generator=g).to(DEV)
    gw = torch.autograd.grad(want, (x, w, b), gy)
    got = T.conv2d_relu_pool(x, w, b, pad)
    assert got.shape == want.shape
    gg = torch.autograd.grad(got, (x, w, b), gy)
    assert (got - want).abs().max().item() <= (1e-5 if stencil else 2e-3) * want.abs().max().item()
    for a, c, n in zip(gg, gw, ("dx", "dw", "db")):
        assert _rel(a, c) < 1e-2, (n, _rel(a, c))
    # without a data gradient (the discriminator's first block) only the bf16 transposed gradient is produced
    got2 = T.conv2d_relu_pool(x.detach(), w, b, pad)
    g2 = torch.autograd.grad(got2, (w, b), gy)
    for a, c, n in zip(g2, gw[1:], ("dw", "db")):
        assert _rel(a, c) < 1e-2, (n, _rel(a, c))


@pytest.mark.parametrize("B,Ci,Hin,Co,k,s,pad,act", [(30, 100, 1, 128, 4, 1, 0, 0), (30, 128, 4, 64, 4, 2, 1, 0), (3, 64, 8, 32, 4, 2, 1, 0), (3, 32, 16, 1, 5, 1, 0, 3)])
def test_conv_transpose2d_tc_vs_torch(B, Ci, Hin, Co, k, s, pad, act):
    from gan_des_midi_music_gen_b200 import functional_tc as T
    g = torch.Generator(device="cpu").manual_seed(Ci + Hin)
    x = torch.randn(B, Ci, Hin, Hin, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(Ci, Co, k, k, generator=g) / Ci ** 0.5).to(DEV).requires_grad_(True)
    f = torch.sigmoid if act == 3 else (lambda z: z)
    want = f(F.conv_transpose2d(rb(x), rb(w), None, s, pad))
    gy = torch.randn(want.shape, generator=g).to(DEV)
    gw = torch.autograd.grad(want, (x, w), gy)
    got = T.conv_transpose2d(x, w, s, pad, act)
    assert got.shape == want.shape
    gg = torch.autograd.grad(got, (x, w), gy)
    assert (got - want).abs().max().item() <= 2e-3 * want.abs().max().item()
    for a, c, n in zip(gg, gw, ("dx", "dw")):
        assert _rel(a, c) < 1e-2, (n, _rel(a, c))
    g32 = torch.autograd.grad(f(F.conv_transpose2d(x, w, None, s, pad)), (x, w), gy)
    for a, c, n in zip(gg, g32, ("dx", "dw")):
        assert _rel(a, c) < 1.5e-2, (n, _rel(a, c))


def test_gandes_loop_body_on_tensor_cores(golden_dir):
    """The loop body of SIMNN.py:275-334 with every contraction on tcgen05 against the reference's own run (gandes_b3.npz): bf16-operand bars
    (outputs 2e-2 abs on sigmoid outputs, loss rel 1e-3, gradients rel-L2 2e-2), and the packed-weight cache follows the optimiser step."""
    from gan_des_midi_music_gen_b200.GAN_DES import SIMNN
    from gan_des_midi_music_gen_b200 import optim as fo
    g = np.load(os.path.join(golden_dir, "gandes_b3.npz"))
    B = int(g["meta"][0])
    gshapes, dshapes = mo.gandes_shapes()
    gen, disc = SIMNN.Generator().to(DEV).enable_tensor_cores(), SIMNN.Discriminator().to(DEV).enable_tensor_cores()
    gen.load_state_dict(mo.synth_state(gshapes, seed=11)); disc.load_state_dict(mo.synth_state(dshapes, seed=12))
    assert list(disc.state_dict()) == list(SIMNN.Discriminator().state_dict())
    criterion = fo.BCEWithLogitsLoss()
    gen_opt = fo.FusedAdam(gen.parameters(), lr=2e-5, betas=(0.5, 0.999))
    disc_opt = fo.FusedAdam(disc.parameters(), lr=2e-5, betas=(0.5, 0.999))
    real, fake, noise = (torch.from_numpy(g[k]).to(DEV) for k in ("real", "fake", "noise"))
    disc_opt.zero_grad()
    p_real = disc(real).reshape(-1)
    l_real = criterion(p_real, torch.ones(B, device=DEV) * 0.9)
    gen_out = gen(noise)
    p_fake = disc(fake.detach()).reshape(-1)
    d_loss = criterion(p_fake, torch.ones(B, device=DEV) * 0.1) + l_real
    d_loss.backward()
    t = lambda k: torch.from_numpy(np.asarray(g[k])).to(DEV)
    assert (gen_out - t("gen_out")).abs().max().item() < 2e-2
    assert (p_real - t("p_real")).abs().max().item() < 5e-3 and (p_fake - t("p_fake")).abs().max().item() < 5e-3
    assert abs(d_loss.item() - float(g["disc_loss"])) <= 1e-3 * abs(float(g["disc_loss"]))
    sl = (slice(0, 128, 16), slice(0, None, 97))
    # the bf16-rounding-point restatement of the D step: torch fp32 operators on bf16-rounded GEMM operands
    ps = {k: p.detach().clone().requires_grad_(True) for k, p in disc.named_parameters()}

    def emu(x):
        x = F.max_pool2d(torch.relu(F.conv2d(x.unsqueeze(1), ps["conv1.weight"], ps["conv1.bias"], 1, 1)), 2)      # fp32 stencil kernel: no rounding point
        x = F.max_pool2d(torch.relu(F.conv2d(rb(x), rb(ps["conv2.weight"]), rb(ps["conv2.bias"]), 1, 1)), 2)
        x = torch.relu(F.linear(rb(x.reshape(-1, 32 * 32 * 54)), rb(ps["fc1.weight"]), ps["fc1.bias"]))
        return torch.sigmoid(F.linear(rb(x), rb(ps["fc2.weight"]), ps["fc2.bias"])).reshape(-1)
    e_loss = F.binary_cross_entropy_with_logits(emu(real), torch.full((B,), 0.9, device=DEV)) + \
        F.binary_cross_entropy_with_logits(emu(fake), torch.full((B,), 0.1, device=DEV))
    eg = dict(zip(ps, torch.autograd.grad(e_loss, list(ps.values()))))
    rep = {}
    for k, p in disc.named_parameters():
        cut = (lambda a: a[sl]) if k == "fc1.weight" else (lambda a: a)
        want = t("grad_d." + k)
        k_e, e_o, k_o = _rel(p.grad, eg[k]), _rel(cut(eg[k]), want), _rel(cut(p.grad), want)
        rep[k] = (k_e, e_o, k_o)
        assert k_e < 1e-2, (k, rep[k])
        assert k_o <= max(2e-2, 1.25 * e_o + k_e), (k, rep[k])
    print("GAN-DES D-step gradient rel-L2 (kernels vs bf16 restatement, restatement vs fp32 reference, kernels vs fp32 reference):",
          {k: tuple(f"{x:.1e}" for x in v) for k, v in rep.items()})
    w_before = disc.fc1.weight.detach().clone()
    disc_opt.step()
    assert not torch.equal(w_before, disc.fc1.weight)
    gen_opt.zero_grad()
    p_g = disc(fake).squeeze()                     # must see the updated weights (cache invalidated by the optimiser)
    g_loss = criterion(p_g, torch.ones(B, device=DEV))
    g_loss.backward()
    assert (p_g - t("p_fake_g")).abs().max().item() < 5e-3
    assert abs(g_loss.item() - float(g["gen_loss"])) <= 1e-3 * abs(float(g["gen_loss"]))
    # the generator's backward through the tensor-core transposed convolutions (training-mode BatchNorm + ReLU between them) against the
    # same restatement: torch operators on bf16-rounded GEMM operands
    gy = torch.randn(B, 1, 20, 20, device=DEV)
    gen.train()
    gen.zero_grad()
    gen(noise).backward(gy)
    gp = {k: p.detach().clone().requires_grad_(True) for k, p in gen.named_parameters()}
    x = noise
    for i, (st, pd) in enumerate(((1, 0), (2, 1), (2, 1)), 1):
        x = F.conv_transpose2d(rb(x), rb(gp[f"conv{i}.weight"]), None, st, pd)
        x = torch.relu(F.batch_norm(x, None, None, gp[f"batch_norm{i}.weight"], gp[f"batch_norm{i}.bias"], True, 0.1, 1e-5))
    out = torch.sigmoid(F.conv_transpose2d(rb(x), rb(gp["conv4.weight"]), None, 1, 0))
    eg = dict(zip(gp, torch.autograd.grad(out, list(gp.values()), gy)))
    rep = {k: _rel(p.grad, eg[k]) for k, p in gen.named_parameters()}
    print("GAN-DES generator gradient rel-L2, kernels vs bf16 restatement:", {k: f"{v:.1e}" for k, v in rep.items()})
    assert max(rep.values()) < 1.5e-2, rep


@pytest.mark.parametrize("tcores,batched", [(False, True), (True, True), (True, False)])
def test_gandes_trainer_graph_replay_matches_module_loop(tcores, batched):
    """GANDESTrainer (segments of SIMNN.py:275-334, captured into CUDA graphs on the second call) against the same loop written with the
    modules, FusedAdam and the fused BCE: identical kernels, so losses agree to fp32 atomics noise and the weights follow the same trajectory.
    ``batched``: the trainer's D step sends real and fake through the discriminator as one batch of 2B (no BatchNorm in D: same logits, same gradients)."""
    from gan_des_midi_music_gen_b200.GAN_DES import SIMNN
    from gan_des_midi_music_gen_b200 import optim as fo
    from gan_des_midi_music_gen_b200.gandes_trainer import GANDESTrainer
    B = 4
    gshapes, dshapes = mo.gandes_shapes()
    g = torch.Generator().manual_seed(21)
    noise, real, fake = (torch.randn(B, 100, 1, 1, generator=g).to(DEV), torch.randn(B, 128, 216, generator=g).to(DEV),
                         torch.randn(B, 128, 216, generator=g).to(DEV))
    out = {}
    for mode in ("loop", "trainer"):
        gen, disc = SIMNN.Generator().to(DEV).enable_tensor_cores(tcores), SIMNN.Discriminator().to(DEV).enable_tensor_cores(tcores)
        gen.load_state_dict(mo.synth_state(gshapes, seed=11)); disc.load_state_dict(mo.synth_state(dshapes, seed=12))
        losses = []
        if mode == "loop":
            crit = fo.BCEWithLogitsLoss()
            disc_opt = fo.FusedAdam(disc.parameters(), lr=2e-4, betas=(0.5, 0.999))
            for it in range(5):
                disc_opt.zero_grad()
                l_real = crit(disc(real).reshape(-1), torch.full((B,), 0.9, device=DEV))
                with torch.no_grad():
                    gm = gen(noise)
                l_fake = crit(disc(fake.detach()).reshape(-1), torch.full((B,), 0.1, device=DEV))
                dl = l_fake + l_real
                dl.backward()
                disc_opt.step()
                gl = crit(disc(fake).squeeze(), torch.ones(B, device=DEV))
                losses.append((dl.item(), gl.item()))
        else:
            tr = GANDESTrainer(gen, disc, lr=2e-4, betas=(0.5, 0.999), batch_d_passes=batched)
            for it in range(5):
                gm = tr.generate(noise)
                dl = tr.d_step(real, fake)
                gl = tr.g_step(fake)
                losses.append((dl.item(), gl.item()))
            assert {k[0] for k in tr._graphs} == {"gen", "d", "g"} and tr.replayed_launches > 0
            assert int(tr.adam_step) == 5
        out[mode] = (losses, [p.detach().clone() for p in disc.parameters()], gm.clone(), int(gen.batch_norm1.num_batches_tracked))
    (l0, p0, g0, n0), (l1, p1, g1, n1) = out["loop"], out["trainer"]
    assert n0 == n1 == 5 and (g0 - g1).abs().max().item() < 1e-5
    for a, b in zip(l0, l1):
        assert abs(a[0] - b[0]) <= 1e-3 * abs(a[0]) and abs(a[1] - b[1]) <= 1e-3 * abs(a[1]), (l0, l1)
    for a, b in zip(p0, p1):
        assert (a - b).abs().max().item() <= 5 * 2e-4 * 0.5 + 1e-7           # Adam: a sign flip of a ~0 gradient element moves a weight by <= lr per step
        assert _rel(a, b) < 2e-2
