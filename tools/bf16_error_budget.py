"""Where the bf16 path's gradient error comes from (CPU, torch only; evidence for DESIGN.md section 2 and the bounds in tests/test_gpu_disc_pass.py).

Emulates DiscriminatorCNN forward + BCE + backward in fp32 with bf16 ROUNDING POINTS switched on one at a time:
    w1 / w2   conv weights rounded to bf16 (the MMA B operands)
    a1 / a2   conv activations rounded to bf16 (the MMA A operands / what the backward reads)
    dz        the gradient tensors DZ2 / DZ1 rounded to bf16 (the wgrad / dgrad MMA operands)
and prints the rel-L2 distance of every parameter gradient to the all-fp32 result.

Finding (synthetic state of the golden files, B = 256): rounding w1, w2 or a1 ALONE costs 1.4-1.8e-2 on conv1.weight (LeakyReLU masks of
near-zero pre-activations flip, which changes a gradient element by a factor 5); rounding dz adds nothing measurable (2.38e-2 -> 2.39e-2).
A hi + lo split of DZ2 / DZ1 would therefore buy nothing; only an fp32-grade FORWARD would.  With the discriminator of the reference's
shipped checkpoint (the conditions of SURVEY 8d's probe; needs /root/reference) all six tensors stay below 1e-2 (worst conv2.weight 6.4e-3).

    python tools/bf16_error_budget.py [B]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.nn.functional as F

import mmgan_oracle as mo


def rb(x):
    return x + (x.bfloat16().float() - x).detach()


class RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def grads(P0, x, y, w1=False, w2=False, a1=False, a2=False, dz=False):
    P = {k: v.clone().float().requires_grad_(True) for k, v in P0.items()}
    I = lambda t: t
    rg = RoundGrad.apply if dz else I
    z1 = F.conv2d(x, (rb if w1 else I)(P["conv1.weight"]), P["conv1.bias"], stride=2, padding=1)
    t1 = (rb if a1 else I)(F.leaky_relu(rg(z1), 0.2))
    z2 = F.conv2d(t1, (rb if w2 else I)(P["conv2.weight"]), P["conv2.bias"], stride=2, padding=1)
    t2 = (rb if a2 else I)(F.leaky_relu(rg(z2), 0.2))
    logit = (t2.reshape(len(x), -1) @ P["fc.weight"].t() + P["fc.bias"]).squeeze(1)
    loss = F.binary_cross_entropy_with_logits(logit, torch.full((len(x),), y))
    return dict(zip(P.keys(), torch.autograd.grad(loss, list(P.values()))))


def table(P0, x, y, title):
    ref = grads(P0, x, y)
    print(f"--- {title}, target {y}")
    rows = [("w1", dict(w1=True)), ("w2", dict(w2=True)), ("a1", dict(a1=True)), ("a2", dict(a2=True)), ("dz", dict(dz=True)),
            ("w1+w2+a1+a2", dict(w1=True, w2=True, a1=True, a2=True)), ("all (the kernel)", dict(w1=True, w2=True, a1=True, a2=True, dz=True))]
    for name, kw in rows:
        got = grads(P0, x, y, **kw)
        print(f"{name:18s}", {k: f"{((got[k] - ref[k]).norm() / ref[k].norm()).item():.2e}" for k in ref})


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    torch.set_num_threads(os.cpu_count() or 1)
    sd = mo.synth_state(mo.mmgan_shapes(), seed=7, d_scale=0.25)
    P0 = {k[len("discriminator."):]: v for k, v in sd.items() if k.startswith("discriminator.")}
    x = torch.from_numpy(mo.synth_rolls(B, 50, seed=8)).float()
    table(P0, x, 0.0, f"synthetic state (mmgan_oracle.synth_state seed 7), B = {B}")
    ck = os.path.join(os.environ.get("MMG_REFERENCE_ROOT", "/root/reference"), "MMGAN_MIDI_DES", "models", "mmgan_64_64_epoch_1.pth")
    if os.path.exists(ck):
        sd = torch.load(ck, map_location="cpu")
        P0 = {k[len("discriminator."):]: v for k, v in sd.items() if k.startswith("discriminator.")}
        for y in (0.0, 1.0):
            table(P0, x, y, f"shipped checkpoint mmgan_64_64_epoch_1.pth (SURVEY 8d probe), B = {B}")
