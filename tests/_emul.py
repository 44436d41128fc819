"""Test helper: DiscriminatorCNN forward + BCE + backward in torch fp32 with the bf16 path's ROUNDING POINTS (bf16 conv weights, bf16 conv
activations, bf16 gradient tensors DZ2 / DZ1; fp32 accumulation, fp32 biases / fc), i.e. what ideal bf16-operand arithmetic computes.  The
CUDA kernels are compared with it tightly; its own distance to the all-fp32 reference is the price of bf16 operands
(tools/bf16_error_budget.py), which no kernel can undercut."""
import torch
import torch.nn.functional as F

NAMES = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "fc.weight", "fc.bias"]


def rb(x):
    return x + (x.bfloat16().float() - x).detach()


class _RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def disc_pass(D, x8, target, n_rows, round_ops=True):
    """loss (float), logits (B,), grads {name: tensor} of one pass over the rolls x8 (uint8 or float), mean over n_rows rows."""
    ps = [p.detach().clone().requires_grad_(True) for p in (D.conv1.weight, D.conv1.bias, D.conv2.weight, D.conv2.bias, D.fc.weight, D.fc.bias)]
    w1, b1, w2, b2, wf, bf = ps
    r = rb if round_ops else (lambda t: t)
    rg = _RoundGrad.apply if round_ops else (lambda t: t)
    grads = [torch.zeros_like(p) for p in ps]
    logits, loss = [], 0.0
    for i in range(0, x8.shape[0], 1024):                          # chunks keep the fp32 activations small
        x = x8[i:i + 1024].float()
        a1 = r(F.leaky_relu(rg(F.conv2d(x, r(w1), b1, stride=2, padding=1)), 0.2))
        a2 = r(F.leaky_relu(rg(F.conv2d(a1, r(w2), b2, stride=2, padding=1)), 0.2))
        lg = (a2.reshape(x.shape[0], -1) @ wf.t() + bf).squeeze(1)
        ls = F.binary_cross_entropy_with_logits(lg, torch.full_like(lg, target), reduction="sum") / n_rows
        for g, gi in zip(grads, torch.autograd.grad(ls, ps)):
            g += gi
        logits.append(lg.detach())
        loss += ls.item()
    return loss, torch.cat(logits), dict(zip(NAMES, grads))
