"""Fused optimiser / loss replacements with the reference's constructor arguments.

``FusedAdam(params, lr, betas, eps)`` is a ``torch.optim.Optimizer`` (so ``StepLR`` and
``zero_grad`` work unchanged, network_tests.py:253-258,293,311,328-329) whose ``step`` is ONE
multi-tensor kernel launch per parameter group instead of a Python loop of ATen ops.  Parameters
whose ``grad is None`` are skipped and get no state, exactly like ``torch.optim.Adam``
(torch/optim/adam.py) -- which is what makes the reference's ``gen_opt.step()`` a no-op.
"""
import torch

from . import functional as Fn
from . import functional_tc as FnTC


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        if weight_decay != 0:
            raise NotImplementedError("the reference uses weight_decay=0")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        updated = False
        for group in self.param_groups:
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                by_step.setdefault(st["step"], []).append(p)
            b1, b2 = group["betas"]
            for step, ps in by_step.items():
                updated = True
                Fn.adam_step([p.data for p in ps], [p.grad.contiguous() for p in ps], [self.state[p]["exp_avg"] for p in ps],
                             [self.state[p]["exp_avg_sq"] for p in ps], step, group["lr"], b1, b2, group["eps"])
        if updated:
            FnTC.invalidate_weight_cache()       # the kernel wrote the parameters through raw pointers: packed bf16 copies are stale
        return loss


class BCEWithLogitsLoss(torch.nn.Module):
    """nn.BCEWithLogitsLoss() (mean) as one fused forward+backward kernel."""

    def forward(self, input, target):
        return Fn.bce_with_logits(input, target)
