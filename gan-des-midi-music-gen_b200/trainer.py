"""Fused MM-GAN training iteration (the reference loop body, network_tests.py:292-315) behind one call.

``MMGANTrainer.step`` performs, for one batch:
    D step: G1/G2 train-mode forward (BN running stats update), D forward on the fake and the real
            rolls, BCE (fake vs 0, real vs 1), backward, Adam on the 6 discriminator tensors;
    G step: G1/G2 forward again, D forward on the G-step fake rolls, BCE vs 1, backward (its D grads
            accumulate on top of the D-step grads, as in the reference, and are discarded by the next
            step), generator Adam is a no-op because no generator parameter has a grad.
The fake rolls are what the host DES bridge returned for the two generator forwards; here they are
inputs.  Rolls may be float32 or uint8 (B,2,128,W) tensors (piano-roll values are integers 0..127 and
durations < 256, so uint8 is exact and quarters the H2D / HBM traffic).

precision='fp32' : the drop-in nn.Modules + autograd over the fp32 kernels (reference tolerance).
Data parallel: with torch.distributed initialised (NCCL), the batch is sharded by rank, the D
gradients are averaged with one all-reduce per optimiser step (84 KB) and the G-step D grads are not
reduced (the reference discards them).
"""
import torch

from . import functional as Fn
from .optim import FusedAdam


class MMGANTrainer:
    def __init__(self, mmgan, lr=0.01, betas=(0.9, 0.999), eps=1e-8, precision="fp32", process_group=None, sync_bn=False):
        if precision != "fp32":
            raise NotImplementedError("bf16 tensor-core path: see trainer_bf16 (not wired yet)")
        self.m = mmgan
        self.precision = precision
        self.disc_opt = FusedAdam(mmgan.discriminator.parameters(), lr=lr, betas=betas, eps=eps)
        self.gen_opt = FusedAdam(list(mmgan.generator1.parameters()) + list(mmgan.generator2.parameters()), lr=lr, betas=betas, eps=eps)
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        self._flat = None

    def _allreduce_d_grads(self):
        if self.world == 1:
            return
        ps = [p for p in self.m.discriminator.parameters()]
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
        torch.distributed.all_reduce(flat, group=self.pg)
        flat.div_(self.world)
        o = 0
        for p in ps:
            n = p.numel()
            p.grad.copy_(flat[o:o + n].view_as(p))
            o += n

    def step(self, noise1, noise2, beats, real, fake_d, fake_g, inner_d=None, inner_g=None):
        m, D = self.m, self.m.discriminator
        B = len(noise1)
        f = lambda t: t if t.dtype == torch.float32 else t.float()
        # ---- D step (:293-308)
        self.disc_opt.zero_grad(set_to_none=True)
        with torch.no_grad():
            self.g1_out = m.generator1(noise1, inner_d)
            self.g2_out = m.generator2(noise2, beats)
        lf = Fn.bce_with_logits(D(f(fake_d)).squeeze(-1), 0.0)
        lr_ = Fn.bce_with_logits(D(f(real)).squeeze(-1), 1.0)
        disc_loss = lf + lr_
        disc_loss.backward()
        self._allreduce_d_grads()
        self.disc_opt.step()
        # ---- G step (:311-315)
        self.gen_opt.zero_grad(set_to_none=True)
        with torch.no_grad():
            self.g1_out = m.generator1(noise1, inner_g)
            self.g2_out = m.generator2(noise2, beats)
        gen_loss = Fn.bce_with_logits(D(f(fake_g)).squeeze(-1), 1.0)
        gen_loss.backward()
        self.gen_opt.step()          # no-op: generator grads are None
        return disc_loss.detach(), gen_loss.detach()
