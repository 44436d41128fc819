"""Drop-in mirror of the reference's ``MMGAN_MIDI_DES/network_tests.py`` model API
(/root/reference/MMGAN_MIDI_DES/network_tests.py:43-206), backed by the sm_100a kernels of
libmmgan_b200.so.  Same class names, constructor arguments, attribute names and state-dict keys
(``generator{1,2}.gen.{0-3}.{0,1}.*``, ``discriminator.{conv1,conv2,fc}.*``), so the reference's
shipped ``.pth`` checkpoints load unchanged and stock ``torch.optim`` / ``nn.BCEWithLogitsLoss`` /
``DataLoader`` keep working.  Parameters stay ordinary fp32 ``nn.Parameter``s; every forward and
backward launches this repo's CUDA kernels (no ATen math, no CPU fallback).

Differences that are deliberate:
  * the host DES bridge (``matrix_to_midi``) is an injected callable (``MultiModalGAN.bridge``): the
    reference's unchanged ``matrix_sim_process.matrix_to_midi`` plugs in there; without one, forward
    raises, because the simulator is outside this repo's scope.
  * ``make_dot_png`` is accepted and ignored (torchviz debug dump, no numerical effect).
"""
import math

import torch
from torch import nn

from .. import functional as Fn
from .._native import require_cuda

__all__ = ["get_noise", "weights_init", "Generator", "BeatGenerator", "Discriminator", "DiscriminatorCNN", "MultiModalGAN"]


def get_noise(n_samples, noise_dim, device="cpu"):
    """network_tests.py:43-44 -- N(0,1) noise of shape (n_samples, noise_dim)."""
    return torch.randn(n_samples, noise_dim, device=device)


def weights_init(m):
    """network_tests.py:47-55 -- conv weights N(0,1); BatchNorm2d and Linear xavier-normal, bias 0.
    (BatchNorm1d is not matched, exactly as in the reference, so it keeps gamma=1, beta=0.)"""
    if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
        nn.init.normal_(m.weight, mean=0, std=1)
    if isinstance(m, (nn.BatchNorm2d, nn.Linear)):
        nn.init.xavier_normal_(m.weight)
        nn.init.constant_(m.bias, 0.0)


class _GenBlock(nn.Sequential):
    """[Linear, BatchNorm1d, Sigmoid] holder (children '0','1','2' as in the reference, so the
    state-dict keys match); forward is two fused launches: GEMM(+bias) and BN(+sigmoid)."""

    def __init__(self, d_in, d_out):
        super().__init__(nn.Linear(d_in, d_out), nn.BatchNorm1d(d_out), nn.Sigmoid())

    def forward(self, x):
        lin, bn = self[0], self[1]
        z = Fn.linear(x, lin.weight, lin.bias)
        use_batch_stats = self.training or bn.running_mean is None
        if self.training and bn.track_running_stats:
            bn.num_batches_tracked += 1
        return Fn.batch_norm(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, use_batch_stats,
                             bn.momentum if bn.momentum is not None else 0.1, bn.eps, Fn.ACT_SIGMOID)


class _GenBase(nn.Module):
    def _build(self, z_dim, hidden_dim, input_dim, out_features, device):
        self.z_dim = z_dim
        self.device = device
        self.input_tensor_dim = z_dim if input_dim is None else input_dim
        widths = [z_dim + self.input_tensor_dim, hidden_dim * 4, hidden_dim * 2, hidden_dim, out_features]
        self.gen = nn.Sequential(*[self.make_gen_block(a, b) for a, b in zip(widths[:-1], widths[1:])])
        self.gen.apply(weights_init)

    def make_gen_block(self, input_dim, output_dim):
        return _GenBlock(input_dim, output_dim)

    def enable_tensor_cores(self, max_batch, enabled=True):
        """Route ``forward`` through the bf16 tcgen05 generator blocks (csrc/gen_tc.cu) whenever autograd is off (``torch.no_grad()`` / inference:
        the reference never back-propagates into the generators; with autograd on the fp32 differentiable kernels are used)."""
        if enabled:
            from ..gen_tc import GenTC
            self._tc = GenTC(self, int(max_batch))
        else:
            self._tc = None
        return self

    def _features(self, noise, input_tensor):
        tc = getattr(self, "_tc", None)
        if tc is not None and not torch.is_grad_enabled() and len(noise) <= tc.cap:
            if input_tensor is None:
                input_tensor = torch.randn(len(noise), self.input_tensor_dim).to(self.device)      # same RNG consumption as below
            require_cuda(noise, input_tensor)
            return tc.forward(noise, input_tensor)
        if input_tensor is None:
            # drawn on the default (CPU) generator and then moved, like the reference (:83-84,:119-120),
            # so the global RNG stream is consumed identically
            input_tensor = torch.randn(len(noise), self.input_tensor_dim).to(self.device)
        x = torch.cat((noise, input_tensor), dim=1).to(self.device)
        require_cuda(x)
        return self.gen(x)


class Generator(_GenBase):
    """network_tests.py:58-90 -- noise (B,z) [+ input (B,input_dim) or fresh randn] -> (B, im_chan, *adj_size) in (0,1)."""

    def __init__(self, z_dim=10, im_chan=1, hidden_dim=64, input_dim=None, adj_size=None, device="cpu"):
        super().__init__()
        self.adj_size = adj_size
        self._build(z_dim, hidden_dim, input_dim, im_chan * adj_size[0] * adj_size[1], device)

    def forward(self, noise, input_tensor=None):
        out = self._features(noise, input_tensor)
        return out.view(len(noise), -1, self.adj_size[0], self.adj_size[1]).to(self.device)


class BeatGenerator(_GenBase):
    """network_tests.py:93-123 -- same stack, output (B, output_dim) simulator parameters."""

    def __init__(self, z_dim=10, hidden_dim=64, input_dim=None, output_dim=None, device="cpu"):
        super().__init__()
        self.output_dim = output_dim
        self._build(z_dim, hidden_dim, input_dim, output_dim, device)

    def forward(self, noise, input_tensor=None):
        return self._features(noise, input_tensor)


class _DiscBlock(nn.Sequential):
    def __init__(self, d_in, d_out):
        super().__init__(nn.Linear(d_in, d_out), nn.LeakyReLU(0.2, inplace=True))

    def forward(self, x):
        return Fn.linear(x, self[0].weight, self[0].bias, Fn.ACT_LRELU)


class Discriminator(nn.Module):
    """network_tests.py:126-144 -- 3 x [Linear, LeakyReLU(0.2)] MLP (unused by MultiModalGAN; API completeness)."""

    def __init__(self, im_chan=1, hidden_dim=16, roll_size=None, device="cpu"):
        super().__init__()
        self.roll_size = roll_size
        self.device = device
        flat = im_chan * roll_size[0] * roll_size[1] * roll_size[2]
        self.disc = nn.Sequential(self.make_disc_block(flat, hidden_dim), self.make_disc_block(hidden_dim, hidden_dim * 2),
                                  self.make_disc_block(hidden_dim * 2, 1))

    def make_disc_block(self, input_dim, output_dim):
        return _DiscBlock(input_dim, output_dim)

    def forward(self, image):
        require_cuda(image)
        return self.disc(image)


class DiscriminatorCNN(nn.Module):
    """network_tests.py:147-160 -- Conv2d(k4,s2,p1)+LeakyReLU x2 -> flatten -> Linear -> (B,1) logits."""

    def __init__(self, roll_size=(2, 128, 30), hidden_dim=16):
        super().__init__()
        self.conv1 = nn.Conv2d(roll_size[0], hidden_dim, kernel_size=4, stride=2, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, hidden_dim * 2, kernel_size=4, stride=2, padding=1)
        self.leaky_relu = nn.LeakyReLU(0.2, inplace=True)
        self.final_size = hidden_dim * 2 * ((roll_size[1] // 4) * (roll_size[2] // 4))
        self.fc = nn.Linear(self.final_size, 1)

    def enable_tensor_cores(self, max_batch, enabled=True):
        """Route ``forward`` (and its autograd backward) through the bf16 tcgen05 kernels (csrc/disc_tc_fused.cu) for inputs of shape
        (B <= max_batch, 2, 128, 50), float32 or uint8: bf16 operands, fp32 accumulation, fp32 master weights (tolerances of SURVEY 8d).  Off by
        default: the module then computes in fp32 like the reference.  The training loop proper should use ``trainer.MMGANTrainer`` (one kernel per
        D pass); this switch is for code that keeps calling the module (the reference loop as written, demo notebooks)."""
        if enabled:
            from ..disc_tc import DiscTC
            self._tc = DiscTC(self, int(max_batch))
        else:
            self._tc = None
        return self

    def forward(self, image):
        require_cuda(image)
        tc = getattr(self, "_tc", None)
        if tc is not None and image.dim() == 4 and tuple(image.shape[1:]) == (2, 128, 50) and image.shape[0] <= tc.cap \
                and image.dtype in (torch.float32, torch.uint8):
            from ..disc_tc import DiscTCFunction
            return DiscTCFunction.apply(image, tc, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias, self.fc.weight, self.fc.bias)
        x = Fn.conv2d(image, self.conv1.weight, self.conv1.bias, 2, 1, Fn.ACT_LRELU)
        x = Fn.conv2d(x, self.conv2.weight, self.conv2.bias, 2, 1, Fn.ACT_LRELU)
        return Fn.linear(x.view(len(x), -1), self.fc.weight, self.fc.bias)


class MultiModalGAN(nn.Module):
    """network_tests.py:163-206 -- container of generator1 / generator2 / discriminator.

    ``forward`` = G1, G2 -> host bridge (DES -> MIDI -> piano roll) -> D, returning
    ``(logits (B,1), failed_sim_count)``.  ``bridge`` is any callable with the signature of the
    reference's ``matrix_to_midi(gen1_output, gen2_output, adj_size=, instrument=, start=, end=, count=,
    generate=)`` returning ``(list of (2,128,W) float64 arrays, failed_count)``."""

    def __init__(self, z_dim=100, hidden_dim=64, adj_size=(28, 28), roll_size=(2, 128, 50), input_dim=50, output_dim=16,
                 instrument=None, start=30, end=80, device="cpu", bridge=None):
        super().__init__()
        self.z_dim = z_dim
        self.generator1 = Generator(z_dim, hidden_dim=hidden_dim, adj_size=adj_size, device=device).to(device)
        self.generator2 = BeatGenerator(z_dim, hidden_dim=hidden_dim, input_dim=input_dim, output_dim=output_dim, device=device).to(device)
        self.discriminator = DiscriminatorCNN(roll_size=roll_size).to(device)
        self.instrument = instrument
        self.start = start
        self.end = end
        self.adj_size = adj_size
        self.device = device
        self.bridge = bridge

    def _simulate(self, g1, g2, **kw):
        if self.bridge is None:
            raise RuntimeError("MultiModalGAN.bridge is not set: plug in the reference's matrix_to_midi (host DES); "
                               "the simulator is outside this repo's scope")
        rolls, failed = self.bridge(g1.detach(), g2.detach(), adj_size=self.adj_size, instrument=self.instrument, start=self.start,
                                    end=self.end, **kw)
        return rolls, failed

    def forward(self, noise1, noise2, input_tensor, count, make_dot_png=True):
        gen_output1 = self.generator1(noise1)
        gen_output2 = self.generator2(noise2, input_tensor)
        rolls, failed_sim_count = self._simulate(gen_output1, gen_output2, count=count)
        if torch.is_tensor(rolls):
            sim_output = rolls.to(self.device, torch.float32)
        else:       # list of numpy (2,128,W) float64: one pinned staging copy instead of B small ones
            import numpy as np
            host = torch.from_numpy(np.stack(rolls)).float()
            sim_output = host.pin_memory().to(self.device, non_blocking=True) if torch.cuda.is_available() else host.to(self.device)
        return self.discriminator(sim_output), failed_sim_count

    def generate_midi(self, noise1, noise2, input_tensor):
        self.generator1.eval()
        self.generator2.eval()
        gen_output1 = self.generator1(noise1)
        gen_output2 = self.generator2(noise2, input_tensor)
        sim_output, _ = self._simulate(gen_output1, gen_output2, generate=True)
        return sim_output
