"""Drop-in mirror of the reference's ``GAN_DES/SIMNN.py`` model API
(/root/reference/GAN_DES/SIMNN.py:37-142): ``get_noise``, ``weights_init``, ``Generator``,
``Discriminator`` with the same constructor arguments and state-dict keys (``conv{1-4}.weight``,
``batch_norm{1-3}.*`` / ``conv{1,2}.*``, ``fc{1,2}.*``), every layer running on the sm_100a kernels
of libmmgan_b200.so.  ``SimNN`` / ``generate_song`` (dead placeholder code) and the
FluidSynth audio bridge are out of scope.
"""
import torch
from torch import nn

from .. import functional as Fn
from .. import functional_tc as FnTC
from .._native import require_cuda

__all__ = ["get_noise", "weights_init", "Generator", "Discriminator"]


def get_noise(n_samples, noise_dim, device="cpu"):
    """SIMNN.py:37-46 -- N(0,1) noise of shape (n_samples, noise_dim, 1, 1)."""
    return torch.randn(n_samples, noise_dim, 1, 1, device=device)


def weights_init(m):
    """SIMNN.py:49-59 -- conv / conv-transpose / BatchNorm2d weights N(0, 0.02), BN bias 0."""
    if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.BatchNorm2d)):
        nn.init.normal_(m.weight, mean=0.0, std=0.02)
    if isinstance(m, nn.BatchNorm2d):
        nn.init.constant_(m.bias, val=0)


class Generator(nn.Module):
    """SIMNN.py:62-112 -- (B,100,1,1) -> ConvT k4 s1 p0 -> BN -> ReLU -> ConvT k4 s2 p1 -> BN -> ReLU
    -> ConvT k4 s2 p1 -> BN -> ReLU -> ConvT k5 s1 p0 -> sigmoid -> (B,1,20,20); all bias-free."""

    def __init__(self, no_of_channels=1, noise_dim=100, gen_dim=32):
        super().__init__()
        ct = lambda i, o, k, s, p: nn.ConvTranspose2d(i, o, kernel_size=k, stride=s, padding=p, bias=False)
        self.conv1 = ct(noise_dim, gen_dim * 4, 4, 1, 0)
        self.conv2 = ct(gen_dim * 4, gen_dim * 2, 4, 2, 1)
        self.conv3 = ct(gen_dim * 2, gen_dim, 4, 2, 1)
        self.conv4 = ct(gen_dim, no_of_channels, 5, 1, 0)
        self.batch_norm1 = nn.BatchNorm2d(gen_dim * 4)
        self.batch_norm2 = nn.BatchNorm2d(gen_dim * 2)
        self.batch_norm3 = nn.BatchNorm2d(gen_dim)
        self._ops = Fn
        self._initialize_weights()

    def enable_tensor_cores(self, enabled=True):
        """Route the four transposed convolutions (forward, data and weight gradients) to the tcgen05 GEMM kernel (bf16 operands, fp32
        accumulation: functional_tc.py); BatchNorm / activations stay on the fp32 kernels.  Parameters and state-dict keys are untouched."""
        self._ops = FnTC if enabled else Fn
        return self

    def _initialize_weights(self):
        for m in self.modules():
            if isinstance(m, nn.ConvTranspose2d):
                nn.init.normal_(m.weight, 0.0, 0.02)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.normal_(m.weight, 1.0, 0.02)
                nn.init.constant_(m.bias, 0)

    def _bn_relu(self, bn, z):
        if self.training and bn.track_running_stats:
            bn.num_batches_tracked += 1
        return Fn.batch_norm(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, self.training,
                             bn.momentum if bn.momentum is not None else 0.1, bn.eps, Fn.ACT_RELU)

    def forward(self, input):
        require_cuda(input)
        ops = self._ops
        x = self._bn_relu(self.batch_norm1, ops.conv_transpose2d(input, self.conv1.weight, 1, 0))
        x = self._bn_relu(self.batch_norm2, ops.conv_transpose2d(x, self.conv2.weight, 2, 1))
        x = self._bn_relu(self.batch_norm3, ops.conv_transpose2d(x, self.conv3.weight, 2, 1))
        return ops.conv_transpose2d(x, self.conv4.weight, 1, 0, Fn.ACT_SIGMOID)


class Discriminator(nn.Module):
    """SIMNN.py:115-142 -- (B,128,216) spectrograms -> conv k2 p1 -> ReLU -> pool2 -> conv k3 p1 -> ReLU
    -> pool2 -> fc 55296->128 -> ReLU -> fc 128->1 -> sigmoid (the loop then applies BCEWithLogits on
    top of this sigmoid: a quirk of the reference that is preserved)."""

    def __init__(self, no_of_channels=1, disc_dim=32):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 16, kernel_size=2, stride=1, padding=1)
        self.conv2 = nn.Conv2d(16, 32, kernel_size=3, stride=1, padding=1)
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2, padding=0)
        self.fc1 = nn.Linear(32 * 32 * 54, 128)
        self.fc2 = nn.Linear(128, 1)
        self._ops = Fn

    def enable_tensor_cores(self, enabled=True):
        """Route conv1 / conv2 / fc1 / fc2 (forward, data and weight gradients) to the tcgen05 GEMM kernel (bf16 operands, fp32 accumulation:
        functional_tc.py); max-pooling and the loss stay on the fp32 kernels.  Parameters and state-dict keys are untouched."""
        self._ops = FnTC if enabled else Fn
        return self

    def forward(self, input):
        require_cuda(input)
        x = torch.unsqueeze(input, 1)
        ops = self._ops
        if ops is FnTC:       # conv + ReLU + pool as one autograd node: its backward undoes pooling and ReLU in one kernel
            x = FnTC.conv2d_relu_pool(x, self.conv1.weight, self.conv1.bias, 1)
            x = FnTC.conv2d_relu_pool(x, self.conv2.weight, self.conv2.bias, 1)
        else:
            x = Fn.max_pool2(Fn.conv2d(x, self.conv1.weight, self.conv1.bias, 1, 1, Fn.ACT_RELU))
            x = Fn.max_pool2(Fn.conv2d(x, self.conv2.weight, self.conv2.bias, 1, 1, Fn.ACT_RELU))
        x = x.view(-1, 32 * 32 * 54)
        x = ops.linear(x, self.fc1.weight, self.fc1.bias, Fn.ACT_RELU)
        return ops.linear(x, self.fc2.weight, self.fc2.bias, Fn.ACT_SIGMOID)
