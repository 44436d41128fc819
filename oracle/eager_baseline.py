"""ORACLE-SIDE BASELINE (test / measurement infrastructure, not product code): the reference's MM-GAN modules and loop body restated with
stock torch.nn layers, to be run under PyTorch EAGER on a CUDA device -- the "library GPU" bar of SURVEY 2.2 / 8d (cuBLAS / cuDNN / ATen
kernels on the same B200).  Class bodies follow /root/reference/MMGAN_MIDI_DES/network_tests.py:58-160 (Generator, BeatGenerator,
DiscriminatorCNN) and the loop body :281-321 with matrix_to_midi replaced by given fake rolls (the host DES is excluded on every arm).
Only bench.py --impl eager imports this."""
import torch
from torch import nn


def _gen_block(i, o):
    return nn.Sequential(nn.Linear(i, o), nn.BatchNorm1d(o), nn.Sigmoid())            # network_tests.py:75-80


class Generator(nn.Module):                                                           # :58-90
    def __init__(self, z_dim=50, hidden_dim=64, input_dim=50, out_features=4096, view=None):
        super().__init__()
        self.input_dim, self.view = input_dim, view
        self.gen = nn.Sequential(_gen_block(z_dim + input_dim, hidden_dim * 4), _gen_block(hidden_dim * 4, hidden_dim * 2),
                                 _gen_block(hidden_dim * 2, hidden_dim), _gen_block(hidden_dim, out_features))

    def forward(self, noise, input_tensor=None):
        if input_tensor is None:
            input_tensor = torch.randn(len(noise), self.input_dim, device=noise.device)
        out = self.gen(torch.cat((noise, input_tensor), dim=1))
        return out.view(len(noise), -1, *self.view) if self.view else out


class DiscriminatorCNN(nn.Module):                                                    # :147-160
    def __init__(self, roll_size=(2, 128, 50), hidden_dim=16):
        super().__init__()
        self.conv1 = nn.Conv2d(roll_size[0], hidden_dim, kernel_size=4, stride=2, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, hidden_dim * 2, kernel_size=4, stride=2, padding=1)
        self.leaky_relu = nn.LeakyReLU(0.2, inplace=True)
        self.fc = nn.Linear(hidden_dim * 2 * ((roll_size[1] // 4) * (roll_size[2] // 4)), 1)

    def forward(self, image):
        x = self.leaky_relu(self.conv1(image))
        x = self.leaky_relu(self.conv2(x))
        return self.fc(x.reshape(len(x), -1))


class EagerMMGAN:
    """The loop body (:292-315) on one device; ``autocast`` = torch.autocast(bfloat16) around the module calls (a bf16 library bar)."""

    def __init__(self, device, autocast=False, channels_last=False):
        self.dev, self.autocast = device, autocast
        self.g1 = Generator(50, 64, 50, 4096, view=(64, 64)).to(device)
        self.g2 = Generator(50, 64, 50, 20).to(device)
        self.d = DiscriminatorCNN().to(device)
        if channels_last:
            self.d = self.d.to(memory_format=torch.channels_last)
        self.crit = nn.BCEWithLogitsLoss()                                              # :248
        self.gen_opt = torch.optim.Adam(list(self.g1.parameters()) + list(self.g2.parameters()), lr=0.01)   # :253
        self.disc_opt = torch.optim.Adam(self.d.parameters(), lr=0.01)                  # :254

    def iteration(self, noise1, noise2, beats, real, fake_d, fake_g):
        B = len(noise1)
        ones, zeros = torch.ones(B, device=self.dev), torch.zeros(B, device=self.dev)
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if self.autocast else torch.autocast("cuda", enabled=False)
        with ctx:
            self.disc_opt.zero_grad()                                                   # :293
            with torch.no_grad():
                self.g1(noise1); self.g2(noise2, beats)                                 # :294 -> :177-178 (graph cut at the host DES)
            loss = self.crit(self.d(fake_d).squeeze(), zeros) + self.crit(self.d(real).squeeze(), ones)   # :304-306
            loss.backward(); self.disc_opt.step()                                       # :307-308
            self.gen_opt.zero_grad()                                                    # :311
            with torch.no_grad():
                self.g1(noise1); self.g2(noise2, beats)                                 # :312
            gl = self.crit(self.d(fake_g).squeeze(), ones)                              # :313
            gl.backward(); self.gen_opt.step()                                          # :314-315 (generator grads are None: no-op)
        return loss, gl
