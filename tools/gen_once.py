"""Three train-mode G1 forwards at the bench batch (eager launches): the target of the ncu --set full capture of the generator kernels."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import __graft_entry__  # noqa: E402,F401
from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
from gan_des_midi_music_gen_b200.gen_tc import GenTC
B = 16384
m = nt.MultiModalGAN(z_dim=50, adj_size=(64, 64), roll_size=(2, 128, 50), input_dim=50, output_dim=20, instrument=0, start=100, end=150, device="cuda").train()
g1 = GenTC(m.generator1, B)
n0, n1, o1 = torch.randn(B, 50, device="cuda"), torch.randn(B, 50, device="cuda"), torch.empty(B, 4096, device="cuda")
for _ in range(3):
    g1.forward(n0, n1, out=o1)
torch.cuda.synchronize()
print("ok", float(o1.mean()))
