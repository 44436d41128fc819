"""Times GenTC forward (train mode) for both generators at the bench batch size."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
from gan_des_midi_music_gen_b200.gen_tc import GenTC
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = nt.MultiModalGAN(z_dim=50, adj_size=(64, 64), roll_size=(2, 128, 50), input_dim=50, output_dim=20, instrument=0, start=100, end=150, device="cuda").train()
g1, g2 = GenTC(m.generator1, B), GenTC(m.generator2, B)
n = [torch.randn(B, 50, device="cuda") for _ in range(4)]
o1, o2 = torch.empty(B, 4096, device="cuda"), torch.empty(B, 20, device="cuda")
for name, fn in (("G1", lambda: g1.forward(n[0], n[1], out=o1)), ("G2", lambda: g2.forward(n[2], n[3], out=o2))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, "forward us", e0.elapsed_time(e1) * 1e3 / 20)
