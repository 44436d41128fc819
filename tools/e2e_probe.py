"""Where does the e2e step time go: H2D alone, compute alone, pipelined."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
from gan_des_midi_music_gen_b200.trainer import MMGANTrainer, HostBatchPipeline
from gan_des_midi_music_gen_b200.benchmark import _synth_rolls_u8
B = 8192
dev = torch.device("cuda", 0)
m = nt.MultiModalGAN(z_dim=50, adj_size=(64, 64), roll_size=(2, 128, 50), input_dim=50, output_dim=20, instrument=0, start=100, end=150, device=dev).train()
tr = MMGANTrainer(m, lr=0.01, precision="bf16", max_batch=B, inner_rng="device")
h = {k: _synth_rolls_u8(B, 50, i, "cpu").pin_memory() for i, k in enumerate(("real", "fake_d", "fake_g"))}
h["beats"] = (25.0 * torch.rand(B, 50)).pin_memory()
pipe = HostBatchPipeline(tr, h)
def t(fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
st = pipe.stage[0]
def copy_only():
    with torch.cuda.stream(pipe.copy_stream):
        for k in pipe.KEYS: st[k].copy_(h[k], non_blocking=True)
    pipe.copy_stream.synchronize()
n = [torch.randn(B, 50, device=dev) for _ in range(2)]
def compute_only():
    tr.step(n[0], n[1], st["beats"], st["real"], st["fake_d"], st["fake_g"])
for _ in range(3): copy_only(); compute_only()
print("copy only ms", t(copy_only))
print("compute only ms", t(compute_only))
def both():
    with torch.cuda.stream(pipe.copy_stream):
        for k in pipe.KEYS: pipe.stage[1][k].copy_(h[k], non_blocking=True)
    compute_only()
print("copy || compute ms", t(both))
t0 = time.perf_counter(); compute_only(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host launch ms", (t1 - t0) * 1e3, "until done ms", (t2 - t0) * 1e3)
def piped():
    for _ in pipe.run([h] * 5): pass
piped(); piped()
print("pipeline ms/step", t(piped, 2) / 5)
