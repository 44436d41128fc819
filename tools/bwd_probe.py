"""Times DiscTC forward pieces and the fused backward alone at the bench batch size."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
from gan_des_midi_music_gen_b200.disc_tc import DiscTC
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D = nt.DiscriminatorCNN(roll_size=(2, 128, 50)).cuda()
tc = DiscTC(D, B)
x = ((torch.rand(B, 2, 128, 50, device="cuda") < 0.02) * 77).to(torch.uint8)
dl = torch.randn(B, device="cuda") / B
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
print("skip", os.environ.get("MMG_DBG_SKIP", "0"), "forward us", round(t(lambda: tc.forward(x)), 1), "backward us", round(t(lambda: tc.backward(dl)), 1))
