"""Times GenTC forward (train mode) for both generators at the bench batch size, replayed from a CUDA graph (what the trainer does), for
2 and 4 builder / epilogue warp groups per CTA (mmg_gen_set_worker_groups); checks that both variants give the same output."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import __graft_entry__  # noqa: E402,F401  (registers the package alias)
from gan_des_midi_music_gen_b200 import _native as N
from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
from gan_des_midi_music_gen_b200.gen_tc import GenTC
_pos = [x for x in sys.argv[1:] if not x.startswith("--")]
B = int(_pos[0]) if _pos else 16384
m = nt.MultiModalGAN(z_dim=50, adj_size=(64, 64), roll_size=(2, 128, 50), input_dim=50, output_dim=20, instrument=0, start=100, end=150, device="cuda").train()
g1, g2 = GenTC(m.generator1, B), GenTC(m.generator2, B)
n = [torch.randn(B, 50, device="cuda") for _ in range(4)]
o1, o2 = torch.empty(B, 4096, device="cuda"), torch.empty(B, 20, device="cuda")


def graph_time(fn, reps=10, rounds=5):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(rounds):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


g1u, g2u = GenTC(m.generator1, B, fused_hidden=False), GenTC(m.generator2, B, fused_hidden=False)
for name, fn in (("G1 per-layer launches", lambda: g1u.forward(n[0], n[1], out=o1)), ("G1 fused hidden blocks", lambda: g1.forward(n[0], n[1], out=o1)),
                 ("G2 per-layer launches", lambda: g2u.forward(n[2], n[3], out=o2)), ("G2 fused hidden blocks", lambda: g2.forward(n[2], n[3], out=o2))):
    print(f"{name}: {graph_time(fn):.1f} us", flush=True)
side = torch.cuda.Stream()
def both():          # what the trainer does: the beat generator on a side stream (fork / join, also under capture)
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        g2.forward(n[2], n[3], out=o2)
    g1.forward(n[0], n[1], out=o1)
    cur.wait_stream(side)
for wg in (2, 4):
    N.lib().mmg_gen_set_worker_groups(wg)
    print(f"G1 on the main stream + G2 on a side stream, {wg} worker groups: {graph_time(both):.1f} us", flush=True)
outs = {}
for wg in (2, 4):
    N.lib().mmg_gen_set_worker_groups(wg)
    for name, fn, o in (("G1", lambda: g1.forward(n[0], n[1], out=o1), o1), ("G2", lambda: g2.forward(n[2], n[3], out=o2), o2),
                        ("G1 eval", lambda: g1.forward(n[0], n[1], training=False, out=o1), o1)):
        t = graph_time(fn)
        outs.setdefault(name, []).append(o.clone())
        print(f"worker groups {wg}: {name} forward {t:.1f} us (B = {B})", flush=True)
for name, v in outs.items():
    d = max((v[0] - x).abs().max().item() for x in v[1:])
    print(name, "max |difference| between variants / repeats:", d, "finite:", bool(torch.isfinite(v[-1]).all()))

# ---- per-call breakdown of one train-mode G1 forward: every C-ABI call of the forward replayed alone (10 x in a graph, warm caches)
if "--breakdown" in sys.argv:
    N.lib().mmg_gen_set_worker_groups(4)
    calls, orig = [], N.call
    def rec(name, *a):
        calls.append((name, a))
        return orig(name, *a)
    N.call = rec
    g1.forward(n[0], n[1], out=o1)
    N.call = orig
    torch.cuda.synchronize()
    tot = 0.0
    for name, a in calls:
        t = graph_time(lambda: orig(name, *a[:-1], N.stream()))
        tot += t
        extra = ""
        if name == "mmg_gen_layer_fwd":
            st = a[0]._obj
            extra = f"K = {st.k0 + st.k1}, N = {st.N}, z_out = {bool(st.z_out)}, out_sums = {bool(st.out_sums)}, y_out = {bool(st.y_out)}"
        print(f"   {name:28s} {t:7.1f} us  {extra}")
    print(f"   sum of the parts {tot:.1f} us")
    raw = N.lib()
    if hasattr(raw, "mmg_gen_get_stamps"):                 # -DMMG_ABLATION builds only: clock stamps of CTA 0 in the output layer's kernel
        import ctypes
        name, a = calls[-1]
        orig(name, *a[:-1], N.stream())
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * (3 * 32 * 4))()
        raw.mmg_gen_get_stamps(buf)
        st = torch.tensor(list(buf)).view(3, 32, 4)
        t0 = int(st[0, 0, 0])
        print("   stamps of CTA 0 (cycles since the producer's first wait)")
        for it in range(13):
            print(f"     item {it:2d}  producer wait {int(st[0, it, 0]) - t0:7d} -> {int(st[0, it, 1]) - t0:7d} | MMA tempty {int(st[1, it, 0]) - t0:7d} aready {int(st[1, it, 1]) - t0:7d} wfull {int(st[1, it, 2]) - t0:7d} "
                  f"committed {int(st[1, it, 3]) - t0:7d} | worker at tfull wait {int(st[2, it, 0]) - t0:7d} tfull {int(st[2, it, 1]) - t0:7d} epilogue done {int(st[2, it, 2]) - t0:7d}")
