"""A/B of the rasteriser's step kernel on BASELINE config 4: speculate-and-verify (mode 0) vs sequential chain (mode 1); per-kernel times via CUDA events
around the whole call, fall-back count, bit-equality of the two outputs."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import raster_oracle as ro
from gan_des_midi_music_gen_b200 import _native as N
from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
dt, meta, off = ro.synth_songs(1276, 15000, 300.0, seed=0)
d = [torch.from_numpy(a).cuda() for a in (dt, meta.view(np.int32), off)]
outs = {}
for mode in (1, 0):
    N.lib().mmg_raster_set_mode(mode)
    ws = torch.zeros(ds.raster_workspace_bytes(1276, len(dt)), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        out = ds.rasterize_events(*d, 300, 0, 300, torch.float32, path="sort", workspace=ws)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = ds.rasterize_events(*d, 300, 0, 300, torch.float32, path="sort", workspace=ws); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    outs[mode] = out
    print(f"mode {mode}: ms {min(ts):.4f} (median {sorted(ts)[3]:.4f}) fallbacks (10 calls) {int(ws[-8:].view(torch.int64).item())} checksum {float(out.double().sum())}", flush=True)
N.lib().mmg_raster_set_mode(0)
print("equal:", bool(torch.equal(outs[0], outs[1])))
