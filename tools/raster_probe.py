"""Times the rasteriser on BASELINE config 4 (used under ncu for the per-kernel split)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import raster_oracle as ro
from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1276
dt, meta, off = ro.synth_songs(S, 15000, 300.0, seed=0)
d = [torch.from_numpy(a).cuda() for a in (dt, meta.view(np.int32), off)]
for dtype in (torch.float32, torch.uint8):
    for _ in range(3):
        out = ds.rasterize_events(*d, 300, 0, 300, dtype)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = ds.rasterize_events(*d, 300, 0, 300, dtype); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(dtype, "ms", min(ts), "checksum", float(out.double().sum()))
