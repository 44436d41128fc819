"""One rasteriser call on BASELINE config 4 (for ncu captures): python tools/raster_one.py <path> [songs]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import raster_oracle as ro
from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
path = sys.argv[1] if len(sys.argv) > 1 else "stream"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1276
dt, meta, off = ro.synth_songs(S, 15000, 300.0, seed=0)
d = [torch.from_numpy(a).cuda() for a in (dt, meta.view(np.int32), off)]
for _ in range(3):
    out = ds.rasterize_events(*d, 300, 0, 300, torch.float32, path=path)
torch.cuda.synchronize()
print("checksum", float(out.double().sum()))
