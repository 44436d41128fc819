"""One rasteriser call on the training loop's call shape (for ncu captures): 32768 songs x 320 messages, W = 50, uint8."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import raster_oracle as ro
from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
dt, meta, off = ro.synth_songs(32768, 320, 60.0, seed=0, p_on=0.4, p_off=0.4)
d = [torch.from_numpy(a).cuda() for a in (dt, meta.view(np.int32), off)]
for _ in range(3):
    out = ds.rasterize_events(*d, 100, 0, 50, torch.uint8, path="stream")
torch.cuda.synchronize()
print("checksum", float(out.double().sum()))
