"""Times the rasteriser paths (stream / sort) on BASELINE config 4 and on the e2e call shape (used under ncu for the per-kernel split)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import raster_oracle as ro
from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for S, E, T, seq, W, p_on in ((1276, 15000, 300.0, 300, 300, 0.2), (32768, 320, 60.0, 100, 50, 0.4)):
    dt, meta, off = ro.synth_songs(S, E, T, seed=0, p_on=p_on, p_off=p_on)
    d = [torch.from_numpy(a).cuda() for a in (dt, meta.view(np.int32), off)]
    for dtype in (torch.float32, torch.uint8):
        ref = None
        for path in ("sort", "stream"):
            for _ in range(3):
                out = ds.rasterize_events(*d, seq, 0, W, dtype, path=path)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); out = ds.rasterize_events(*d, seq, 0, W, dtype, path=path); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            same = "" if ref is None else f" equal_to_sort={bool(torch.equal(out, ref))}"
            ref = out if ref is None else ref
            print(f"S={S} E={E} W={W} {dtype} {path}: ms {min(ts):.4f} checksum {float(out.double().sum())}{same}", flush=True)
