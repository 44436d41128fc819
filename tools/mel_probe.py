"""The mel front-end leg of the bench alone: one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402,F401

from gan_des_midi_music_gen_b200 import benchmark as bm  # noqa: E402

print(json.dumps(bm._mel_leg("cuda", bm._peaks())))
