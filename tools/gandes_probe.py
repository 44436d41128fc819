"""Runs a few GAN-DES training iterations (for ncu launch lists)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from gan_des_midi_music_gen_b200 import benchmark as bm
print(bm._gandes_leg(torch.device("cuda", 0)))
