"""GPU parity of the GAN-DES mel front end (GAN_DES/util.py mirror; csrc/mel.cu + the tf32 projection GEMM) against the golden vectors of the
UNMODIFIED reference function (tests/golden/mel_cases.npz) and the float64 oracle.  Tolerances: 0.02 dB on the dB spectrograms (fp32 FFT;
the projection rounds its operands to tf32: all terms are positive, so the relative error of a mel bin is <= 2^-10 = 0.004 dB), 2e-3 of the
maximum on the power spectrogram."""
import os

import numpy as np
import pytest
import torch

import mel_oracle as mel

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_mel_db_vs_reference_golden(golden_dir):
    from gan_des_midi_music_gen_b200.GAN_DES import util
    g = np.load(os.path.join(golden_dir, "mel_cases.npz"))
    for i, (L, sr, seed) in enumerate(g["meta"]):
        w = torch.from_numpy(mel.synth_wave(int(L), int(seed))).to(DEV)
        got = util.get_melspectrogram_db_tensor(w, int(sr)).cpu().numpy()
        want = g[f"db{i}"]
        assert got.shape == want.shape
        assert np.abs(got - want).max() < 2e-2, (i, np.abs(got - want).max())
    w = torch.from_numpy(mel.synth_wave(220500, 12)).to(DEV)
    p = util.get_melspectrogram_db_tensor_maestro(w, 44100).cpu().numpy()
    assert np.abs(p - g["power1"]).max() <= 2e-3 * g["power1"].max()


def test_mel_batch_equals_single_calls_and_oracle():
    from gan_des_midi_music_gen_b200.GAN_DES import util
    L, B = 220500, 5                                              # MaestroDataset.__getitem__: 5 s windows at 44.1 kHz (datasets.py:85-90)
    waves = np.stack([mel.synth_wave(L, 40 + b) for b in range(B)])
    waves[3] *= 1e-3                                              # a quiet window: its floor sits 80 dB under ITS OWN maximum
    x = torch.from_numpy(waves).to(DEV)
    got = util.get_melspectrogram_db_tensor(x)
    assert got.shape == (B, 128, 216)
    for b in range(B):
        one = util.get_melspectrogram_db_tensor(x[b])
        assert torch.equal(one, got[b])
        want = mel.get_melspectrogram_db_tensor(waves[b])
        assert np.abs(got[b].cpu().numpy() - want).max() < 2e-2, b
    with pytest.raises(ValueError):
        util.melspectrogram_batch(torch.zeros(2, 100, device=DEV))
    with pytest.raises(ValueError):                              # reflect padding needs more than n_fft / 2 samples (torch.stft raises as well)
        util.melspectrogram_batch(torch.zeros(1, 900, device=DEV))
