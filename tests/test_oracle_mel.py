"""CPU: the mel front-end oracle (oracle/mel_oracle.py, numpy float64) against the golden vectors the UNMODIFIED reference function
GAN_DES/util.py:37-87 produced through torchaudio in the build container (tests/golden/mel_cases.npz, oracle/make_golden.py mel), and the
host-side filter bank of the mirror against the oracle's."""
import os

import numpy as np
import torch

import mel_oracle as mel


def test_mel_oracle_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "mel_cases.npz"))
    for i, (L, sr, seed) in enumerate(g["meta"]):
        w = mel.synth_wave(int(L), int(seed))
        got = mel.get_melspectrogram_db_tensor(w, int(sr))
        want = g[f"db{i}"]
        assert got.shape == want.shape == (128, 216)
        assert np.abs(got - want).max() < 5e-3, (i, np.abs(got - want).max())          # float64 restatement vs torchaudio's float32 pipeline, in dB
        assert abs(want.max() - want.min()) <= 80.0 + 1e-4                               # top_db floor
    p = mel.mel_power(mel.synth_wave(220500, 12), 44100)
    assert np.abs(p - g["power1"]).max() <= 2e-5 * g["power1"].max()                     # util.py:63-87 returns the power spectrogram


def test_mirror_filter_bank_matches_oracle():
    from gan_des_midi_music_gen_b200.GAN_DES import util
    fbt = util._mel_fbanks_t(1025, 20, 8300, 128, 44100, "cpu")
    want = mel.melscale_fbanks(1025, 20.0, 8300.0, 128, 44100)
    assert fbt.shape == (128, 1028) and torch.all(fbt[:, 1025:] == 0)
    assert np.abs(fbt[:, :1025].numpy().T - want).max() < 5e-5            # the mirror evaluates torchaudio's formula in float32, like torchaudio
    assert (fbt.sum(1) > 0).all()
    try:
        import torchaudio.functional as AF
    except Exception:                                                       # torchaudio is optional: the formula check above always runs
        return
    ref = AF.melscale_fbanks(1025, 20.0, 8300.0, 128, 44100, norm=None, mel_scale="htk")
    assert torch.equal(fbt[:, :1025].T.contiguous(), ref)


def test_split_audio_tensor_windows():
    from gan_des_midi_music_gen_b200.GAN_DES import util
    sr = 100
    w = torch.arange(1234, dtype=torch.float32)
    parts = util.split_audio_tensor(w, sr)                       # util.py:103-119: 5 s windows, the last one taken from the end
    assert [len(p) for p in parts] == [500, 500, 500]
    assert parts[0][0] == 0 and parts[1][0] == 500 and parts[2][-1] == 1233 and parts[2][0] == 734
