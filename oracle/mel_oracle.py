"""ORACLE (test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this).

numpy float64 restatement of the GAN-DES mel front end, /root/reference/GAN_DES/util.py:37-61 (``get_melspectrogram_db_tensor``) and
:63-87 (``..._maestro``, which returns the mel power WITHOUT the dB step), i.e. of what torchaudio 2.x computes for
``T.MelSpectrogram(sample_rate, n_fft, hop_length, n_mels, f_min, f_max)`` + ``T.AmplitudeToDB(top_db)`` with their defaults:
win_length = n_fft, periodic Hann window, centre = True with reflect padding, power = 2, one-sided, not normalised; HTK mel scale without
area normalisation (torchaudio.functional.melscale_fbanks); dB = 10 log10(max(x, 1e-10)), floored at (max of the spectrogram) - top_db.
Pinned by tests/golden/mel_cases.npz (oracle/make_golden.py: the UNMODIFIED reference function run through torchaudio in the build container).
"""
import numpy as np


def melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') -> (n_freqs, n_mels)"""
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * np.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * np.log10(1.0 + f_max / 700.0)
    m_pts = np.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up))


def mel_power(wave, sr=44100, n_fft=2048, n_mels=128, fmin=20, fmax=8300, mel_length=216):
    """util.py:40-57: hop from the length, crop, MelSpectrogram -> (n_mels, T) float64"""
    wave = np.asarray(wave, dtype=np.float64)
    hop = len(wave) // (mel_length - 1)
    wave = wave[:mel_length * hop]
    pad = n_fft // 2
    x = np.pad(wave, (pad, pad), mode="reflect")
    T = 1 + len(wave) // hop
    n = np.arange(n_fft)
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)
    frames = np.stack([x[t * hop:t * hop + n_fft] for t in range(T)]) * win
    spec = np.abs(np.fft.rfft(frames, axis=1)) ** 2                      # (T, n_fft/2 + 1)
    fb = melscale_fbanks(n_fft // 2 + 1, float(fmin), float(fmax), n_mels, sr)
    return (spec @ fb).T


def amplitude_to_db(x, top_db=80.0):
    db = 10.0 * np.log10(np.maximum(x, 1e-10))
    if top_db is not None:
        db = np.maximum(db, db.max() - top_db)
    return db


def get_melspectrogram_db_tensor(wave, sr=44100, n_fft=2048, hop_length=512, n_mels=128, fmin=20, fmax=8300, top_db=80, mel_length=216):
    return amplitude_to_db(mel_power(wave, sr, n_fft, n_mels, fmin, fmax, mel_length), top_db)


def synth_wave(L, seed):
    """a deterministic test signal: a few decaying partials + noise, float32"""
    rng = np.random.default_rng(seed)
    t = np.arange(L) / 44100.0
    x = 0.02 * rng.standard_normal(L)
    for _ in range(6):
        f0 = rng.uniform(60.0, 6000.0)
        x += rng.uniform(0.05, 0.4) * np.sin(2 * np.pi * f0 * t + rng.uniform(0, 6.28)) * np.exp(-t * rng.uniform(0.2, 3.0))
    return x.astype(np.float32)
