"""Drop-in mirror of the tensor half of the reference's ``GAN_DES/util.py`` (/root/reference/GAN_DES/util.py:37-87, :103-119): the mel front
end that turns audio windows into the (128, 216) dB spectrograms the GAN-DES discriminator reads -- on the device.

``get_melspectrogram_db_tensor(waveform, sr, ...)`` keeps the reference's signature and quirks (the hop length is derived from the length:
``hop = len // (mel_length - 1)``, the ``hop_length`` argument is ignored; the ``_maestro`` variant returns the POWER spectrogram, its dB step
is computed and dropped).  Three launches: framing + Hann window + 2048-point FFT + |X|^2 (csrc/mel.cu), the mel projection as a tf32
tcgen05 GEMM stored straight to (B, n_mels, T) (csrc/gemm_tc.cu), dB conversion with the per-spectrogram ``top_db`` floor.  A (B, L) batch of
equal-length windows (what ``MaestroDataset.__getitem__`` loops over, datasets.py:85-90) goes through the same three launches at once.

The librosa / file based helpers (util.py:8-35, :89-100) need librosa and an audio decoder, which this image lacks: out of scope.
"""
import torch

from .. import _native as N

__all__ = ["get_melspectrogram_db_tensor", "get_melspectrogram_db_tensor_maestro", "melspectrogram_batch", "split_audio_tensor"]

_fb_cache = {}


def _mel_fbanks_t(n_freqs, f_min, f_max, n_mels, sample_rate, device):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk"), float32 like torchaudio, TRANSPOSED to (n_mels, pitch) with the row
    pitch padded to 16 bytes: the K-major N operand of the projection GEMM."""
    key = (n_freqs, float(f_min), float(f_max), n_mels, int(sample_rate), str(device))
    fbt = _fb_cache.get(key)
    if fbt is None:
        import math
        all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
        m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
        m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
        m_pts = torch.linspace(m_min, m_max, n_mels + 2)
        f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
        f_diff = f_pts[1:] - f_pts[:-1]
        slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
        down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
        up = slopes[:, 2:] / f_diff[1:]
        fb = torch.max(torch.zeros(1), torch.min(down, up))                 # (n_freqs, n_mels)
        pitch = (n_freqs + 3) // 4 * 4
        fbt = torch.zeros(n_mels, pitch)
        fbt[:, :n_freqs] = fb.T
        fbt = _fb_cache[key] = fbt.to(device)
    return fbt


def melspectrogram_batch(waveforms, sr=44100, n_fft=2048, n_mels=128, fmin=20, fmax=8300, top_db=80, mel_length=216, db=True):
    """(B, L) float32 CUDA windows -> (B, n_mels, T) dB (or power, ``db=False``) spectrograms, T = 1 + min(L, mel_length * hop) // hop."""
    N.require_cuda(waveforms)
    if waveforms.dim() != 2:
        raise ValueError("melspectrogram_batch expects (B, L)")
    x = waveforms.float()
    B, L = x.shape
    hop = L // (mel_length - 1)
    if hop < 1:
        raise ValueError(f"waveform of {L} samples is shorter than mel_length - 1 = {mel_length - 1}")
    x = x[:, :mel_length * hop].contiguous()
    L = x.shape[1]
    T = 1 + L // hop
    n_freqs = n_fft // 2 + 1
    pitch = (n_freqs + 3) // 4 * 4
    power = torch.empty(B * T, pitch, device=x.device)
    N.call("mmg_stft_power_f32", N.ptr(x), B, L, x.stride(0), n_fft, hop, N.ptr(power), pitch, N.stream())
    fbt = _mel_fbanks_t(n_freqs, fmin, fmax, n_mels, sr, x.device)
    mel = torch.empty(B, n_mels, T, device=x.device)
    N.call("mmg_gemm_tc", N.ptr(power), 0, pitch, N.ptr(fbt), 0, pitch, N.ptr(mel), n_mels, B * T, n_mels, n_freqs, 1, 1, 1, T, 0, None, 0, 0, N.stream())
    if not db:
        return mel
    out = torch.empty_like(mel)
    N.call("mmg_power_to_db_f32", N.ptr(mel), N.ptr(out), B, n_mels * T, -1.0 if top_db is None else float(top_db), N.stream())
    return out


def get_melspectrogram_db_tensor(waveform, sr=44100, n_fft=2048, hop_length=512, n_mels=128, fmin=20, fmax=8300, top_db=80, mel_length=216):
    """util.py:37-61 -- 1-D waveform -> (n_mels, T) dB mel spectrogram (a (B, L) batch gives (B, n_mels, T))."""
    if waveform.dim() == 1:
        return melspectrogram_batch(waveform.unsqueeze(0), sr, n_fft, n_mels, fmin, fmax, top_db, mel_length)[0]
    return melspectrogram_batch(waveform, sr, n_fft, n_mels, fmin, fmax, top_db, mel_length)


def get_melspectrogram_db_tensor_maestro(waveform, sr=44100, n_fft=2048, hop_length=512, n_mels=128, fmin=20, fmax=8300, top_db=80, mel_length=216):
    """util.py:63-87 -- the same pipeline, but the reference returns ``mel_spectrogram`` (power), not the dB tensor it has just computed."""
    if waveform.dim() == 1:
        return melspectrogram_batch(waveform.unsqueeze(0), sr, n_fft, n_mels, fmin, fmax, top_db, mel_length, db=False)[0]
    return melspectrogram_batch(waveform, sr, n_fft, n_mels, fmin, fmax, top_db, mel_length, db=False)


def split_audio_tensor(waveform, sample_rate, hop_length_audio=5, window_size=5):
    """util.py:103-119 without the file read: a mono waveform cut into ``window_size``-second windows every ``hop_length_audio`` seconds, the
    last one taken from the end so that it is as long as the others.  Returns a list of views."""
    n, step, win = waveform.shape[-1], hop_length_audio * sample_rate, window_size * sample_rate
    out = []
    for i in range(0, n + 1, step):
        out.append(waveform[..., -win:] if i + step > n else waveform[..., i:i + win])
    return out
