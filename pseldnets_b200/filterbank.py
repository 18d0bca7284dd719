"""Init-time constants of the SELD feature front-end: analysis windows and mel banks.

These are the two persistent buffers the reference extractors own
(`stft_extractor.window`, `mel_scale.fb`; /root/reference/src/utils/feature.py:28-34)
plus the librosa-style bank of the MIC extractor (feature.py:126).  They are built once on
the host; the CUDA path only ever sees the resulting fp32 tables.

The FOA bank repeats, op for op in fp32 torch, what
torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='htk') does
(torchaudio 2.2.1 functional/functional.py:518-587), so the buffer is bit-identical to the
reference's and checkpoints keep loading (`af_extractor.mel_scale.fb`).
"""
import math

import numpy as np
import torch

# feature.py:9-14 -- the four window names the reference accepts (periodic torch windows).
window_fn_dict = {
    'hann': torch.hann_window,
    'hamming': torch.hamming_window,
    'blackman': torch.blackman_window,
    'bartlett': torch.bartlett_window,
}


def make_window(name, n_fft):
    """torchaudio Spectrogram builds `window_fn(win_length)` (periodic=True default), fp32."""
    assert name in window_fn_dict.keys(), \
        "window must be in {}, but got {}".format(window_fn_dict.keys(), name)
    return window_fn_dict[name](n_fft)


def _hz_to_mel_htk(freq):
    return 2595.0 * math.log10(1.0 + (freq / 700.0))


def melscale_fbanks_htk_slaney(n_freqs, f_min, f_max, n_mels, sample_rate):
    """(n_freqs, n_mels) fp32 triangular bank: HTK mel scale, Slaney area normalisation.

    Same arithmetic order as torchaudio (`all_freqs = linspace(0, sr//2, n_freqs)`,
    `m_pts = linspace(m_min, m_max, n_mels+2)`, `f_pts = 700*(10**(m/2595)-1)`, min of down/up
    slopes clamped at 0, scaled by `2/(f_pts[2:]-f_pts[:-2])`).
    """
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = _hz_to_mel_htk(f_min)
    m_max = _hz_to_mel_htk(f_max)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    zero = torch.zeros(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(zero, torch.min(down_slopes, up_slopes))
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    fb *= enorm.unsqueeze(0)
    return fb


def _hz_to_mel_slaney(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz,
                    min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def librosa_mel_bank(sr, n_fft, n_mels):
    """(n_freqs, n_mels) fp32 bank of the MIC extractor: `librosa.filters.mel(sr, n_fft, n_mels).T`
    (feature.py:126) with librosa 0.10.1 defaults fmin=0, fmax=sr/2, htk=False (Slaney scale),
    norm='slaney', dtype float32.  librosa is not installable here, so its published
    algorithm is restated: fp64 ramps between mel-spaced centre frequencies, lower/upper
    slope minimum clamped at 0, each band scaled by 2/(f[i+2]-f[i]), result cast to fp32.
    """
    fmax = float(sr) / 2
    n_freqs = 1 + n_fft // 2
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_freqs), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return torch.from_numpy(np.ascontiguousarray(weights.T.astype(np.float32)))
