"""Torch-CPU port of the reference extractors -- TEST / BASELINE INFRASTRUCTURE ONLY.

Same library calls, in the same order, as /root/reference/src/utils/feature.py makes through
torchaudio (Spectrogram -> torch.stft, MelScale -> matmul, AmplitudeToDB -> clamp/log10), so
on a given host it reproduces the reference's CPU output bit for bit and costs what the
reference costs.  bench.py times it as `cpu_baseline` (kind "port") and as the
`--impl reference` arm, because the reference itself (a Python tree under /root/reference)
cannot travel to the GPU box.  Never imported by pseldnets_b200/.
"""
import torch

EPS = torch.finfo(torch.float32).eps  # feature.py:8


def _spectrogram(x, window, n_fft, hop):
    # torchaudio functional/functional.py:117-141 (power=None, normalized=False, pad=0)
    shape = x.size()
    spec = torch.stft(x.reshape(-1, shape[-1]), n_fft=n_fft, hop_length=hop, win_length=n_fft,
                      window=window, center=True, pad_mode='reflect', normalized=False,
                      onesided=True, return_complex=True)
    return spec.reshape(shape[:-1] + spec.shape[-2:])


def _mel_scale(spec, fb):
    # torchaudio transforms/_transforms.py:417
    return torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)


def _amp2db(x):
    # torchaudio functional/functional.py:390-391 with multiplier=10, amin=1e-10, db_multiplier=0
    x_db = 10.0 * torch.log10(torch.clamp(x, min=1e-10))
    x_db -= 10.0 * 0.0
    return x_db


def intensityvector(spec_re, spec_im, fb):
    """(B, >=4, T, F) real / imaginary parts -> (B, 3, T, M).  Same operations, in the same order, as
    feature.py:101-115: three real cross terms with channel 0, their Euclidean norm + eps, three divisions,
    three (T, F) @ (F, M) products, stacked."""
    w_re, w_im = spec_re[:, 0], spec_im[:, 0]
    cross = [w_re * spec_re[:, j] + w_im * spec_im[:, j] for j in (1, 2, 3)]
    norm = torch.sqrt(cross[0] ** 2 + cross[1] ** 2 + cross[2] ** 2) + EPS
    return torch.stack([torch.matmul(c / norm, fb) for c in cross], dim=1)


@torch.no_grad()
def logmel_iv(x, window, fb, n_fft, hop):
    # feature.py:46-56
    if x.ndim != 3:
        raise ValueError("x shape must be (batch_size, num_channels, data_length)")
    X = _spectrogram(x, window, n_fft, hop)
    mel = _mel_scale(torch.abs(X) ** 2, fb)
    logmel = _amp2db(mel).transpose(-1, -2)
    iv = intensityvector(X.real.transpose(-1, -2), X.imag.transpose(-1, -2), fb)
    return torch.cat((logmel, iv), dim=1)


@torch.no_grad()
def logmel(x, window, fb, n_fft, hop):
    # feature.py:85-91
    if x.ndim != 3:
        raise ValueError("x shape must be (batch_size, num_channels, data_length)")
    X = _spectrogram(x, window, n_fft, hop)
    mel = _mel_scale(torch.abs(X) ** 2, fb)
    return _amp2db(mel).transpose(-1, -2)
