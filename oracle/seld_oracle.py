"""CPU oracle for the SELD feature front-end -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file; nothing under pseldnets_b200/ does.  It restates, in plain numpy, the
algorithm of /root/reference/src/utils/feature.py:

  * logmel_iv()   <- LogmelIV_Extractor.forward (feature.py:39-56) + intensityvector (:93-117)
  * logmel()      <- Logmel_Extractor.forward (feature.py:76-91)
  * logmel_gcc()  <- Features_Extractor_MIC._spectrogram/_get_logmel_spectrogram/_get_gcc
                     (feature.py:146-175) assembled as preprocess.py:546-556

and the first stage of the backbones that consume the feature map:

  * scalar_eval()    <- the per-channel eval-mode BatchNorm2d "scalar" loop
                        (src/models/accdoa.py:222-227, 318-321; einv2.py:106-109, 292-295)
  * reshape_wav2img()<- HTSAT_Swin_Transformer.reshape_wav2img
                        (src/models/components/htsat.py:493-511)

The arithmetic of those functions lives in third-party packages that are not vendored in the
reference tree: torchaudio==2.2.1 (transforms.Spectrogram / MelScale / AmplitudeToDB),
torch==2.2.1 (torch.stft, matmul), librosa==0.10.1 (stft, filters.mel, power_to_db) and
numpy (fft.irfft) -- /root/reference/requirements.txt:2,4,10,11.  Their published algorithms
are restated here (citations inline).

Pinning: the reference ships no tests, golden vectors or fixtures, so the pin is the reference
itself executed in the build container: tests/golden/make_golden.py imports the unmodified
/root/reference/src/utils/feature.py, runs it on seeded inputs and commits the outputs as
tests/golden/*.npz; tests/test_oracle.py checks this file against those vectors.  The MIC path
needs librosa, which is not installable offline -> its restatement is "parity unpinned"
(checked only against an independent scipy evaluation and its own fp64 mode).  The two
backbone-input functions are pinned by tests/golden/make_golden_epilogue.py, which executes the
reference's own reshape_wav2img and torch.nn.BatchNorm2d loop (tests/golden/epilogue.npz).

`dtype=np.float32` follows the reference's fp32 arithmetic stage by stage (the FFT itself is
numpy/pocketfft in fp32); `dtype=np.float64` is the "truth" evaluation used to bound the
fp32 rounding noise of both the reference and the CUDA path.
"""
import numpy as np

EPS32 = np.float32(np.finfo(np.float32).eps)  # feature.py:8


def _frames(xp, n_fft, hop, n_frames):
    """(..., Lp) -> (..., n_frames, n_fft) strided view (no copy)."""
    shape = xp.shape[:-1] + (n_frames, n_fft)
    strides = xp.strides[:-1] + (hop * xp.strides[-1], xp.strides[-1])
    return np.lib.stride_tricks.as_strided(xp, shape=shape, strides=strides, writeable=False)


def stft(x, window, n_fft, hop, pad_mode, dtype=np.float32, n_frames=None):
    """One-sided STFT of x (..., L) -> (..., T, F) complex, T = 1 + L//hop.

    torch.stft(center=True, pad_mode='reflect', onesided=True, normalized=False) as called by
    torchaudio functional/functional.py:123-134: pad n_fft//2 on both sides, frame t covers
    padded samples [t*hop, t*hop+n_fft), multiply by the window, DFT with e^{-2 pi i k n / N}.
    `pad_mode='constant'` is librosa.stft's 0.10 default (used by the MIC path).
    """
    x = np.asarray(x, dtype=dtype)
    pad = n_fft // 2
    width = [(0, 0)] * (x.ndim - 1) + [(pad, pad)]
    xp = np.pad(x, width, mode=pad_mode)
    T = 1 + x.shape[-1] // hop
    if n_frames is not None:
        T = min(T, n_frames)
    fr = _frames(xp, n_fft, hop, T) * np.asarray(window, dtype=dtype)
    ctype = np.complex64 if dtype == np.float32 else np.complex128
    return np.fft.rfft(fr, n=n_fft, axis=-1).astype(ctype, copy=False)


def _power_to_db(p, amin, dtype):
    """AmplitudeToDB('power', top_db=None) = 10*log10(clamp(x, amin)) - 10*log10(max(amin, 1))
    (torchaudio functional.py:390-391 with ref=1 -> db_multiplier = 0)."""
    return (dtype(10.0) * np.log10(np.maximum(p, dtype(amin)))).astype(dtype)


def logmel(x, window, fb, n_fft, hop, dtype=np.float32):
    """Logmel_Extractor.forward: x (B, C, L) -> (B, C, T, M).  feature.py:88-91."""
    x = np.asarray(x)
    if x.ndim != 3:
        raise ValueError("x shape must be (batch_size, num_channels, data_length)")
    X = stft(x, window, n_fft, hop, 'reflect', dtype)              # (B, C, T, F)
    P = (np.abs(X) ** 2).astype(dtype)                              # torch.abs(x)**2, feature.py:50
    mel = P @ np.asarray(fb, dtype=dtype)                           # MelScale: (.., T, F) @ (F, M)
    return _power_to_db(mel, 1e-10, dtype)


def intensity_vector(X, fb, dtype=np.float32):
    """intensityvector(): X (B, >=4, T, F) complex -> (B, 3, T, M).  feature.py:101-115."""
    re, im = X.real.astype(dtype), X.imag.astype(dtype)
    iv = [re[:, 0] * re[:, j] + im[:, 0] * im[:, j] for j in (1, 2, 3)]
    normal = np.sqrt(iv[0] ** 2 + iv[1] ** 2 + iv[2] ** 2) + dtype(EPS32)
    fbm = np.asarray(fb, dtype=dtype)
    return np.stack([(v / normal) @ fbm for v in iv], axis=1).astype(dtype)


def logmel_iv(x, window, fb, n_fft, hop, dtype=np.float32):
    """LogmelIV_Extractor.forward: x (B, C>=4, L) -> (B, C+3, T, M).  feature.py:49-55."""
    x = np.asarray(x)
    if x.ndim != 3:
        raise ValueError("x shape must be (batch_size, num_channels, data_length)")
    X = stft(x, window, n_fft, hop, 'reflect', dtype)
    P = (np.abs(X) ** 2).astype(dtype)
    lm = _power_to_db(P @ np.asarray(fb, dtype=dtype), 1e-10, dtype)
    iv = intensity_vector(X, fb, dtype)
    return np.concatenate((lm, iv), axis=1)


def logmel_gcc(x, window, mel_bank, n_fft, hop, n_mels=None, top_db=80.0, dtype=np.float32):
    """MIC features: x (B, C, L) -> (B, C + C(C-1)/2, T, M), T = int(L/hop) (preprocess.py:546).

    feature.py:146-153  per-channel librosa.stft (zero 'constant' centre padding, periodic
                        window, complex64 result), cropped to the first T frames;
    feature.py:155-162  |X|^2 @ mel_bank, librosa.power_to_db(ref=1, amin=1e-10, top_db=80):
                        10*log10(max(amin, S)) then floor at (max over the whole (T, M) plane
                        of that channel) - top_db;
    feature.py:164-175  for m < n: R = conj(X_m) X_n, cc = irfft(exp(1j*angle(R))) (n = n_fft,
                        1/N scaling), keep lags [-M/2, M/2) as concat(cc[-M/2:], cc[:M/2]);
    preprocess.py:549-556  concat(logmel, gcc) channel-first.
    The reference takes (L, C) time-major audio; the batch/channel-first layout here is the
    extractor API's.
    """
    x = np.asarray(x)
    if x.ndim != 3:
        raise ValueError("x shape must be (batch_size, num_channels, data_length)")
    B, C, L = x.shape
    M = mel_bank.shape[1] if n_mels is None else n_mels
    T = int(L / hop)
    ctype = np.complex64 if dtype == np.float32 else np.complex128
    X = stft(x, window, n_fft, hop, 'constant', dtype, n_frames=T).astype(ctype)   # (B, C, T, F)
    P = (np.abs(X) ** 2)
    mel = P @ np.asarray(mel_bank, dtype=P.dtype)
    db = 10.0 * np.log10(np.maximum(1e-10, mel))
    if top_db is not None:
        db = np.maximum(db, db.max(axis=(-2, -1), keepdims=True) - top_db)
    feats = [db[:, c] for c in range(C)]
    for m in range(C):
        for n in range(m + 1, C):
            R = np.conj(X[:, m]) * X[:, n]
            ph = np.exp(1.j * np.angle(R)).astype(ctype)
            cc = np.fft.irfft(ph, n=n_fft, axis=-1)
            feats.append(np.concatenate((cc[..., -M // 2:], cc[..., :M // 2]), axis=-1))
    return np.stack(feats, axis=1).astype(dtype)


def _fma32(a, b, c):
    """fp32 fused multiply-add: the product of two fp32 numbers is exact in fp64."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def scalar_eval(x, mean, var, weight, bias, eps=1e-5):
    """Eval-mode "scalar" of the backbones: x (B, C, T, M) -> same shape.

    accdoa.py:222-227: x is viewed as (B, M, T, C) and channel c goes through its own
    BatchNorm2d(M): y = (x - running_mean[c, m]) / sqrt(running_var[c, m] + eps) * weight[c, m]
    + bias[c, m].  Evaluated the way torch's CPU kernel rounds it (bit-exact against
    tests/golden/epilogue.npz): one multiplier and one offset per (c, m),
        a = weight * (1 / sqrt(var + eps)),  b = fma(-mean, a, bias),  y = fma(x, a, b).
    mean / var / weight / bias: (C, M).  fp64 input: plain double arithmetic.
    """
    x = np.asarray(x)
    dt = x.dtype.type
    inv = dt(1.0) / np.sqrt(np.asarray(var, x.dtype) + dt(eps))
    a = np.asarray(weight, x.dtype) * inv
    a4 = a[None, :, None, :]
    if x.dtype == np.float32:
        b = _fma32(-np.asarray(mean, np.float32), a, np.asarray(bias, np.float32))
        return _fma32(x, np.broadcast_to(a4, x.shape), np.broadcast_to(b[None, :, None, :], x.shape))
    b = np.asarray(bias, x.dtype) - np.asarray(mean, x.dtype) * a
    return x * a4 + b[None, :, None, :]


def reshape_wav2img(x, spec_size=256):
    """htsat.py:493-511: x (B, C, T, M) -> (B, C, spec_size, spec_size) "image".

    freq_ratio r = spec_size // M (htsat.py:442); the time axis is zero-padded (or, when longer,
    cropped: F.pad with a negative amount) to r * spec_size frames and cut into r consecutive
    pieces that are stacked along the mel axis:  img[b, c, k*M + m, t] = x[b, c, k*spec_size + t, m].
    """
    x = np.asarray(x)
    B, C, T, M = x.shape
    r = spec_size // M
    target_T = spec_size * r
    xp = np.zeros((B, C, target_T, M), dtype=x.dtype)
    n = min(T, target_T)
    xp[:, :, :n] = x[:, :, :n]
    return np.ascontiguousarray(xp.reshape(B, C, r, target_T // r, M).transpose(0, 1, 2, 4, 3)).reshape(
        B, C, r * M, target_T // r)
