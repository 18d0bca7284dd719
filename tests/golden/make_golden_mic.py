"""Generate tests/golden/mic.npz: MIC-format features (log-mel + GCC-PHAT) from an evaluation that is INDEPENDENT
of oracle/seld_oracle.py and of pseldnets_b200/.

The reference's MIC class (/root/reference/src/utils/feature.py:119-175, driven by preprocess.py:546-556) needs
librosa, which cannot be installed here, so it cannot be run to produce vectors the way the FOA goldens are made.
What can be done is to evaluate the same published definitions with a different library stack and hold both the
numpy restatement (the oracle) and the CUDA kernels to it:

    librosa.stft(center=True, pad_mode='constant', window='hann')  <->  torch.stft(center=True, pad_mode='constant',
                                                                        window=torch.hann_window(periodic=True))
    librosa.filters.mel(sr, n_fft, n_mels) (Slaney scale, Slaney norm) <->  torchaudio.functional.melscale_fbanks(
                                                                        f_min=0, f_max=sr/2, norm='slaney', mel_scale='slaney')
    librosa.power_to_db(ref=1, amin=1e-10, top_db=80)              <->  clamp / log10 / floor at plane max - 80, in torch
    np.fft.irfft(np.exp(1j*np.angle(R)))                           <->  torch.fft.irfft(torch.polar(1, torch.angle(R)))

Nothing here imports numpy's FFT, the oracle, or the package's bank code.  Outputs are stored in fp64 (truth) and
fp32 (what an fp32 pipeline of these library calls gives).  Runs in the build container (needs torchaudio):

    python tests/golden/make_golden_mic.py        # rewrites tests/golden/mic.npz

Planes with a digitally silent microphone: `angle(R)` of an exactly-zero cross-spectrum depends on the SIGNS of the
zeros (atan2(+0, -0) = pi), which differ between FFT libraries; this script records, per case, which GCC planes are
affected (`<case>/signed_zero_planes`) so the tests can treat them separately (tests/test_mic_gpu.py).
"""
import os
import sys

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402  (input recipes only)

N_FFT, N_MELS = 1024, 64


def mic_features(x, sr, hop, dtype, top_db=80.0):
    """x (B, C, L) -> (B, C + C(C-1)/2, int(L/hop), N_MELS), all arithmetic in `dtype`."""
    xt = torch.from_numpy(np.ascontiguousarray(x)).to(dtype)
    B, C, L = xt.shape
    T = int(L / hop)                                                   # preprocess.py:546
    win = torch.hann_window(N_FFT, periodic=True, dtype=dtype)
    X = torch.stft(xt.reshape(-1, L), N_FFT, hop, N_FFT, win, center=True, pad_mode='constant',
                   return_complex=True)                                # feature.py:150
    X = X.reshape(B, C, N_FFT // 2 + 1, -1)[..., :T].transpose(-1, -2)  # feature.py:152  (B, C, T, F)
    fb = torchaudio.functional.melscale_fbanks(N_FFT // 2 + 1, 0.0, sr / 2, N_MELS, sr, norm='slaney',
                                               mel_scale='slaney').to(dtype)   # feature.py:126
    db = 10.0 * torch.log10(torch.clamp((X.abs() ** 2) @ fb, min=1e-10))      # feature.py:158-160
    if top_db is not None:
        db = torch.maximum(db, db.amax(dim=(-2, -1), keepdim=True) - top_db)
    feats = [db[:, c] for c in range(C)]
    for m in range(C):                                                 # feature.py:168-174
        for n in range(m + 1, C):
            R = torch.conj(X[:, m]) * X[:, n]
            cc = torch.fft.irfft(torch.polar(torch.ones_like(R.real), torch.angle(R)), n=N_FFT, dim=-1)
            feats.append(torch.cat((cc[..., -N_MELS // 2:], cc[..., :N_MELS // 2]), dim=-1))
    return torch.stack(feats, dim=1).numpy()


def inputs():
    """name -> (sr, hop, x, planes whose phasors hinge on signed zeros)"""
    out = {}
    out['white_24k'] = (24000, 240, synth.white(51, 2, 4, 3600), [])
    out['uniform_24k'] = (24000, 240, synth.uniform(52, (1, 4, 3600)).astype(np.float32), [])
    s = synth.white(53, 1, 1, 3600 + 64)[0, 0]
    delays = (0, 7, -3, 12)                                            # mic c hears s delayed by delays[c] samples
    out['delay_24k'] = (24000, 240, np.stack([s[32 - d:32 - d + 3600] for d in delays])[None].copy(), [])
    x = synth.white(54, 1, 4, 3600)
    x[:, :, 1800:] *= np.float32(1e-6)                                 # tail 120 dB down: the top_db floor decides
    out['quiet_tail_24k'] = (24000, 240, x, [])
    x = synth.white(55, 1, 4, 3600)
    x[:, :, 1800:] = 0.0                                               # np.pad(..., 'constant') fill: every phasor 1
    out['silent_tail_24k'] = (24000, 240, x, [])
    x = synth.white(56, 1, 4, 3600)
    x[0, 2] = 0.0
    out['dead_mic_24k'] = (24000, 240, x, [1, 3, 5])                   # pairs (0,2) (1,2) (2,3)
    out['white_32k'] = (32000, 320, synth.white(57, 1, 4, 4800), [])
    out['ragged_24k'] = (24000, 240, synth.white(58, 1, 4, 3607), [])
    return out


if __name__ == '__main__':
    store = {}
    names = []
    for name, (sr, hop, x, sz) in inputs().items():
        x = np.ascontiguousarray(x, dtype=np.float32)
        store[name + '/x'] = x
        store[name + '/sr_hop'] = np.array([sr, hop])
        store[name + '/y64'] = mic_features(x, sr, hop, torch.float64)
        store[name + '/y32'] = mic_features(x, sr, hop, torch.float32)
        store[name + '/y64_notopdb'] = mic_features(x, sr, hop, torch.float64, top_db=None)[:, :4]
        store[name + '/signed_zero_planes'] = np.array(sz, dtype=np.int64)
        names.append(name)
        print(name, store[name + '/y64'].shape)
    for sr in (24000, 32000):
        store['bank_%d' % sr] = torchaudio.functional.melscale_fbanks(N_FFT // 2 + 1, 0.0, sr / 2, N_MELS, sr, norm='slaney',
                                                                      mel_scale='slaney').numpy()
    store['names'] = np.array(names)
    store['versions'] = np.array(['torch ' + torch.__version__, 'torchaudio ' + torchaudio.__version__])
    path = os.path.join(HERE, 'mic.npz')
    np.savez_compressed(path, **store)
    print('wrote', path, os.path.getsize(path), 'bytes')
