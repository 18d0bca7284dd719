"""GPU: finite level imbalance between channels that share a packed transform.

The FOA kernels transform two real channels per complex FFT (channels 0/1 and 2/3; in the log-mel-only mode
four consecutive (frame, channel) jobs per warp).  Splitting the packed spectrum leaves each channel with its
partner's fp32 rounding noise (about -130 dB relative to the partner), which the reference -- one real FFT per
channel, `feature.py:49` -- does not have.  Digital silence is covered by test_digitally_silent_channels; these
tests cover everything in between: one channel (or every second job) attenuated by 40 ... 115 dB against its
partner, white and coherent plane-wave input, both channel pairs, FOA and log-mel-only mode, all compared with
the fp64 oracle at the north_star tolerance (1e-4 of the block maximum).  Reference semantics:
`feature.py:50-54, 107-114`."""
import numpy as np
import pytest
import torch

from conftest import RTOL_BLOCK, block_err, make_cfg
import pseldnets_b200 as pb

pytestmark = pytest.mark.gpu

ATTEN_DB = [40, 80, 100, 115]


def _run(ext, x):
    y = ext(torch.from_numpy(np.ascontiguousarray(x)).cuda())
    torch.cuda.synchronize()
    return y.cpu().numpy()


def _check(y, x, ext, n_logmel, iv, what):
    """Within 1e-4 of the block maximum of the fp64 evaluation.  Where the reference's own fp32 arithmetic (the torch
    port, pinned to the reference's goldens) is itself further than that from fp64 -- ill-conditioned normalised
    intensities -- the budget is 1.5 x the reference's own distance: no fp32 implementation can promise more."""
    from oracle import seld_oracle as so, torch_port as tp
    w, fb = ext.stft_extractor.window.cpu(), ext.mel_scale.fb.cpu()
    ref = (so.logmel_iv if iv else so.logmel)(x, w.numpy(), fb.numpy(), 1024, 240, np.float64)
    r32 = (tp.logmel_iv if iv else tp.logmel)(torch.from_numpy(x), w, fb, 1024, 240).numpy()
    assert y.shape == ref.shape and np.isfinite(y).all(), what
    blocks = [('log-mel', slice(0, n_logmel))] + ([('IV', slice(n_logmel, None))] if iv else [])
    for name, sl in blocks:
        tol = max(RTOL_BLOCK, 1.5 * block_err(r32, ref, sl))
        e = block_err(y, ref, sl)
        assert e <= tol, '%s: %s block error %.3e > %.1e' % (what, name, e, tol)


def _signal(kind, seed, B, L):
    from oracle import synth
    if kind == 'white':
        return synth.uniform(seed, (B, 4, L)).astype(np.float32)                 # full scale
    return (4.0 * synth.plane_wave_foa(seed, B, L, noise=1e-3)).astype(np.float32)  # coherent, IV far from 0


@pytest.mark.parametrize('kind', ['white', 'plane'])
@pytest.mark.parametrize('quiet', [0, 1, 2, 3])
@pytest.mark.parametrize('db', ATTEN_DB)
def test_foa_channel_attenuated_against_its_partner(kind, quiet, db):
    ext = pb.LogmelIV_Extractor(make_cfg()).cuda()
    x = _signal(kind, 4100 + 10 * quiet + db, 2, 6000)
    x[:, quiet] *= np.float32(10.0 ** (-db / 20.0))
    x[1, quiet, 3000:] = 0.0                                                     # ... and a digitally silent tail in clip 1
    _check(_run(ext, x), x, ext, 4, True, '%s, channel %d at -%d dB' % (kind, quiet, db))


@pytest.mark.parametrize('C', [1, 2, 4])
@pytest.mark.parametrize('db', ATTEN_DB)
def test_logmel_only_jobs_attenuated_against_their_partner(C, db):
    """Log-mel-only mode: the four transform slots of a warp take consecutive (frame, channel) jobs, so partners
    are neighbouring channels (C >= 2) or neighbouring frames (C = 1: a loud passage next to a quiet one)."""
    from oracle import synth
    ext = pb.Logmel_Extractor(make_cfg(feat='logmel')).cuda()
    L = 9600
    x = synth.uniform(4300 + C + db, (2, C, L)).astype(np.float32)
    g = np.float32(10.0 ** (-db / 20.0))
    if C == 1:
        x[:, :, L // 2:] *= g                                                    # level step in time
    else:
        x[:, 1::2] *= g                                                          # every second channel quiet
    _check(_run(ext, x), x, ext, C, False, 'log-mel only, C = %d, partner at -%d dB' % (C, db))


def test_band_limited_partner():
    """A channel that is quiet only in part of the spectrum (low-passed) next to a full-band partner: the
    imbalance exists per mel band, not per channel."""
    from oracle import synth
    ext = pb.LogmelIV_Extractor(make_cfg()).cuda()
    x = synth.uniform(4500, (1, 4, 7200)).astype(np.float32)
    # crude low-pass of channel 1: 64-tap moving average twice (-100 dB and below above a few kHz is not reached,
    # but 50-70 dB is), plus a channel 3 that only has a low tone
    k = np.ones(64, np.float32) / 64.0
    x[0, 1] = np.convolve(np.convolve(x[0, 1], k, 'same'), k, 'same')
    t = np.arange(7200, dtype=np.float64)
    x[0, 3] = (0.5 * np.sin(2 * np.pi * 200.0 / 24000.0 * t)).astype(np.float32)
    _check(_run(ext, x), x, ext, 4, True, 'band-limited partners')
