import ast
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

# north_star tolerance: log-mel and IV within 1e-4 relative in fp32, evaluated per feature block
# against the block's max |ref| (SURVEY.md 7.2 "Tolerance definition"); GCC-PHAT 1e-4 absolute.
RTOL_BLOCK = 1e-4


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def make_cfg(sr=24000, hop=240, window='hann', feat='logmelIV', n_mels=64, nfft=1024):
    return {'data': {'sample_rate': sr, 'nfft': nfft, 'hoplen': hop, 'n_mels': n_mels,
                     'window': window, 'audio_feature': feat}}


def block_err(new, ref, sl):
    """max|new-ref| over channel slice `sl`, relative to max|ref| of that block (abs if ref == 0)."""
    d = float(np.abs(np.asarray(new, np.float64)[:, sl] - np.asarray(ref, np.float64)[:, sl]).max())
    m = float(np.abs(np.asarray(ref, np.float64)[:, sl]).max())
    return d / m if m > 0 else d


def assert_blocks_close(new, ref, n_logmel, tol=RTOL_BLOCK, what=''):
    assert new.shape == ref.shape, (new.shape, ref.shape)
    assert np.isfinite(new).all(), what + ': non-finite output'
    e = block_err(new, ref, slice(0, n_logmel))
    assert e <= tol, '%s: log-mel block error %.3e > %.1e' % (what, e, tol)
    if ref.shape[1] > n_logmel:
        e = block_err(new, ref, slice(n_logmel, None))
        assert e <= tol, '%s: IV block error %.3e > %.1e' % (what, e, tol)


@pytest.fixture(scope='session')
def golden_small():
    g = np.load(os.path.join(GOLDEN_DIR, 'foa_small.npz'))
    meta = ast.literal_eval(str(g['meta']))
    return g, meta


@pytest.fixture(scope='session')
def golden_cfg1():
    return np.load(os.path.join(GOLDEN_DIR, 'foa_cfg1_full.npz'))


def golden_input(recipe):
    """Re-create a golden input from its recipe with oracle.synth (tests only)."""
    from oracle import synth
    kind, seed, B, C, L = recipe
    if kind == 'white':
        return synth.white(seed, B, C, L)
    if kind == 'uniform':
        return synth.uniform(seed, (B, C, L))
    if kind == 'plane':
        return synth.plane_wave_foa(seed, B, L)
    if kind == 'zeros':
        return np.zeros((B, C, L), dtype=np.float32)
    if kind == 'half_silent':
        return synth.half_silent(seed, B, C, L)
    if kind == 'quiet':
        return synth.white(seed, B, C, L, scale=3e-5)
    raise KeyError(kind)
