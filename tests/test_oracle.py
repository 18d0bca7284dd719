"""CPU: pin the oracle (numpy restatement + torch port) to golden vectors made from the real
reference (tests/golden/make_golden.py).  The reference ships no vectors of its own."""
import os

import numpy as np
import pytest
import torch

from conftest import assert_blocks_close, block_err, golden_input
from oracle import seld_oracle as so
from oracle import seld_oracle as oracle
from oracle import synth
from oracle import torch_port as tp
from pseldnets_b200 import filterbank as fbk


def _tables(sr, win):
    w = fbk.make_window(win, 1024)
    fb = fbk.melscale_fbanks_htk_slaney(513, 20, sr / 2, 64, sr)
    return w, fb


def test_synth_inputs_are_reproducible(golden_small):
    g, meta = golden_small
    for name, kind, sr, hop, win, recipe in meta:
        assert np.array_equal(golden_input(recipe), g[name + '/x']), name


def test_numpy_oracle_fp64_is_the_reference_algorithm(golden_small):
    """fp64 evaluation agrees with the reference module run in fp64 to ~1e-13: same algorithm."""
    g, meta = golden_small
    for name, kind, sr, hop, win, recipe in meta:
        w, fb = _tables(sr, win)
        f = so.logmel_iv if kind == 'logmelIV' else so.logmel
        y = f(g[name + '/x'], w.numpy(), fb.numpy(), 1024, hop, np.float64)
        C = g[name + '/x'].shape[1]
        assert_blocks_close(y, g[name + '/y64'], C, tol=1e-11, what=name)


def test_numpy_oracle_fp32_matches_reference_fp32(golden_small):
    g, meta = golden_small
    for name, kind, sr, hop, win, recipe in meta:
        w, fb = _tables(sr, win)
        f = so.logmel_iv if kind == 'logmelIV' else so.logmel
        y = f(g[name + '/x'], w.numpy(), fb.numpy(), 1024, hop, np.float32)
        assert y.dtype == np.float32
        C = g[name + '/x'].shape[1]
        assert_blocks_close(y, g[name + '/y32'], C, tol=1e-5, what=name)


def test_torch_port_matches_reference_fp32(golden_small):
    """Same library calls as the reference -> same numbers up to the host's FFT/BLAS rounding."""
    g, meta = golden_small
    for name, kind, sr, hop, win, recipe in meta:
        w, fb = _tables(sr, win)
        f = tp.logmel_iv if kind == 'logmelIV' else tp.logmel
        y = f(torch.from_numpy(g[name + '/x']), w, fb, 1024, hop).numpy()
        C = g[name + '/x'].shape[1]
        assert_blocks_close(y, g[name + '/y32'], C, tol=1e-5, what=name)


def test_oracle_cfg1_full_size(golden_cfg1):
    from oracle import synth
    g = golden_cfg1
    x = synth.white(1234, 1, 4, 240000)
    w, fb = _tables(24000, 'hann')
    y = so.logmel_iv(x, w.numpy(), fb.numpy(), 1024, 240, np.float32)
    assert tuple(y.shape) == tuple(g['shape'])
    assert_blocks_close(y[:, :, g['frames']], g['y32'], 4, tol=1e-5, what='cfg1')
    y64 = so.logmel_iv(x, w.numpy(), fb.numpy(), 1024, 240, np.float64)
    np.testing.assert_allclose(y64.sum(axis=(2, 3)), g['sum64'], rtol=1e-9)
    np.testing.assert_allclose((y64 ** 2).sum(axis=(2, 3)), g['sumsq64'], rtol=1e-9)


def test_oracle_edge_semantics():
    w, fb = _tables(24000, 'hann')
    z = np.zeros((1, 4, 2400), np.float32)
    y = so.logmel_iv(z, w.numpy(), fb.numpy(), 1024, 240)
    # silence -> amin clamp (-100 dB up to fp32 log10 rounding) and exactly zero IV
    assert np.abs(y[:, :4] + 100.0).max() < 1e-4 and np.all(y[:, 4:] == 0.0)
    with pytest.raises(ValueError):
        so.logmel_iv(np.zeros((4, 2400), np.float32), w.numpy(), fb.numpy(), 1024, 240)
    with pytest.raises(ValueError):
        tp.logmel(torch.zeros(4, 2400), w, fb, 1024, 240)


def test_mic_oracle_self_consistency():
    """MIC restatement ("parity unpinned": librosa is not installable): fp32 vs fp64 evaluation,
    shapes, GCC of a pure delay peaks at that lag, top_db floor."""
    from oracle import synth
    sr, hop = 24000, 240
    w = fbk.make_window('hann', 1024).numpy()
    bank = fbk.librosa_mel_bank(sr, 1024, 64).numpy()
    x = synth.white(5, 1, 1, 4800 + 8)[0, 0]
    d = 5
    mics = np.stack([x[8:8 + 4800], x[8 - d:8 - d + 4800], x[8:8 + 4800], x[8 - 2:8 - 2 + 4800]])[None]
    y32 = so.logmel_gcc(mics, w, bank, 1024, hop, dtype=np.float32)
    y64 = so.logmel_gcc(mics, w, bank, 1024, hop, dtype=np.float64)
    assert y32.shape == (1, 10, 20, 64)
    assert block_err(y32, y64, slice(0, 4)) < 1e-5
    assert np.abs(y32[:, 4:] - y64[:, 4:]).max() < 1e-4
    # pair (0,1): mic 1 lags mic 0 by d samples -> peak at lag +d -> column 32 + d
    assert int(np.argmax(y64[0, 4, 10])) == 32 + d
    assert int(np.argmax(y64[0, 5, 10])) == 32          # identical channels: lag 0
    assert y64[0, :4].min() >= y64[0, :4].max() - 80.0 - 1e-9


def test_epilogue_oracle_matches_reference_goldens():
    """scalar_eval / reshape_wav2img against the reference's own BatchNorm2d loop and
    HTSAT_Swin_Transformer.reshape_wav2img (tests/golden/make_golden_epilogue.py): bit-exact."""
    import hashlib
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'epilogue.npz'))
    for name in ['pad', 'exact', 'crop', 'tiny', 'r1']:
        seed, B, C, T, M, S = (int(v) for v in g[name + '/recipe'])
        x = synth.feature_like(seed, B, C, T, M)
        xs = oracle.scalar_eval(x, *synth.scalar_params(seed + 1000, C, M), 1e-5)
        assert np.array_equal(oracle.reshape_wav2img(x, S), g[name + '/img']), name
        assert np.array_equal(xs, g[name + '/scalar']), name
        assert np.array_equal(oracle.reshape_wav2img(xs, S), g[name + '/scalar_img']), name
    seed, B, C, T, M, S = (int(v) for v in g['full/recipe'])
    x = synth.feature_like(seed, B, C, T, M)
    img = oracle.reshape_wav2img(x, S)
    assert img.shape == (B, C, S, S)
    assert hashlib.sha256(img.tobytes()).digest() == g['full/img_sha256'].tobytes()
    xs = oracle.scalar_eval(x, *synth.scalar_params(seed + 1000, C, M), 1e-5)
    assert np.array_equal(xs[:, :, ::11, ::3], g['full/scalar_sub'])
    assert np.array_equal(oracle.reshape_wav2img(xs, S)[:, :, ::7, ::5], g['full/scalar_img_sub'])


def test_epilogue_oracle_fold_semantics():
    x = np.arange(2 * 1 * 10 * 4, dtype=np.float32).reshape(2, 1, 10, 4)
    img = oracle.reshape_wav2img(x, 8)                       # r = 2, target_T = 16: 6 frames of padding
    assert img.shape == (2, 1, 8, 8)
    assert img[1, 0, 2, 3] == x[1, 0, 3, 2]                  # piece 0: row m, column t
    assert img[0, 0, 4 + 1, 1] == x[0, 0, 8 + 1, 1]          # piece 1 sits below piece 0
    assert not img[:, :, 4:, 2:].any()                       # frames 10..15 are zero padding


def _mic_golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'mic.npz'))


def test_mic_bank_against_torchaudio_slaney():
    """`librosa.filters.mel` restated (filterbank.librosa_mel_bank, feature.py:126) against torchaudio's independent
    Slaney-scale / Slaney-norm bank stored in tests/golden/mic.npz."""
    g = _mic_golden()
    for sr in (24000, 32000):
        ours = fbk.librosa_mel_bank(sr, 1024, 64).numpy()
        ref = g['bank_%d' % sr]
        assert ours.shape == ref.shape == (513, 64)
        assert np.abs(ours - ref).max() <= 1e-7 * 1.0 and np.abs(ours - ref).max() / np.abs(ref).max() <= 5e-6
        assert np.array_equal(ours != 0, ref != 0) or np.abs(ours[(ours != 0) != (ref != 0)]).max() < 1e-7


def test_mic_oracle_is_pinned_to_the_independent_evaluation():
    """The MIC oracle (numpy restatement of feature.py:146-175 + preprocess.py:546-556) against vectors produced by a
    different library stack (torch.stft / torchaudio bank / torch.fft.irfft: tests/golden/make_golden_mic.py) on
    eight seeded inputs: white, full scale, pure delays, 120 dB quiet tail (top_db floor), zero-filled tail, dead
    microphone, 32 kHz, ragged length.  fp64 oracle: log-mel 1e-5 of the block maximum, GCC 1e-6 absolute; fp32 oracle:
    the north_star tolerances."""
    g = _mic_golden()
    for name in g['names']:
        sr, hop = (int(v) for v in g[name + '/sr_hop'])
        x, y64 = g[name + '/x'], g[name + '/y64']
        w = fbk.make_window('hann', 1024).numpy()
        bank = fbk.librosa_mel_bank(sr, 1024, 64).numpy()
        sz = set(int(p) for p in g[name + '/signed_zero_planes'])
        ok = [4 + p for p in range(6) if p not in sz]
        for dtype, tol_lm, tol_gcc in ((np.float64, 1e-5, 1e-6), (np.float32, 1e-4, 1e-4)):
            y = so.logmel_gcc(x, w, bank, 1024, hop, dtype=dtype)
            assert y.shape == y64.shape, name
            assert block_err(y, y64, slice(0, 4)) <= tol_lm, (name, dtype)
            assert np.abs(y[:, ok].astype(np.float64) - y64[:, ok]).max() <= tol_gcc, (name, dtype)
        # without the top_db floor (power_to_db(top_db=None))
        y = so.logmel_gcc(x, w, bank, 1024, hop, top_db=None, dtype=np.float64)
        assert block_err(y[:, :4], g[name + '/y64_notopdb'], slice(0, 4)) <= 1e-5, name
        if sz:
            # digitally silent microphone: the cross-spectrum is an exact zero whose SIGN pattern decides angle() --
            # numpy and torch do not agree with each other there (the reference is implementation-defined)
            bad = [4 + p for p in sorted(sz)]
            y = so.logmel_gcc(x, w, bank, 1024, hop, dtype=np.float64)
            assert np.abs(y[:, bad] - y64[:, bad]).max() > 1e-2
