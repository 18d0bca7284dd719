"""Backbone-input stage on the GPU, through the C ABI, against the oracle and the goldens
made from the reference's own reshape_wav2img / BatchNorm2d loop (tests/golden/make_golden_epilogue.py).

Bar: bit-exact.  The fold only moves data; the scalar rounds as torch's CPU kernel does
(a = w * (1 / sqrt(var + eps)), b = fma(-mean, a, bias), y = fma(x, a, b)), which the kernel
reproduces operation by operation.
"""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import seld_oracle as oracle
from oracle import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def golden():
    return np.load(os.path.join(HERE, 'golden', 'epilogue.npz'))


def make_scalar(params, eps=1e-5, device='cuda', affine=True):
    mean, var, weight, bias = params
    C, M = mean.shape
    scalar = torch.nn.ModuleList([torch.nn.BatchNorm2d(M, eps=eps, affine=affine) for _ in range(C)])
    for c in range(C):
        scalar[c].running_mean.copy_(torch.from_numpy(mean[c]))
        scalar[c].running_var.copy_(torch.from_numpy(var[c]))
        if affine:
            scalar[c].weight.data.copy_(torch.from_numpy(weight[c]))
            scalar[c].bias.data.copy_(torch.from_numpy(bias[c]))
    return scalar.to(device).eval()


@pytest.mark.parametrize('name', ['pad', 'exact', 'crop', 'tiny', 'r1'])
def test_small_goldens(golden, name):
    import pseldnets_b200 as pb
    seed, B, C, T, M, S = (int(v) for v in golden[name + '/recipe'])
    x = synth.feature_like(seed, B, C, T, M)
    scalar = make_scalar(synth.scalar_params(seed + 1000, C, M))
    xd = torch.from_numpy(x).cuda()
    assert np.array_equal(pb.reshape_wav2img(xd, S).cpu().numpy(), golden[name + '/img'])
    assert np.array_equal(pb.scalar_wav2img(xd, scalar, S).cpu().numpy(), golden[name + '/scalar_img'])
    assert np.array_equal(xd.cpu().numpy(), x), 'the fused call must not touch its input'
    y = pb.apply_scalar(xd, scalar)
    assert y is xd
    assert np.array_equal(xd.cpu().numpy(), golden[name + '/scalar'])


def test_htsat_shape_golden(golden):
    """(B, 7, 1001, 64) -> (B, 7, 256, 256), the shape of every reference config."""
    import pseldnets_b200 as pb
    seed, B, C, T, M, S = (int(v) for v in golden['full/recipe'])
    x = synth.feature_like(seed, B, C, T, M)
    sp = pb.ScalarParams(make_scalar(synth.scalar_params(seed + 1000, C, M)))
    xd = torch.from_numpy(x).cuda()
    img = pb.reshape_wav2img(xd, S).cpu().numpy()
    assert img.shape == (B, C, S, S)
    assert hashlib.sha256(img.tobytes()).digest() == golden['full/img_sha256'].tobytes()
    simg = pb.scalar_wav2img(xd, sp, S).cpu().numpy()
    assert np.array_equal(simg[:, :, ::7, ::5], golden['full/scalar_img_sub'])
    assert np.array_equal(simg, oracle.reshape_wav2img(oracle.scalar_eval(x, *synth.scalar_params(seed + 1000, C, M)), S))
    pb.apply_scalar(xd, sp)
    assert np.array_equal(xd.cpu().numpy()[:, :, ::11, ::3], golden['full/scalar_sub'])


@pytest.mark.parametrize('B,C,T,M,S', [
    (3, 7, 1001, 64, 256),       # reference shape
    (1, 4, 1001, 64, 256),       # Logmel / MIC style channel counts
    (2, 10, 1000, 64, 256),      # MIC: T = 1000
    (1, 7, 1025, 64, 256),       # one frame too many: cropped
    (1, 2, 77, 128, 256),        # two mel tiles, r = 2
    (2, 3, 500, 32, 128),        # partial tiles in both directions
    (1, 1, 40, 20, 60),          # M, S multiples of 4 only; kGeneral scalar kernel (256 % 5 != 0)
    (5, 2, 9, 4, 4),
])
def test_shapes_against_oracle(B, C, T, M, S):
    import pseldnets_b200 as pb
    x = synth.feature_like(7 * B + T, B, C, T, M)
    params = synth.scalar_params(T + M, C, M)
    scalar = make_scalar(params)
    xd = torch.from_numpy(x).cuda()
    assert np.array_equal(pb.reshape_wav2img(xd, S).cpu().numpy(), oracle.reshape_wav2img(x, S))
    xs = oracle.scalar_eval(x, *params)
    assert np.array_equal(pb.scalar_wav2img(xd, scalar, S).cpu().numpy(), oracle.reshape_wav2img(xs, S))
    pb.apply_scalar(xd, scalar)
    assert np.array_equal(xd.cpu().numpy(), xs)


def test_matches_torch_modules_on_gpu():
    """The reference's loop itself, run by torch on the same GPU (cuDNN / native BatchNorm round
    differently from the CPU kernel: tolerance 1e-5 of the map's largest magnitude)."""
    import pseldnets_b200 as pb
    B, C, T, M, S = 2, 7, 1001, 64, 256
    x = torch.from_numpy(synth.feature_like(5, B, C, T, M)).cuda()
    scalar = make_scalar(synth.scalar_params(6, C, M))
    ref = x.clone()
    with torch.no_grad():
        ref = ref.transpose(1, 3)
        for nch in range(ref.shape[-1]):
            ref[..., [nch]] = scalar[nch](ref[..., [nch]])
        ref = ref.transpose(1, 3).contiguous()
    got = pb.apply_scalar(x.clone(), scalar)
    tol = 1e-5 * ref.abs().max().item()
    assert (got - ref).abs().max().item() <= tol


def test_no_affine_and_identity():
    import pseldnets_b200 as pb
    B, C, T, M = 2, 3, 50, 16
    x = synth.feature_like(9, B, C, T, M)
    mean, var, _, _ = synth.scalar_params(10, C, M)
    scalar = make_scalar((mean, var, None, None), affine=False)
    ones, zeros = np.ones_like(mean), np.zeros_like(mean)
    xd = torch.from_numpy(x).cuda()
    pb.apply_scalar(xd, scalar)
    assert np.array_equal(xd.cpu().numpy(), oracle.scalar_eval(x, mean, var, ones, zeros))
    xd = torch.from_numpy(x).cuda()
    assert pb.apply_scalar(xd, None) is xd and np.array_equal(xd.cpu().numpy(), x)


def test_strided_input_and_empty_batch():
    import pseldnets_b200 as pb
    x = torch.from_numpy(synth.feature_like(11, 2, 4, 100, 64)).cuda()
    view = x[:, 1:3]                                         # non-contiguous channel slice
    assert np.array_equal(pb.reshape_wav2img(view, 256).cpu().numpy(),
                          oracle.reshape_wav2img(view.cpu().numpy(), 256))
    with pytest.raises(ValueError):
        pb.apply_scalar(view, make_scalar(synth.scalar_params(1, 2, 64)))
    e = torch.empty((0, 7, 1001, 64), device='cuda')
    assert pb.reshape_wav2img(e, 256).shape == (0, 7, 256, 256)
    assert pb.apply_scalar(e, make_scalar(synth.scalar_params(1, 7, 64))) is e


def test_errors():
    import pseldnets_b200 as pb
    from pseldnets_b200 import _abi
    x = torch.zeros((1, 2, 10, 64), device='cuda')
    with pytest.raises(ValueError):
        pb.reshape_wav2img(x[0], 256)
    with pytest.raises(ValueError):
        pb.reshape_wav2img(x, 100)                           # not a multiple of mel_bins
    with pytest.raises(RuntimeError):
        pb.reshape_wav2img(x.cpu(), 256)
    with pytest.raises(TypeError):
        pb.reshape_wav2img(x.double(), 256)
    with pytest.raises(ValueError):
        pb.apply_scalar(x, make_scalar(synth.scalar_params(1, 3, 64)))       # channel count mismatch
    train = make_scalar(synth.scalar_params(1, 2, 64)).train()
    with pytest.raises(RuntimeError):
        pb.apply_scalar(x, train)
    odd = torch.zeros((1, 2, 10, 6), device='cuda')
    with pytest.raises(_abi.SeldError) as ei:
        pb.reshape_wav2img(odd, 6)
    assert ei.value.code == _abi.SELD_EUNSUPPORTED


def test_extractor_to_image_end_to_end():
    """waveform -> LogmelIV_Extractor -> scalar -> image, against the oracle chain."""
    import pseldnets_b200 as pb
    from pseldnets_b200 import filterbank
    cfg = {'data': {'sample_rate': 24000, 'nfft': 1024, 'hoplen': 240, 'n_mels': 64, 'window': 'hann',
                    'audio_feature': 'logmelIV'}}
    ext = pb.get_afextractor(cfg).cuda()
    x = synth.white(31, 1, 4, 24000)
    feat = ext(torch.from_numpy(x).cuda())
    params = synth.scalar_params(32, 7, 64)
    img = pb.scalar_wav2img(feat, make_scalar(params), 256)
    want = oracle.reshape_wav2img(oracle.scalar_eval(feat.cpu().numpy(), *params), 256)
    assert np.array_equal(img.cpu().numpy(), want)
    assert img.shape == (1, 7, 256, 256) and not np.any(img.cpu().numpy()[:, :, :, 101:])   # 101 frames, rest padded
