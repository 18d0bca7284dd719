"""Deterministic synthetic waveforms for parity fixtures -- TEST INFRASTRUCTURE.

A counter-based generator (splitmix64 of (seed, index)) written with integer numpy ops only, so
the same (seed, shape) gives bit-identical fp32 samples in the build container (where the
goldens are made from the real reference) and on the GPU box (where they are re-generated and
fed to the CUDA path).  The value families follow SURVEY.md section 4 / 8d.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform(seed, shape):
    """fp32 uniform in [-1, 1) on a 2^-23 grid."""
    n = int(np.prod(shape))
    with np.errstate(over='ignore'):
        idx = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x100000001B3)
        bits = _splitmix64(_splitmix64(idx)) >> np.uint64(40)          # 24 bits
    u = bits.astype(np.float64) * (2.0 / (1 << 24)) - 1.0
    return u.astype(np.float32).reshape(shape)


def normal(seed, shape):
    """Approximately N(0,1) fp32 (Irwin-Hall sum of four uniforms, unit variance)."""
    acc = np.zeros(shape, dtype=np.float64)
    for k in range(4):
        acc += uniform(seed * 4 + k + 1000003, shape).astype(np.float64)
    return (acc * np.sqrt(0.75)).astype(np.float32)


def white(seed, B, C, L, scale=0.1):
    """SURVEY 8d primary distribution: 0.1*N(0,1), all bins populated."""
    return (np.float32(scale) * normal(seed, (B, C, L))).astype(np.float32)


def plane_wave_foa(seed, B, L, noise=1e-3):
    """Coherent FOA plane wave g*s(t) + noise: intensity vector far from zero."""
    s = normal(seed, (B, 1, L)) * np.float32(0.2)
    g = np.array([1.0, 0.55, -0.35, 0.75], dtype=np.float32).reshape(1, 4, 1)
    return (g * s + np.float32(noise) * normal(seed + 77, (B, 4, L))).astype(np.float32)


def half_silent(seed, B, C, L, scale=0.1):
    """Clip whose tail is the np.pad(..., 'constant') zero fill of data.py:86,214."""
    x = white(seed, B, C, L, scale)
    x[..., L // 2:] = 0.0
    return x


def feature_like(seed, B, C, T, M):
    """Stand-in for an extracted feature map: dB-like values in the leading channels, [-1, 1) after."""
    u = uniform(seed, (B, C, T, M))
    x = u.copy()
    n_db = max(1, C - 3)
    x[:, :n_db] = np.float32(40.0) * u[:, :n_db] - np.float32(50.0)
    return x.astype(np.float32)


def scalar_params(seed, C, M):
    """Running statistics / affine terms of C BatchNorm2d(M) "scalar" modules: four (C, M) fp32 arrays."""
    u = uniform(seed, (4, C, M))
    mean = np.float32(30.0) * u[0] - np.float32(20.0)
    var = np.float32(50.0) * (u[1] + np.float32(1.0)) + np.float32(0.01)
    weight = np.float32(1.0) + np.float32(0.5) * u[2]
    bias = np.float32(0.3) * u[3]
    return (mean.astype(np.float32), var.astype(np.float32), weight.astype(np.float32), bias.astype(np.float32))
