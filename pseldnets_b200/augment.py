"""Waveform-domain augmentation of the staged batch, on the CUDA path.

In the reference's training step the batch goes  Rotation -> WavMix -> extractor
(src/models/model_module.py:53-58).  Rotation rewrites the waveform clip by clip in a Python loop
(torch.stack of four signed channel views, src/augment/rotate.py:10-44, 47-99); WavMix gathers two
sets of clips, blends them and scatters the result back (src/augment/wavmix.py:50).  Here the
waveform arithmetic of the whole batch is one kernel launch each, in place and bit-identical:

    rotate_waveforms(batch_x, codes)                 one launch for all rotated clips
    wavmix_waveforms(batch_x, dst, src, lambs)       wavmix.py:50

and `Rotation` is a drop-in for the reference class (same constructor, same call signature, same
random draws in the same order from the same generators, same label bookkeeping) that defers the
waveform work of its loop to one `rotate_waveforms` call.  WavMix's label bookkeeping
(wavmix.py:52-116) is control plane and stays with the caller: replace its line 50 by
`wavmix_waveforms(batch_x, idx_ov1[:N], new_idx_ov[:N], lambs)`.
No CPU fallback: CUDA tensors only.
"""
import ctypes
import random

import numpy as np
import torch

from . import _abi

ROT_IDENTITY = -1

# rotate.py:61-68 / 88-91: DOA axis order (xx, yy, zz) -> waveform channels (s_x, s_y, s_z); the new clip is
# stack(x[0], sign_y * x[s_x], sign_z * x[s_y], sign_x * x[s_z])
TRANS_48 = {(0, 1, 2): (1, 2, 3), (0, 2, 1): (2, 1, 3), (1, 0, 2): (3, 2, 1),
            (1, 2, 0): (2, 3, 1), (2, 0, 1): (3, 1, 2), (2, 1, 0): (1, 3, 2)}
TRANS_16 = {(0, 1, 2): (1, 2, 3), (1, 0, 2): (3, 2, 1)}


def rotation_code(sources, signs):
    """int32 code of one clip for seld_foa_rotate_f32 (SELD_ROT_CODE): `sources` = the channels (1..3)
    that become output channels 1, 2, 3; `signs` = their factors (+1 / -1)."""
    s1, s2, s3 = (int(s) for s in sources)
    if not all(1 <= s <= 3 for s in (s1, s2, s3)):
        raise ValueError('rotation sources must be channels 1..3, got %r' % (sources,))
    if not all(int(g) in (-1, 1) for g in signs):
        raise ValueError('rotation signs must be +1 / -1, got %r' % (signs,))
    n1, n2, n3 = (int(g) < 0 for g in signs)
    return s1 | (s2 << 2) | (s3 << 4) | (0x100 if n1 else 0) | (0x200 if n2 else 0) | (0x400 if n3 else 0)


def _check_wave(x, name, min_ch):
    if x.ndim != 3:
        raise ValueError('%s: batch shape must be (batch_size, num_channels, data_length)' % name)
    if not x.is_cuda:
        raise RuntimeError('%s runs on the CUDA path only; got a %s tensor' % (name, x.device))
    if x.dtype != torch.float32:
        raise TypeError('%s: float32 waveforms expected, got %s' % (name, x.dtype))
    if x.shape[1] < min_ch:
        raise ValueError('%s needs at least %d channels' % (name, min_ch))
    if x.numel() and x.stride(2) != 1:
        raise ValueError('%s works in place and needs unit stride along time' % name)


def rotate_waveforms(batch_x, codes):
    """Apply one signed channel permutation per clip, in place; returns batch_x.

    codes: B ints (list / numpy / CPU or CUDA int32 tensor) from rotation_code(), ROT_IDENTITY for
    clips to leave alone."""
    _check_wave(batch_x, 'rotate_waveforms', 4)
    B, C, L = batch_x.shape
    if not torch.is_tensor(codes):
        codes = torch.as_tensor(np.asarray(codes, dtype=np.int32))
    if codes.numel() != B:
        raise ValueError('one rotation code per clip expected: %d codes for %d clips' % (codes.numel(), B))
    if B == 0 or L == 0:
        return batch_x
    codes = codes.to(device=batch_x.device, dtype=torch.int32, non_blocking=True).contiguous()
    with _abi.device_guard(batch_x.device):
        rc = _abi.lib().seld_foa_rotate_f32(batch_x.data_ptr(), B, C, L, batch_x.stride(0), batch_x.stride(1),
                                            codes.data_ptr(), torch.cuda.current_stream(batch_x.device).cuda_stream)
    _abi.check(rc, 'seld_foa_rotate_f32')
    return batch_x


class MixOp(ctypes.Structure):            # seld_mix_op of include/seldfeat.h
    _fields_ = [('dst', ctypes.c_int32), ('src', ctypes.c_int32), ('lam', ctypes.c_float), ('flags', ctypes.c_int32)]


def wavmix_order(dst, src, lambs, batch_size):
    """seld_wavmix_order: the (dst, src, lam) pairs ordered along their chains -> (n, 4) int32 array of
    seld_mix_op records (dst, src, lam bits, flags).  Host only."""
    dst = np.ascontiguousarray(np.asarray(dst, dtype=np.int64).reshape(-1))
    src = np.ascontiguousarray(np.asarray(src, dtype=np.int64).reshape(-1))
    lam = np.ascontiguousarray(np.asarray(lambs, dtype=np.float32).reshape(-1))
    n = dst.size
    if src.size != n or lam.size != n:
        raise ValueError('dst, src and lambs must have one entry per mixed clip')
    ops = np.zeros((n, 4), dtype=np.int32)
    rc = _abi.lib().seld_wavmix_order(dst.ctypes.data, src.ctypes.data, lam.ctypes.data, n, batch_size, ops.ctypes.data)
    if rc == _abi.SELD_EINVAL:
        raise ValueError('wavmix: clip indices must lie in [0, %d) and not repeat within dst or within src' % batch_size)
    _abi.check(rc, 'seld_wavmix_order')
    return ops


def wavmix_waveforms(batch_x, dst, src, lambs):
    """wavmix.py:50 in place:  batch_x[dst] = lambs * batch_x[dst] + (1 - lambs) * batch_x[src]
    (right-hand sides taken before any assignment); returns batch_x.  lambs: (N,) tensor or array."""
    _check_wave(batch_x, 'wavmix_waveforms', 1)
    B, C, L = batch_x.shape
    if torch.is_tensor(lambs):
        lambs = lambs.detach().reshape(-1).to('cpu', torch.float32).numpy()
    ops = wavmix_order(dst, src, lambs, B)
    if len(ops) == 0 or L == 0:
        return batch_x
    ops_dev = torch.from_numpy(ops).to(batch_x.device, non_blocking=True)
    with _abi.device_guard(batch_x.device):
        rc = _abi.lib().seld_wavmix_f32(batch_x.data_ptr(), B, C, L, batch_x.stride(0), batch_x.stride(1),
                                        ops_dev.data_ptr(), len(ops),
                                        torch.cuda.current_stream(batch_x.device).cuda_stream)
    _abi.check(rc, 'seld_wavmix_f32')
    return batch_x


class Rotation:
    """Drop-in for augment.Rotation (src/augment/rotate.py:5-99).

    Same draws, in the reference's order: per clip np.random.uniform() against p, then
    random.choice over the axis permutations and np.random.choice([-1, 1], size=3) for the signs.  The
    labels are rotated clip by clip as the reference does (they are tiny); the waveforms of all drawn
    clips are rotated afterwards by ONE launch instead of a torch.stack + copy per clip."""

    def __init__(self, p, rotation_type):
        if rotation_type not in (16, 48):
            raise ValueError('rotation_type must be 16 or 48')
        self.p = p
        self.type = rotation_type
        self._table = TRANS_48 if rotation_type == 48 else TRANS_16

    def draw(self):
        """One clip's rotation: ((xx, yy, zz), (s_x, s_y, s_z), (signx, signy, signz))."""
        axes = random.choice(list(self._table.keys()))
        signs = np.random.choice([-1, 1], size=3)
        return axes, self._table[axes], signs

    @staticmethod
    def _rotate_doa(doa, axes, signs):
        xx, yy, zz = axes
        sx, sy, sz = (int(s) for s in signs)
        return torch.stack((sx * doa[..., xx], sy * doa[..., yy], sz * doa[..., zz]), dim=-1)

    def __call__(self, batch_x, batch_target):
        N = batch_x.shape[0]
        if batch_x.ndim != 3 or batch_x.shape[1] != 4:
            raise ValueError('Rotation expects FOA batches (batch_size, 4, data_length)')   # rotate.py:36 fails otherwise
        codes = np.full(N, ROT_IDENTITY, dtype=np.int32)
        for n in range(N):
            if np.random.uniform() >= self.p:
                continue
            if 'accdoa_label' in batch_target:
                key = 'accdoa_label'
                T, C = batch_target[key].shape[1:]
                doa = batch_target[key][n].reshape(T, 3, C // 3).transpose(1, 2)
            elif 'doa_label' in batch_target:
                key = 'doa_label'
                doa = batch_target[key][n]
            elif 'adpit_label' in batch_target:
                key = 'adpit_label'
                seddoa = batch_target[key][n].transpose(-1, -2)
                doa = seddoa[..., 1:]
            else:
                raise KeyError('Rotation needs accdoa_label, doa_label or adpit_label in the targets')
            axes, (s_x, s_y, s_z), signs = self.draw()
            signx, signy, signz = (int(s) for s in signs)
            codes[n] = rotation_code((s_x, s_y, s_z), (signy, signz, signx))     # rotate.py:72 / 95
            y = self._rotate_doa(doa, axes, signs)
            if key == 'accdoa_label':
                y = y.transpose(1, 2).reshape(T, -1)
            elif key == 'adpit_label':
                y = torch.cat([seddoa[..., :1], y], dim=-1).transpose(-1, -2)
            batch_target[key][n] = y
        if (codes != ROT_IDENTITY).any():
            rotate_waveforms(batch_x, codes)
        return batch_x, batch_target
