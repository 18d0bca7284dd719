"""Developer probe for the tensor-core mel projection (run on the B200 through gpurun):

    python tools/umma_probe.py > gpurun_out/umma_probe.txt

1. layout check: builds the shared-memory images exactly as the iv5 kernel lays them out (bf16 hi/lo
   planes of 8 frames x 8 rows x 528 bins, MN-major, no swizzle, frame stride 16896 B; mel bank as banded
   K-major chunks of 16 bins), issues the 3-term sequence hi*hi + lo*hi + hi*lo chunk by chunk into
   column windows of the M = 64 accumulator, and compares the TMEM result with numpy;
2. the same with the bank in fp16 (scaled by 32) against bf16 data (mixed a/b formats);
3. timing (clock64 per tile, 148 blocks): banded N = 16 windows vs dense N = 64, and the tensor-pipe cost
   of the two 32-point DFT stages done as 3-term GEMMs (18 MMAs of 128 x 64 x 16 per frame).
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pseldnets_b200 import filterbank  # noqa: E402

LIB = ctypes.CDLL(os.path.join(ROOT, 'tools', 'libumma_probe.so'))
LIB.umma_probe_run.restype = ctypes.c_int
LIB.umma_probe_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                               ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                               ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]

NB = 528                     # bins per frame (33 chunks of 16)
PLANE = NB * 16              # bytes per (frame, plane): 16 B per bin = 8 rows of 2 B
SLOT = 2 * PLANE             # hi plane then lo plane
NF = 8


def desc_hi(lbo, sbo):
    return ((lbo >> 4) & 0x3fff) << 16 | ((sbo >> 4) & 0x3fff) << 32 | 1 << 46


def idesc(M, N, afmt, bfmt, a_mn=1, b_mn=0):
    return (1 << 4) | (afmt << 7) | (bfmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def split16(x, dtype):
    """fp32 array -> (hi, lo) both representable in `dtype` (torch.bfloat16 / torch.float16), as fp32 values."""
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    hi = t.to(dtype).to(torch.float32)
    lo = (t - hi).to(dtype).to(torch.float32)
    return hi.numpy(), lo.numpy()


def bits16(x, dtype):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dtype).view(torch.int16).numpy()


def build_a_image(q):
    """q: (8 frames, 8 rows, NB bins) fp32 -> byte image, plus the split values."""
    hi, lo = split16(q, torch.bfloat16)
    img = np.zeros((NF, 2, NB, 8), dtype=np.int16)
    img[:, 0] = bits16(hi, torch.bfloat16).transpose(0, 2, 1)
    img[:, 1] = bits16(lo, torch.bfloat16).transpose(0, 2, 1)
    return img.reshape(-1).view(np.uint8), hi, lo


def chunk_windows(fb, n_max, align=1):
    """per chunk of 16 bins: (col0, ncols) covering the non-zero bands, ncols a multiple of 8, col0 of `align`"""
    out = []
    M = fb.shape[1]
    for c in range(NB // 16):
        rows = fb[16 * c:16 * c + 16]
        nz = np.nonzero(np.abs(rows).sum(0))[0]
        if len(nz) == 0:
            out.append((0, 8))
            continue
        lo, hi = int(nz[0]), int(nz[-1]) + 1
        lo -= lo % align
        n = max(8, -(-(hi - lo) // 8) * 8)
        n = min(n, n_max)
        if lo + n > M:
            lo = M - n
        assert lo + n >= hi and lo % align == 0, (c, lo, n, hi)
        out.append((lo, n))
    return out


def b_tile_bytes(vals, dtype):
    """vals: (16 k, n) fp32 -> K-major no-swizzle tile: byte(n,k) = (n%8)*16 + (n/8)*256 + (k%8)*2 + (k/8)*128"""
    n = vals.shape[1]
    b = bits16(vals, dtype)
    img = np.zeros((n // 8, 2, 8, 8), dtype=np.int16)      # [n/8][k/8][n%8][k%8]
    for kb in range(2):
        blk = b[8 * kb:8 * kb + 8, :]                      # (k%8, n)
        img[:, kb] = blk.T.reshape(n // 8, 8, 8)           # [n/8][n%8][k%8]
    return img.reshape(-1).view(np.uint8)


def run(a_img, b_img, ops, a_hi, b_hi, reps=1, tmem_cols=64, grid=1, out_cols=64, want_d=True):
    dev = 'cuda:0'
    a = torch.from_numpy(a_img.copy()).to(dev)
    b = torch.from_numpy(b_img.copy()).to(dev)
    o = torch.from_numpy(np.asarray(ops, dtype=np.uint32).reshape(-1).copy()).to(dev)
    d = torch.zeros(128, out_cols, dtype=torch.float32, device=dev)
    cyc = torch.zeros(grid, dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    rc = LIB.umma_probe_run(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), o.data_ptr(), len(ops),
                            a_hi, b_hi, reps, tmem_cols, grid, d.data_ptr() if want_d else None, out_cols, cyc.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0, 'launch failed: cudaError %d' % rc
    return d.cpu().numpy(), cyc.cpu().numpy(), e0.elapsed_time(e1)


def mel_case(bdtype, bscale, label, align=1, dense_chunks=0):
    rng = np.random.default_rng(5)
    fb = filterbank.melscale_fbanks_htk_slaney(513, 20.0, 12000.0, 64, 24000).numpy()
    fbp = np.zeros((NB, 64), np.float32)
    fbp[:513] = fb
    q = np.zeros((NF, 8, NB), np.float32)
    q[:, :4, :513] = (rng.standard_normal((NF, 4, 513)) ** 2 * 10.0 ** rng.uniform(-6, 3, (NF, 4, 1))).astype(np.float32)
    q[:, 4:7, :513] = rng.uniform(-1, 1, (NF, 3, 513)).astype(np.float32)
    a_img, ahi, alo = build_a_image(q)
    bhi, blo = split16(fbp * bscale, bdtype)
    wins = chunk_windows(fbp, 64, align)
    if dense_chunks:                                       # plain GEMM check: N = 64 at column 0, first chunks only
        q[:, :, 16 * dense_chunks:] = 0
        a_img, ahi, alo = build_a_image(q)
        wins = [(0, 64)] * dense_chunks
    bfmt = 1 if bdtype == torch.bfloat16 else 0
    tiles, ops = [], []
    off = 0

    def add_tile(vals):
        nonlocal off
        t = b_tile_bytes(vals, bdtype)
        tiles.append(t)
        o = off
        off += t.size
        return o // 16

    # first op: chunk 0, hi*hi over all 64 columns with D overwritten (zero-initialises the accumulator)
    first = add_tile(bhi[0:16, :])
    ops.append((0, first, 0, idesc(64, 64, 1, bfmt)))
    for c, (c0, n) in enumerate(wins):
        o_hi = add_tile(bhi[16 * c:16 * c + 16, c0:c0 + n])
        o_lo = add_tile(blo[16 * c:16 * c + 16, c0:c0 + n])
        a_h, a_l = (c * 256) // 16, (PLANE + c * 256) // 16
        idn = idesc(64, n, 1, bfmt)
        if c > 0:
            ops.append((a_h, o_hi, c0 | 1 << 31, idn))
        ops.append((a_l, o_hi, c0 | 1 << 31, idn))
        ops.append((a_h, o_lo, c0 | 1 << 31, idn))
    b_img = np.concatenate(tiles)
    d, cyc, _ = run(a_img, b_img, ops, desc_hi(128, SLOT), desc_hi(128, 256))
    rows = np.arange(64)
    got = d[(rows % 16) + 32 * (rows // 16)].reshape(NF, 8, 64) / bscale
    a64, al64 = ahi.astype(np.float64), alo.astype(np.float64)
    want = (np.einsum('fqk,km->fqm', a64 + al64, (bhi + blo).astype(np.float64))
            - np.einsum('fqk,km->fqm', al64, blo.astype(np.float64))) / bscale
    true = np.einsum('fqk,km->fqm', q.astype(np.float64), fbp.astype(np.float64))
    e_split = np.abs(got - want).max(axis=(0, 2)) / np.abs(want).max(axis=(0, 2)).clip(1e-30)
    e_true = np.abs(got - true).max(axis=(0, 2)) / np.abs(true).max(axis=(0, 2)).clip(1e-30)
    rel_true = (np.abs(got - true) / np.abs(true).clip(1e-30))[:, :4].max()
    print('[%s] %d MMAs, B image %d B, windows N: %s' % (label, len(ops), b_img.size, sorted(set(n for _, n in wins))))
    print('[%s] per-row max err / row max: vs 3-term model %s' % (label, np.array2string(e_split[:7], precision=2)))
    print('[%s]                              vs fp64 truth   %s' % (label, np.array2string(e_true[:7], precision=2)))
    print('[%s] power rows: max element-wise relative error vs truth %.3e (dB error %.2e)' % (label, rel_true, 4.343 * rel_true))
    print('[%s] tile cycles (1 block, incl. issue + commit + wait): %d' % (label, cyc[0]))
    return a_img, b_img, ops, wins


def timing(a_img, b_img, ops, wins):
    # banded sequence as the kernel would issue it, 148 blocks
    _, cyc, ms = run(a_img, b_img, ops, desc_hi(128, SLOT), desc_hi(128, 256), reps=400, grid=148, want_d=False)
    print('[time] banded 3-term mel (%d MMAs/tile): %.0f cycles/tile median over 148 blocks, kernel %.3f ms for 400 tiles/block'
          % (len(ops), np.median(cyc) / 400, ms))
    # dense N = 64 for every chunk (timing only: reuses the first tile as B)
    dense = [(0, ops[0][1], 0, idesc(64, 64, 1, 1))]
    for c in range(NB // 16):
        for t in range(3):
            if c == 0 and t == 0:
                continue
            dense.append(((c * 256) // 16 if t != 1 else (PLANE + c * 256) // 16, ops[0][1], 0 | 1 << 31, idesc(64, 64, 1, 1)))
    _, cyc, ms = run(a_img, b_img, dense, desc_hi(128, SLOT), desc_hi(128, 256), reps=400, grid=148, want_d=False)
    print('[time] dense N=64 3-term mel (%d MMAs/tile): %.0f cycles/tile, kernel %.3f ms' % (len(dense), np.median(cyc) / 400, ms))
    # single MMA cost by N (M = 64, K = 16), back to back on the same accumulator
    for n in (8, 16, 32, 64, 128):
        seq = [(0, ops[0][1], 0 | (1 << 31 if i else 0), idesc(64, n, 1, 1)) for i in range(64)]
        _, cyc, _ = run(a_img, b_img, seq, desc_hi(128, SLOT), desc_hi(128, 256), reps=100, grid=148, want_d=False, tmem_cols=128,
                        out_cols=64)
        print('[time] M=64 N=%3d K=16: %.1f cycles per MMA (64 back to back, same accumulator)' % (n, np.median(cyc) / 100 / 64))
    for n in (8, 64):                                     # ... and on 4 independent accumulators (columns 0, 64, 128, 192)
        seq = [(0, ops[0][1], (64 * (i % 4)) | (1 << 31 if i >= 4 else 0), idesc(64, n, 1, 1)) for i in range(64)]
        _, cyc, _ = run(a_img, b_img, seq, desc_hi(128, SLOT), desc_hi(128, 256), reps=100, grid=148, want_d=False, tmem_cols=256,
                        out_cols=64)
        print('[time] M=64 N=%3d K=16: %.1f cycles per MMA (64 back to back, 4 independent accumulators)' % (n, np.median(cyc) / 100 / 64))
    for n in (64, 128, 256):
        seq = [(0, ops[0][1], 0 | (1 << 31 if i else 0), idesc(128, n, 1, 1)) for i in range(64)]
        _, cyc, _ = run(a_img, b_img, seq, desc_hi(128, 512), desc_hi(128, 256), reps=100, grid=148, want_d=False, tmem_cols=256,
                        out_cols=64)
        print('[time] M=128 N=%3d K=16: %.1f cycles per MMA (64 back to back, same accumulator)' % (n, np.median(cyc) / 100 / 64))
    # DFT-as-GEMM: per frame 18 MMAs of M=128, N=64, K=16 (two 32-point stages, 3-term split); 433 frames per SM at cfg2
    seq = [((i % 8) * 16, ops[0][1], 0 | (1 << 31 if i % 6 else 0), idesc(128, 64, 1, 1)) for i in range(18)]
    _, cyc, ms = run(a_img, b_img, seq, desc_hi(128, 512), desc_hi(128, 256), reps=433, grid=148, want_d=False)
    print('[time] DFT stages as 3-term bf16 GEMMs: 18 x (128x64x16) per frame, 433 frames per block, 148 blocks: '
          '%.0f cycles/frame, kernel %.3f ms (tensor pipe only: no operand conversion, no twiddles, no TMEM read-back)'
          % (np.median(cyc) / 433, ms))
    seq1 = [((i % 8) * 16, ops[0][1], 0 | (1 << 31 if i % 6 else 0), idesc(128, 64, 1, 1)) for i in range(6)]
    _, cyc, ms = run(a_img, b_img, seq1, desc_hi(128, 512), desc_hi(128, 256), reps=433, grid=148, want_d=False)
    print('[time] ... single-pass (no split, fails the IV tolerance): 6 MMAs per frame: %.0f cycles/frame, kernel %.3f ms'
          % (np.median(cyc) / 433, ms))


if __name__ == '__main__':
    case = sys.argv[1] if len(sys.argv) > 1 else 'band1'
    print('== case', case, '|', torch.cuda.get_device_name(0))
    if case == 'dense':
        mel_case(torch.bfloat16, 1.0, 'dense 8 chunks', dense_chunks=8)
    elif case.startswith('band'):
        mel_case(torch.bfloat16, 1.0, 'bf16 x bf16, col0 %% %s' % case[4:], align=int(case[4:]))
    elif case == 'fp16':
        mel_case(torch.float16, 32.0, 'bf16 x fp16(x32)', align=int(sys.argv[2]))
    elif case == 'timing':
        a_img, b_img, ops, wins = mel_case(torch.bfloat16, 1.0, 'bf16 x bf16', align=int(sys.argv[2]))
        timing(a_img, b_img, ops, wins)
