"""CPU: host-side logic of the drop-in (constructor, buffers, errors, factory), the C-ABI
library's exported symbols, and the host model of the warp FFT index algebra."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, make_cfg
import pseldnets_b200 as pb
from pseldnets_b200 import _abi, filterbank as fbk


def test_buffers_match_reference(golden_small):
    g, _ = golden_small
    ext = pb.LogmelIV_Extractor(make_cfg())
    sd = ext.state_dict()
    assert repr(sorted(sd.keys())) == str(g['buffers/keys'])
    assert np.array_equal(sd['stft_extractor.window'].numpy(), g['buffers/window'])
    assert np.array_equal(sd['mel_scale.fb'].numpy(), g['buffers/fb_24k'])
    ext32 = pb.Logmel_Extractor(make_cfg(32000, 320, feat='logmel'))
    assert np.array_equal(ext32.mel_scale.fb.numpy(), g['buffers/fb_32k'])
    assert len(list(ext.parameters())) == 0


def test_state_dict_roundtrip_strict():
    a = pb.LogmelIV_Extractor(make_cfg())
    b = pb.LogmelIV_Extractor(make_cfg(window='hamming'))
    b.load_state_dict(a.state_dict(), strict=True)
    assert torch.equal(b.stft_extractor.window, a.stft_extractor.window)


def test_ctor_and_forward_errors():
    with pytest.raises(AssertionError):
        pb.LogmelIV_Extractor(make_cfg(window='kaiser'))
    ext = pb.LogmelIV_Extractor(make_cfg())
    with pytest.raises(ValueError):
        ext(torch.zeros(4, 2400))
    with pytest.raises(RuntimeError):          # no CPU path: fail loudly
        ext(torch.zeros(1, 4, 2400))


def test_factory():
    assert isinstance(pb.get_afextractor(make_cfg(feat='logmelIV')), pb.LogmelIV_Extractor)
    assert isinstance(pb.get_afextractor(make_cfg(feat='logmel')), pb.Logmel_Extractor)
    assert pb.get_afextractor(make_cfg(feat='salsalite')) is None


def test_windows_and_banks():
    for name in ('hann', 'hamming', 'blackman', 'bartlett'):
        w = fbk.make_window(name, 1024)
        assert w.shape == (1024,) and w.dtype == torch.float32
    fb = fbk.melscale_fbanks_htk_slaney(513, 20, 12000, 64, 24000)
    assert fb.shape == (513, 64) and int((fb != 0).sum()) == 998
    assert int(((fb != 0).sum(dim=1)).max()) <= 2             # every bin feeds <= 2 bands
    bank = fbk.librosa_mel_bank(24000, 1024, 64)
    assert bank.shape == (513, 64) and bank.dtype == torch.float32
    try:
        import torchaudio
    except Exception:
        return
    ref = torchaudio.functional.melscale_fbanks(513, 20, 12000, 64, 24000, norm='slaney', mel_scale='htk')
    assert torch.equal(fb, ref)
    ref2 = torchaudio.functional.melscale_fbanks(513, 0, 12000, 64, 24000, norm='slaney', mel_scale='slaney')
    assert (bank - ref2).abs().max() < 1e-7


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'seldfeat.h')).read()
    declared = set(re.findall(r'\b(seld_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), 'libseldfeat.so does not export %s' % name
    assert declared == set(_abi.SIGNATURES), declared ^ set(_abi.SIGNATURES)
    assert _abi.lib().seld_version().decode().endswith('sm_100a')
    assert _abi.lib().seld_strerror(_abi.SELD_ESHORT).decode().startswith('clip too short')


def test_abi_argument_checks_without_gpu():
    """Entry points validate before touching CUDA: callable on a CPU-only host."""
    l = _abi.lib()
    assert l.seld_logmel_iv_f32(None, None, 1, 4, 2400, 9600, 2400, None, None) == _abi.SELD_EINVAL
    h = ctypes.c_void_p()
    w = np.ones(1024, np.float32)
    fb = np.ones((513, 64), np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    assert l.seld_plan_create(ctypes.byref(h), 0, w.ctypes.data_as(fp), fb.ctypes.data_as(fp),
                              512, 240, 64, 1e-10, 1e-7) == _abi.SELD_EUNSUPPORTED
    assert l.seld_plan_create(ctypes.byref(h), 0, w.ctypes.data_as(fp), fb.ctypes.data_as(fp),
                              1024, 0, 64, 1e-10, 1e-7) == _abi.SELD_EINVAL


def test_host_model_of_warp_fft(tmp_path):
    """The kernel's 32x32 index algebra (same fft32.cuh, compiled for the host) equals rfft."""
    so_path = str(tmp_path / 'fft_model.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-o', so_path,
                           os.path.join(ROOT, 'tests', 'host', 'fft_model.cpp')])
    lib = ctypes.CDLL(so_path)
    fp = ctypes.POINTER(ctypes.c_float)
    rng = np.random.default_rng(0)
    r = rng.standard_normal(32).astype(np.float32)
    i = rng.standard_normal(32).astype(np.float32)
    ref = np.fft.fft(r.astype(np.float64) + 1j * i)
    lib.model_fft32(r.ctypes.data_as(fp), i.ctypes.data_as(fp))
    assert np.abs((r + 1j * i) - ref).max() < 1e-6 * np.abs(ref).max()
    a = rng.standard_normal(1024).astype(np.float32)
    b = rng.standard_normal(1024).astype(np.float32)
    A = np.zeros(1026, np.float32)
    Bv = np.zeros(1026, np.float32)
    lib.model_fft1024_pair(a.ctypes.data_as(fp), b.ctypes.data_as(fp), A.ctypes.data_as(fp), Bv.ctypes.data_as(fp))
    Ar, Br = np.fft.rfft(2 * a.astype(np.float64)), np.fft.rfft(2 * b.astype(np.float64))
    assert np.abs((A[0::2] + 1j * A[1::2]) - Ar).max() < 1e-6 * np.abs(Ar).max()
    assert np.abs((Bv[0::2] + 1j * Bv[1::2]) - Br).max() < 1e-6 * np.abs(Br).max()


def test_segment_index_matches_reference_cases():
    """inference.segment_index restates src/utils/data_utilities.py:6-64: short clip, exact fit,
    long remainder (zero-padded last chunk), short remainder (last chunk shifted back)."""
    from pseldnets_b200 import inference as inf
    assert inf.segment_index(5, 10, 5) == ([(0, 5)], [(0, 5)])
    assert inf.segment_index(20, 10, 5) == ([(0, 10), (5, 15), (10, 20)], [(0, 0)] * 3)
    assert inf.segment_index(22, 10, 5) == ([(0, 10), (5, 15), (10, 20), (15, 22)], [(0, 0)] * 3 + [(0, 3)])
    assert inf.segment_index(23, 10, 10) == ([(0, 10), (10, 20), (13, 23)], [(0, 0)] * 3)
    assert inf.segment_index(26, 10, 10) == ([(0, 10), (10, 20), (20, 26)], [(0, 0)] * 2 + [(0, 4)])
    assert inf.segment_index(23, 10, 10, True)[0][-1] == (20, 23)
    try:                                   # same answers as the reference itself where it is mounted
        import sys
        sys.path.insert(0, '/root/reference/src')
        import numpy as np
        from utils.data_utilities import segment_index as ref_si
    except Exception:
        return
    for x_len, ch, hp in ((5, 10, 5), (20, 10, 5), (22, 10, 5), (23, 10, 10), (26, 10, 10), (415200, 240000, 12000)):
        for flag in (False, True):
            assert tuple(map(list, ref_si(np.zeros((1, x_len)), ch, hp, flag))) == tuple(map(list, inf.segment_index(x_len, ch, hp, flag)))


def test_scalar_params_host_logic():
    """ScalarParams stacks the reference's `scalar` ModuleList; everything that is not eval-mode
    BatchNorm with running statistics is refused, and the compute calls refuse CPU tensors."""
    C, M = 3, 8
    scalar = torch.nn.ModuleList([torch.nn.BatchNorm2d(M) for _ in range(C)])
    for c, bn in enumerate(scalar):
        bn.running_mean.fill_(float(c))
        bn.weight.data.fill_(2.0 + c)
    with pytest.raises(RuntimeError):
        pb.ScalarParams(scalar)                              # still in training mode
    sp = pb.ScalarParams(scalar.eval())
    assert (sp.C, sp.M) == (C, M) and sp.eps == 1e-5
    assert torch.equal(sp.mean[:, 0], torch.arange(3.0)) and torch.equal(sp.weight[:, 0], torch.tensor([2.0, 3.0, 4.0]))
    assert sp.var.shape == sp.bias.shape == (C, M)
    na = pb.ScalarParams(torch.nn.ModuleList([torch.nn.BatchNorm2d(M, affine=False)]).eval())
    assert na.weight is None and na.bias is None and na.pointers()[2:] == (0, 0)
    with pytest.raises(RuntimeError):
        pb.ScalarParams(torch.nn.ModuleList([torch.nn.BatchNorm2d(M, track_running_stats=False)]).eval())
    with pytest.raises(ValueError):
        pb.ScalarParams(torch.nn.ModuleList([torch.nn.BatchNorm2d(M, eps=1e-3), torch.nn.BatchNorm2d(M)]).eval())
    x = torch.zeros(1, C, 5, M)
    with pytest.raises(RuntimeError):
        pb.apply_scalar(x, sp)                               # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        pb.reshape_wav2img(x, 8)
    with pytest.raises(ValueError):
        pb.reshape_wav2img(x[0], 8)


def test_epilogue_abi_argument_checks_without_gpu():
    l = _abi.lib()
    assert l.seld_scalar_f32(None, 1, 7, 100, 64, None, None, None, None, 1e-5, None) == _abi.SELD_OK   # no scalar: no-op
    assert l.seld_scalar_f32(None, 1, 0, 100, 64, None, None, None, None, 1e-5, None) == _abi.SELD_EINVAL
    assert l.seld_scalar_f32(None, 1, 7, 100, 62, None, None, None, None, 1e-5, None) == _abi.SELD_EUNSUPPORTED
    assert l.seld_scalar_f32(None, 1, 7, 100, 64, 16, None, None, None, 1e-5, None) == _abi.SELD_EINVAL  # mean without var
    assert l.seld_scalar_wav2img_f32(None, 0, 7, 1001, 64, 256, None, None, None, None, 0.0, None, None) == _abi.SELD_OK
    assert l.seld_scalar_wav2img_f32(None, 1, 7, 1001, 64, 250, None, None, None, None, 0.0, None, None) == _abi.SELD_EINVAL
    assert l.seld_scalar_wav2img_f32(None, 1, 7, 1001, 64, 256, None, None, None, None, 0.0, None, None) == _abi.SELD_EINVAL
    assert l.seld_scalar_wav2img_f32(None, 1, 7, 1001, 6, 6, None, None, None, None, 0.0, None, None) == _abi.SELD_EUNSUPPORTED


def test_wavmix_order_host_logic():
    """seld_wavmix_order (host only): every pair is emitted once, and walking the list the way the kernel
    does reproduces the gather-then-scatter semantics of wavmix.py:50 on a CPU model."""
    import pseldnets_b200.augment as aug
    rng = np.random.default_rng(0)
    BEGIN, USE_HEAD = 1, 2
    for trial in range(200):
        B = int(rng.integers(1, 12))
        n = int(rng.integers(0, B + 1))
        dst = rng.permutation(B)[:n]
        src = rng.permutation(B)[:n] if trial % 2 else rng.permutation(dst)
        lam = rng.random(n).astype(np.float32)
        ops = aug.wavmix_order(dst, src, lam, B)
        assert sorted(zip(ops[:, 0], ops[:, 1])) == sorted(zip(dst, src))
        x = rng.standard_normal(B).astype(np.float32)
        want = x.copy()
        want[dst] = lam * x[dst] + (np.float32(1) - lam) * x[src]
        got, cur, head = x.copy(), None, None
        for d, s, lbits, fl in ops:
            l = np.array([lbits], np.int32).view(np.float32)[0]
            if fl & BEGIN:
                cur = head = got[d]
            else:
                assert d == prev_src                          # a chain continues at the previous source
            nxt = head if fl & USE_HEAD else got[s]
            got[d] = l * cur + (np.float32(1) - l) * nxt
            cur, prev_src = nxt, s
        assert np.array_equal(got, want), (dst, src, ops)
    with pytest.raises(ValueError):
        aug.wavmix_order([0, 0], [1, 2], [0.5, 0.5], 4)
    with pytest.raises(ValueError):
        aug.wavmix_order([0, 1], [2, 2], [0.5, 0.5], 4)
    with pytest.raises(ValueError):
        aug.wavmix_order([0], [4], [0.5], 4)
    assert aug.rotation_code((1, 2, 3), (1, 1, 1)) == 1 | (2 << 2) | (3 << 4)
    assert aug.rotation_code((3, 2, 1), (-1, 1, -1)) == (3 | (2 << 2) | (1 << 4) | 0x100 | 0x400)


def test_augment_abi_argument_checks_without_gpu():
    l = _abi.lib()
    assert l.seld_foa_rotate_f32(None, 0, 4, 100, 400, 100, None, None) == _abi.SELD_OK
    assert l.seld_foa_rotate_f32(None, 2, 3, 100, 300, 100, None, None) == _abi.SELD_EINVAL     # needs channels 0..3
    assert l.seld_foa_rotate_f32(None, 2, 4, 100, 400, 100, None, None) == _abi.SELD_EINVAL     # null pointers
    assert l.seld_wavmix_f32(None, 2, 4, 100, 400, 100, None, 0, None) == _abi.SELD_OK
    assert l.seld_wavmix_f32(None, 2, 4, 100, 400, 100, None, 1, None) == _abi.SELD_EINVAL
    assert l.seld_wavmix_order(None, None, None, 0, 4, None) == _abi.SELD_OK
    assert l.seld_wavmix_order(None, None, None, 1, 4, None) == _abi.SELD_EINVAL


def test_bench_reference_arm_prints_one_json_line():
    """The driver contract: `bench.py --impl reference` prints exactly ONE JSON line on stdout with the agreed keys
    (everything else a library may print goes to stderr)."""
    import json
    import sys
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                          '--batch', '4'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['unit'] == 'audio-s/s' and d['higher_is_better'] is True
    # 'reference' = the unmodified feature.py vendored into baseline/_ref by __graft_entry__.build(); 'port' without it
    vendored = os.path.exists(os.path.join(ROOT, 'baseline', '_ref', 'utils', 'feature.py'))
    assert d['cpu_baseline']['kind'] == ('reference' if vendored else 'port')
    assert d['cpu_baseline']['cores'] >= 1 and d['value'] > 0
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and 'workload' in d['config']


def test_paired_row_layout_of_the_iv2_kernel():
    """Index algebra of the pair-rows in csrc/seld_foa_iv2.cu (restated here): the untangle step, which holds bin
    lane + 32*kb, and the mel walk, which wants bins 16*lane .. 16*lane + 15, must address the same float2 for the
    same bin, injectively, and both access patterns must be free of shared-memory bank conflicts."""
    K_PAIR_WORDS = 1056

    def writer_word(lane, kb):                      # float2 of bin lane + 32*kb: word 64*kb + wofs[kb & 3]
        x = kb & 3
        wofs = 32 * (lane >> 4) + 4 * ((((lane & 15) >> 1) ^ ((2 * x + (lane >> 4)) & 7))) + 2 * (lane & 1)
        return 64 * kb + wofs

    def reader_word(lane, j):                       # bin 16*lane + j: sub-chunk i = j >> 1 at 32*lane + 4*(i ^ (lane & 7))
        return 32 * lane + 4 * ((j >> 1) ^ (lane & 7)) + 2 * (j & 1)

    where = {}
    for kb in range(17):
        for lane in range(32 if kb < 16 else 1):
            k = lane + 32 * kb
            w = writer_word(lane, kb)
            assert 0 <= w and w + 1 < K_PAIR_WORDS
            where[k] = w
    assert len(set(where.values())) == 513                                     # injective
    for k in range(512):
        assert where[k] == reader_word(k >> 4, k & 15), k
    assert where[512] == 1024                                                   # lane 31 reads it at row + 1024
    # STS.64 of the untangle step: each half-warp's 16 stores cover 32 distinct banks (2 wavefronts per store)
    for kb in range(16):
        for half in range(2):
            banks = set()
            for lane in range(16 * half, 16 * half + 16):
                w = writer_word(lane, kb)
                banks.update({w % 32, (w + 1) % 32})
            assert len(banks) == 32
    # LDS.128 of the walk: the 8 lanes of a phase hit 8 different 16-byte bank groups
    for i in range(8):
        for phase in range(4):
            groups = {((32 * lane + 4 * (i ^ (lane & 7))) % 32) // 4 for lane in range(8 * phase, 8 * phase + 8)}
            assert len(groups) == 8


def test_extractor_survives_copy_and_pickle_with_live_plans():
    """The reference extractor is a plain nn.Module: copy.deepcopy / pickle / torch.save of a model that owns it
    work at any time (feature.py:20-37; tests/golden/make_golden.py deep-copies it).  The drop-in caches device
    plans holding raw ctypes handles after its first forward; those must stay out of the copied state, and two
    copies must never share (and double-free) one handle."""
    import copy
    import io
    import pickle

    class FakePlan:                      # stands in for feature._Plan after a forward: an unpicklable ctypes handle
        def __init__(self):
            self.handle = ctypes.c_void_p(0xdead0000)

    for cls, feat in ((pb.LogmelIV_Extractor, 'logmelIV'), (pb.Logmel_Extractor, 'logmel'),
                      (pb.LogmelGCC_Extractor, 'logmelgcc')):
        ext = cls(make_cfg(feat=feat))
        ext._plans[0] = (('key',), FakePlan())
        with pytest.raises(Exception):
            pickle.dumps(ext._plans)                       # the cache itself is not picklable: that was the bug
        for clone in (copy.deepcopy(ext), copy.copy(ext), pickle.loads(pickle.dumps(ext))):
            assert clone._plans == {} and clone._plans is not ext._plans
            assert torch.equal(clone.stft_extractor.window, ext.stft_extractor.window)
            assert torch.equal(clone.mel_scale.fb, ext.mel_scale.fb)
            assert clone.n_fft == ext.n_fft and clone.hop == ext.hop
        assert len(ext._plans) == 1                        # the original keeps its plan
        holder = torch.nn.Sequential(ext)                  # as a sub-module of a model (af_extractor)
        buf = io.BytesIO()
        torch.save(holder, buf)
        buf.seek(0)
        back = torch.load(buf, weights_only=False)
        assert back[0]._plans == {}
        assert sorted(back.state_dict().keys()) == sorted(holder.state_dict().keys())
        ext._plans.clear()
