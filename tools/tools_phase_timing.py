#!/usr/bin/env python
"""Developer tool: build libseldfeat_timing.so with -DSELD_PHASE_TIMING and print where a warp of the
iv2 kernel spends its cycles (clock64 per phase, averaged per frame).  Run on the GPU box."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the timing build is made where nvcc is (tools/build_variant.sh out.so seld_foa_iv2.cu -DSELD_PHASE_TIMING) and passed in SELD_LIB
from pseldnets_b200 import _abi
assert 'SELD_LIB' in os.environ, 'SELD_LIB=<timing build> python tools/tools_phase_timing.py'
import torch
import pseldnets_b200 as pb
cfg = {'data': {'sample_rate': 24000, 'nfft': 1024, 'hoplen': 240, 'n_mels': 64, 'window': 'hann', 'audio_feature': 'logmelIV'}}
ext = pb.get_afextractor(cfg).cuda()
x = 0.1 * torch.randn(64, 4, 240000, device='cuda')
for _ in range(3):
    ext(x)
lib = _abi.lib()
lib.seld_dev_phase_cycles.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
buf = (ctypes.c_ulonglong * 16)()
lib.seld_dev_phase_cycles(buf, 1)
N = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    ext(x)
e1.record()
torch.cuda.synchronize()
lib.seld_dev_phase_cycles(buf, 0)
frames = 64 * 1001 * N
names = ['loop/index', 'issue loads', 'loads land + window', 'fft32 #1', 'twiddle + exchange', 'fft32 #2', 'pointwise -> rows',
         'mel walk (accumulate)', 'mel combine + store', 'mel walk (load rows)']
tot = sum(buf[i] for i in range(10))
print('kernel %.1f us (timing build)' % (1e3 * e0.elapsed_time(e1) / N))
for i in range(10):
    print('%-22s %8.0f cycles/frame  %5.1f%%' % (names[i], buf[i] / frames, 100.0 * buf[i] / tot))
print('%-22s %8.0f cycles/frame' % ('total per warp-frame', tot / frames))
