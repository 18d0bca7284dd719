#!/usr/bin/env python
"""Developer timing: the FOA (and optionally MIC) extractor at cfg2 / cfg3 for several builds of libseldfeat.
   python tools/time_variants.py [--mic] [--rounds R] libA.so libB.so ...     (one subprocess per build and round)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import pseldnets_b200 as pb
mic = %r
cfg = {'data': {'sample_rate': 24000, 'nfft': 1024, 'hoplen': 240, 'n_mels': 64, 'window': 'hann', 'audio_feature': 'logmelgcc' if mic else 'logmelIV'}}
ext = pb.get_afextractor(cfg).cuda()
g = torch.Generator(device='cuda'); g.manual_seed(1)
x = 0.1 * torch.randn(64, 4, 240000, device='cuda', generator=g)
for _ in range(20): y = ext(x)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): y = ext(x)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 100)
print('%%.1f' %% (1e3 * best))
'''
def main():
    args = sys.argv[1:]
    mic = '--mic' in args
    if mic: args.remove('--mic')
    rounds = 1
    if '--rounds' in args:
        i = args.index('--rounds'); rounds = int(args[i + 1]); del args[i:i + 2]
    res = {a: [] for a in args}
    for r in range(rounds):
        for a in args:
            env = dict(os.environ, SELD_LIB=os.path.abspath(a))
            try:
                out = subprocess.run([sys.executable, '-c', CHILD % (ROOT, mic)], env=env, capture_output=True, text=True, timeout=300)
                res[a].append(out.stdout.strip().splitlines()[-1] if out.returncode == 0 and out.stdout.strip() else 'ERR ' + out.stderr.strip()[-200:])
            except subprocess.TimeoutExpired:
                res[a].append('TIMEOUT')
    for a in args:
        print('%-28s %s us' % (os.path.basename(a), '  '.join(res[a])), flush=True)
main()
