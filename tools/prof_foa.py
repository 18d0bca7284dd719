#!/usr/bin/env python
"""Developer: a few calls of the FOA extractor at cfg2 (for ncu captures):  ncu ... python tools/prof_foa.py [mic]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pseldnets_b200 as pb
mic = len(sys.argv) > 1 and sys.argv[1] == 'mic'
cfg = {'data': {'sample_rate': 24000, 'nfft': 1024, 'hoplen': 240, 'n_mels': 64, 'window': 'hann', 'audio_feature': 'logmelgcc' if mic else 'logmelIV'}}
ext = pb.get_afextractor(cfg).cuda()
g = torch.Generator(device='cuda'); g.manual_seed(1)
x = 0.1 * torch.randn(64, 4, 240000, device='cuda', generator=g)
for _ in range(8): y = ext(x)
torch.cuda.synchronize()
