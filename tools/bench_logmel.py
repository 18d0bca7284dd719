#!/usr/bin/env python
"""Developer timing of Logmel_Extractor (log-mel only) for a few channel counts."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pseldnets_b200 as pb
cfg = {'data': {'sample_rate': 24000, 'nfft': 1024, 'hoplen': 240, 'n_mels': 64, 'window': 'hann', 'audio_feature': 'logmel'}}
ext = pb.get_afextractor(cfg).cuda()
for C, B in ((1, 256), (2, 128), (4, 64), (8, 32)):
    x = 0.1 * torch.randn(B, C, 240000, device='cuda')
    for _ in range(5): ext(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): ext(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print('C=%d B=%d: %.3f ms  -> %.2f M channel-seconds/s' % (C, B, ms, B * C * 10 / ms / 1e3))
