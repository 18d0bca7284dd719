#!/usr/bin/env python
"""Developer timing: end-to-end host pipeline with int16 PCM input vs float32 input (cfg2 shapes)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pseldnets_b200 as pb
cfg = {'data': {'sample_rate': 24000, 'nfft': 1024, 'hoplen': 240, 'n_mels': 64, 'window': 'hann', 'audio_feature': 'logmelIV'}}
ext = pb.get_afextractor(cfg).cuda()
pcm = torch.randint(-3000, 3000, (64, 4, 240000), dtype=torch.int32).to(torch.int16).pin_memory()
xf = (pcm.float() / 32768.0).pin_memory()
out = torch.empty((64, 7, 1001, 64), dtype=torch.float32).pin_memory()
for name, x in (('float32', xf), ('int16', pcm)):
    for _ in range(3): ext.forward_host(x, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ext.forward_host(x, out=out, synchronize=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('%s host input: %.2f ms/step -> %.0f audio-s/s end to end' % (name, ms, 640 / ms * 1e3))
xd = pcm.cuda()
for _ in range(5): ext(xd)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): ext(xd)
e1.record(); torch.cuda.synchronize()
print('int16 resident: %.3f ms/step' % (e0.elapsed_time(e1) / 50))
