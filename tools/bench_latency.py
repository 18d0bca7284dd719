"""Small-batch latency of the front-end: eager call vs CUDA-graph replay (pseldnets_b200/graphs.py).
Wall-clock per call with a synchronize after each (what a serving loop sees)."""
import json
import sys
import time

import torch

sys.path.insert(0, '.')
import pseldnets_b200 as pb
from pseldnets_b200.graphs import GraphedFrontEnd

cfg = {'data': {'sample_rate': 24000, 'nfft': 1024, 'hoplen': 240, 'n_mels': 64, 'window': 'hann', 'audio_feature': 'logmelIV'}}
ext = pb.get_afextractor(cfg).cuda()
scalar = torch.nn.ModuleList([torch.nn.BatchNorm2d(64) for _ in range(7)]).cuda().eval()
sp = pb.ScalarParams(scalar)


def wall(fn, n=300, warm=30):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


res = {}
for B in (1, 2, 4, 16):
    x = 0.1 * torch.randn(B, 4, 240000, device='cuda')
    g = GraphedFrontEnd(ext, x.shape, scalar=sp, spec_size=256)
    res['B=%d' % B] = {'eager_us': round(wall(lambda: pb.scalar_wav2img(ext(x), sp, 256)), 1),
                       'graph_us': round(wall(lambda: g(x, clone=False)), 1)}
print(json.dumps({'what': 'waveform (B,4,240000) -> image (B,7,256,256), wall-clock per call incl. synchronize', **res}))
