#!/usr/bin/env python
"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pseldnets_b200 as pb
def cfg(feat, sr=24000, hop=240): return {'data': {'sample_rate': sr, 'nfft': 1024, 'hoplen': hop, 'n_mels': 64, 'window': 'hann', 'audio_feature': feat}}
torch.manual_seed(0)
x = 0.1 * torch.randn(2, 4, 3000, device='cuda')
iv = pb.get_afextractor(cfg('logmelIV')).cuda(); print('iv', iv(x).abs().sum().item())
print('iv i16', iv((x * 20000).to(torch.int16)).abs().sum().item())
print('iv 8ch', iv(0.1 * torch.randn(1, 8, 3000, device='cuda')).abs().sum().item())
lm = pb.get_afextractor(cfg('logmel')).cuda(); print('lm', lm(0.1 * torch.randn(3, 3, 2900, device='cuda')).abs().sum().item())
mic = pb.get_afextractor(cfg('logmelgcc')).cuda(); print('mic', mic(x).abs().sum().item())
iv.mel_scale.fb.copy_(torch.rand(513, 64, device='cuda') * 0.01); print('dense fb (general kernel)', iv(x).abs().sum().item())
print('host', iv.forward_host(x.cpu().pin_memory(), chunk_clips=1).abs().sum().item()); torch.cuda.synchronize()
