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
# unbalanced transform partners: frames are marked and redone by the grid the last block launches from the device
xd = x.clone(); xd[0, 1] = 0.0; xd[1, 2] *= 1e-6
print('iv redo', iv(xd).abs().sum().item(), 'mic redo', mic(xd).abs().sum().item(), 'lm redo', lm(xd[:, :3]).abs().sum().item())
iv.mel_scale.fb.copy_(torch.rand(513, 64, device='cuda') * 0.01); print('dense fb (general kernel)', iv(x).abs().sum().item())
iv = pb.get_afextractor(cfg('logmelIV')).cuda()
# backbone-input stage: partial tiles in both directions, crop, general scalar kernel
sc = torch.nn.ModuleList([torch.nn.BatchNorm2d(64) for _ in range(7)]).cuda().eval()
f = iv(x); print('wav2img', pb.scalar_wav2img(f, sc, 256).abs().sum().item(), pb.apply_scalar(f, sc).abs().sum().item())
f2 = torch.randn(2, 3, 300, 20, device='cuda'); sc2 = torch.nn.ModuleList([torch.nn.BatchNorm2d(20) for _ in range(3)]).cuda().eval()
print('wav2img odd', pb.scalar_wav2img(f2, sc2, 60).abs().sum().item(), pb.apply_scalar(f2, sc2).abs().sum().item())
# waveform augmentation: vector and scalar paths, cycles and open chains
w = 0.1 * torch.randn(6, 4, 3000, device='cuda')
pb.augment.rotate_waveforms(w, [pb.augment.ROT_IDENTITY, pb.augment.rotation_code((3, 2, 1), (-1, 1, -1))] * 3)
pb.augment.wavmix_waveforms(w, [0, 1, 2, 4], [1, 2, 0, 5], [0.3, 0.5, 0.7, 0.9])
v = w[:, :, 1:2998]; pb.augment.rotate_waveforms(v, [pb.augment.rotation_code((2, 3, 1), (1, -1, 1))] * 6); pb.augment.wavmix_waveforms(v, [0, 3], [3, 0], [0.2, 0.4])
print('augment', w.abs().sum().item())
from pseldnets_b200.graphs import GraphedFrontEnd
print('graph', GraphedFrontEnd(iv, (2, 4, 3000), scalar=sc, spec_size=256)(x).abs().sum().item())
print('host', iv.forward_host(x.cpu().pin_memory(), chunk_clips=1).abs().sum().item()); torch.cuda.synchronize()
