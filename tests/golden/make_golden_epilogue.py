"""Golden vectors for the backbone-input stage from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, read-only).  Imports
/root/reference/src/models/components/htsat.py as-is -- modules it imports but that are not
installable offline and not touched here (lightning, omegaconf, hydra ...) are replaced by empty
stubs -- and calls HTSAT_Swin_Transformer.reshape_wav2img unbound on a namespace carrying the two
attributes it reads (spec_size, freq_ratio).  The "scalar" is the loop of accdoa.py:222-227 run
on torch.nn.BatchNorm2d modules in eval mode with seeded statistics.

    python tests/golden/make_golden_epilogue.py     # rewrites tests/golden/epilogue.npz
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/src')


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        m = _Stub(self.__name__ + '.' + name)
        sys.modules[m.__name__] = m
        return m

    def __call__(self, *args, **kwargs):       # e.g. the rank_zero_only decorator: pass the function through
        return args[0] if args else None


for _ in range(32):
    try:
        import models.components.htsat as ref_htsat  # noqa: E402  (the reference, unmodified)
        break
    except ModuleNotFoundError as e:
        sys.modules[e.name] = _Stub(e.name)

from oracle import synth  # noqa: E402


def ref_scalar(x, mean, var, weight, bias, eps):
    C, M = mean.shape
    scalar = torch.nn.ModuleList([torch.nn.BatchNorm2d(M, eps=eps) for _ in range(C)]).eval()
    for c in range(C):
        scalar[c].running_mean.copy_(torch.from_numpy(mean[c]))
        scalar[c].running_var.copy_(torch.from_numpy(var[c]))
        scalar[c].weight.data.copy_(torch.from_numpy(weight[c]))
        scalar[c].bias.data.copy_(torch.from_numpy(bias[c]))
    x = torch.from_numpy(x.copy())
    with torch.no_grad():                       # accdoa.py:222-227
        x = x.transpose(1, 3)
        for nch in range(x.shape[-1]):
            x[..., [nch]] = scalar[nch](x[..., [nch]])
        x = x.transpose(1, 3)
    return x.contiguous().numpy()


def ref_wav2img(x, spec_size):
    M = x.shape[-1]
    ns = types.SimpleNamespace(spec_size=spec_size, freq_ratio=spec_size // M)   # htsat.py:417,442
    with torch.no_grad():
        return ref_htsat.HTSAT_Swin_Transformer.reshape_wav2img(ns, torch.from_numpy(x)).contiguous().numpy()


# name -> (seed, B, C, T, M, spec_size)
SMALL = {
    'pad':   (101, 2, 3, 250, 16, 64),      # T < target_T: zero padding
    'exact': (102, 1, 2, 256, 16, 64),
    'crop':  (103, 1, 2, 300, 16, 64),      # T > target_T: F.pad with a negative amount crops
    'tiny':  (104, 1, 1, 5, 8, 16),
    'r1':    (105, 1, 2, 30, 32, 32),       # freq_ratio 1: plain transpose
}
# the HTS-AT shape of the reference configs: (B, 7, 1001, 64) -> (B, 7, 256, 256); stored as digests
FULL = ('full', 106, 2, 7, 1001, 64, 256)


def main():
    out = {}
    for name, (seed, B, C, T, M, S) in SMALL.items():
        x = synth.feature_like(seed, B, C, T, M)
        p = synth.scalar_params(seed + 1000, C, M)
        xs = ref_scalar(x, *p, 1e-5)
        out[name + '/recipe'] = np.array([seed, B, C, T, M, S])
        out[name + '/scalar'] = xs
        out[name + '/img'] = ref_wav2img(x, S)
        out[name + '/scalar_img'] = ref_wav2img(xs, S)
    name, seed, B, C, T, M, S = FULL
    x = synth.feature_like(seed, B, C, T, M)
    p = synth.scalar_params(seed + 1000, C, M)
    xs = ref_scalar(x, *p, 1e-5)
    img = ref_wav2img(x, S)
    simg = ref_wav2img(xs, S)
    out['full/recipe'] = np.array([seed, B, C, T, M, S])
    out['full/img_sha256'] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), dtype=np.uint8)
    out['full/scalar_img_sub'] = simg[:, :, ::7, ::5].copy()          # strided subsample (tolerance check)
    out['full/scalar_sub'] = xs[:, :, ::11, ::3].copy()
    np.savez_compressed(os.path.join(HERE, 'epilogue.npz'), **out)
    print('wrote epilogue.npz:', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
