"""Backbone-input stage of the feature map, on the CUDA path.

Every backbone of the reference starts by pushing the extracted feature map (B, C, T, M) through a
per-channel eval-mode BatchNorm2d "scalar" (src/models/accdoa.py:222-227, 318-321;
src/models/einv2.py:106-109, 292-295) -- a Python loop of C sliced BatchNorm calls writing back in
place -- and HTS-AT then pads and folds the map into a square image (reshape_wav2img,
src/models/components/htsat.py:493-511: pad, permute, reshape, permute, reshape; three full copies).
Here each is one streaming kernel of libseldfeat.so (csrc/seld_epilogue.cu):

    apply_scalar(x, scalar)                  in place, like the reference loop
    reshape_wav2img(x, spec_size)            (B, C, T, M) -> (B, C, spec_size, spec_size)
    scalar_wav2img(x, scalar, spec_size)     both in one pass over the map

`scalar` is the reference's own container: an nn.ModuleList of C BatchNorm2d(M) modules (state-dict
keys `scalar.<c>.running_mean` ...), or a ScalarParams built from it once.  Eval mode only -- batch
statistics (training) are not part of this path and raise.  No CPU fallback.
"""
import torch

from . import _abi


class ScalarParams:
    """The (C, M) running_mean / running_var / weight / bias rows of a `scalar` ModuleList, stacked on
    one device so the kernels can index them by channel.  Build once per eval run (the statistics do
    not change in eval mode); `apply_scalar` / `scalar_wav2img` also accept the ModuleList itself and
    stack it on every call."""

    def __init__(self, scalar, device=None):
        mods = list(scalar)
        if not mods:
            raise ValueError('scalar: empty ModuleList')
        for bn in mods:
            if bn.training:
                raise RuntimeError('scalar: eval-mode BatchNorm only (batch statistics are outside this path); call .eval()')
            if bn.running_mean is None or bn.running_var is None:
                raise RuntimeError('scalar: BatchNorm without running statistics uses batch statistics')
        eps = {float(bn.eps) for bn in mods}
        if len(eps) != 1:
            raise ValueError('scalar: the BatchNorm modules disagree on eps')
        self.eps = eps.pop()
        dev = torch.device(device) if device is not None else mods[0].running_mean.device

        def rows(get):
            vals = [get(bn) for bn in mods]
            if all(v is None for v in vals):
                return None
            if any(v is None for v in vals):
                raise ValueError('scalar: mixed affine / non-affine BatchNorm modules')
            return torch.stack([v.detach().to(device=dev, dtype=torch.float32) for v in vals]).contiguous()

        self.mean = rows(lambda bn: bn.running_mean)
        self.var = rows(lambda bn: bn.running_var)
        self.weight = rows(lambda bn: bn.weight)
        self.bias = rows(lambda bn: bn.bias)
        if (self.weight is None) != (self.bias is None):
            raise ValueError('scalar: weight and bias must both be present or both absent')
        self.C, self.M = self.mean.shape

    def pointers(self):
        p = lambda t: 0 if t is None else t.data_ptr()
        return p(self.mean), p(self.var), p(self.weight), p(self.bias)


def _params(scalar, x):
    if scalar is None:
        return None
    sp = scalar if isinstance(scalar, ScalarParams) else ScalarParams(scalar, x.device)
    if sp.mean.device != x.device:
        raise RuntimeError('scalar statistics live on %s, the feature map on %s' % (sp.mean.device, x.device))
    if (sp.C, sp.M) != (x.shape[1], x.shape[3]):
        raise ValueError('scalar is for (C, M) = (%d, %d), the feature map has (%d, %d)'
                         % (sp.C, sp.M, x.shape[1], x.shape[3]))
    return sp


def _check_map(x, name):
    if x.ndim != 4:
        raise ValueError('%s: feature map shape must be (batch_size, num_channels, time_frames, mel_bins)' % name)
    if not x.is_cuda:
        raise RuntimeError('%s runs on the CUDA path only; got a %s tensor' % (name, x.device))
    if x.dtype != torch.float32:
        raise TypeError('%s: float32 feature map expected, got %s' % (name, x.dtype))


def apply_scalar(x, scalar):
    """accdoa.py:222-227 in one launch: x[:, c, :, m] <- BatchNorm2d_c(eval) applied IN PLACE; returns x.

    x: (B, C, T, M) float32 CUDA tensor (must be contiguous: the reference writes through views of the
    extractor output, which is); scalar: ModuleList of C BatchNorm2d(M) in eval mode, or ScalarParams."""
    _check_map(x, 'apply_scalar')
    if not x.is_contiguous():
        raise ValueError('apply_scalar works in place and needs a contiguous feature map')
    sp = _params(scalar, x)
    if sp is None or x.numel() == 0:
        return x
    B, C, T, M = x.shape
    with _abi.device_guard(x.device):
        rc = _abi.lib().seld_scalar_f32(x.data_ptr(), B, C, T, M, *sp.pointers(), sp.eps,
                                        torch.cuda.current_stream(x.device).cuda_stream)
    _abi.check(rc, 'seld_scalar_f32')
    return x


def scalar_wav2img(x, scalar, spec_size=256):
    """Scalar (or none) + htsat.py:493-511 in one pass: (B, C, T, M) -> new (B, C, spec_size, spec_size).

    The input is left untouched (the reference's scalar loop overwrites it; nothing downstream of
    HTS-AT's input reads it again)."""
    _check_map(x, 'scalar_wav2img')
    B, C, T, M = x.shape
    if spec_size % M != 0:
        raise ValueError('spec_size (%d) must be a multiple of mel_bins (%d) (htsat.py:442)' % (spec_size, M))
    x = x.contiguous()
    sp = _params(scalar, x)
    img = torch.empty((B, C, spec_size, spec_size), dtype=torch.float32, device=x.device)
    if B == 0:
        return img
    ptrs = sp.pointers() if sp is not None else (0, 0, 0, 0)
    with _abi.device_guard(x.device):
        rc = _abi.lib().seld_scalar_wav2img_f32(x.data_ptr(), B, C, T, M, spec_size, *ptrs,
                                                sp.eps if sp is not None else 0.0, img.data_ptr(),
                                                torch.cuda.current_stream(x.device).cuda_stream)
    _abi.check(rc, 'seld_scalar_wav2img_f32')
    return img


def reshape_wav2img(x, spec_size=256):
    """HTSAT_Swin_Transformer.reshape_wav2img (htsat.py:493-511) as one kernel; bit-identical."""
    return scalar_wav2img(x, None, spec_size)
