"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, read-only).  It imports
/root/reference/src/utils/feature.py as-is -- the one concession is an empty stub module for
`librosa`, which feature.py imports at line 3 but which is not installable offline and is not
touched by the FOA extractors -- runs LogmelIV_Extractor / Logmel_Extractor on the seeded
inputs of oracle/synth.py (torch CPU, fp32, and the same module in fp64 as "truth"), and stores
inputs' recipe + outputs.  The reference has no golden vectors of its own.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import copy
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/src')
sys.modules.setdefault('librosa', types.ModuleType('librosa'))
import utils.feature as ref_feature  # noqa: E402  (the reference, unmodified)

from oracle import synth  # noqa: E402


def cfg(sr, hop, window='hann', feat='logmelIV'):
    return {'data': {'sample_rate': sr, 'nfft': 1024, 'hoplen': hop, 'n_mels': 64,
                     'window': window, 'audio_feature': feat}}


def run(ext, x):
    x = torch.from_numpy(x)
    with torch.no_grad():
        y32 = ext(x).contiguous().numpy()
        ext64 = copy.deepcopy(ext).double()
        y64 = ext64(x.double()).contiguous().numpy()
    return y32, y64


# name -> (kind, cfg, recipe) ; recipe is re-evaluated by tests via oracle.synth
SMALL = {
    'white_24k':      ('logmelIV', cfg(24000, 240), ('white', 11, 2, 4, 4800)),
    'uniform_24k':    ('logmelIV', cfg(24000, 240), ('uniform', 12, 1, 4, 4800)),
    'plane_24k':      ('logmelIV', cfg(24000, 240), ('plane', 13, 1, 4, 4800)),
    'zeros_24k':      ('logmelIV', cfg(24000, 240), ('zeros', 0, 1, 4, 2400)),
    'halfsilent_24k': ('logmelIV', cfg(24000, 240), ('half_silent', 14, 1, 4, 4800)),
    'ragged_24k':     ('logmelIV', cfg(24000, 240), ('white', 15, 1, 4, 5003)),
    'short_24k':      ('logmelIV', cfg(24000, 240), ('white', 16, 1, 4, 700)),
    'white_32k':      ('logmelIV', cfg(32000, 320), ('white', 17, 1, 4, 6400)),
    'quiet_24k':      ('logmelIV', cfg(24000, 240), ('quiet', 18, 1, 4, 4800)),
    'hamming_24k':    ('logmelIV', cfg(24000, 240, 'hamming'), ('white', 19, 1, 4, 2400)),
    'blackman_24k':   ('logmelIV', cfg(24000, 240, 'blackman'), ('white', 20, 1, 4, 2400)),
    'bartlett_24k':   ('logmelIV', cfg(24000, 240, 'bartlett'), ('white', 21, 1, 4, 2400)),
    'eightch_32k':    ('logmelIV', cfg(32000, 320), ('white', 22, 1, 8, 3200)),
    'mono_logmel':    ('logmel', cfg(24000, 240, feat='logmel'), ('white', 23, 2, 1, 4800)),
    'three_logmel':   ('logmel', cfg(24000, 240, feat='logmel'), ('white', 24, 1, 3, 2400)),
}


def make_input(recipe):
    kind, seed, B, C, L = recipe
    if kind == 'white':
        return synth.white(seed, B, C, L)
    if kind == 'uniform':
        return synth.uniform(seed, (B, C, L))
    if kind == 'plane':
        return synth.plane_wave_foa(seed, B, L)
    if kind == 'zeros':
        return np.zeros((B, C, L), dtype=np.float32)
    if kind == 'half_silent':
        return synth.half_silent(seed, B, C, L)
    if kind == 'quiet':
        return synth.white(seed, B, C, L, scale=3e-5)
    raise KeyError(kind)


def main():
    torch.manual_seed(0)
    out = {}
    meta = []
    for name, (kind, c, recipe) in SMALL.items():
        ext = ref_feature.LogmelIV_Extractor(c) if kind == 'logmelIV' else ref_feature.Logmel_Extractor(c)
        x = make_input(recipe)
        y32, y64 = run(ext, x)
        out[name + '/x'] = x
        out[name + '/y32'] = y32
        out[name + '/y64'] = y64.astype(np.float64)
        meta.append((name, kind, c['data']['sample_rate'], c['data']['hoplen'], c['data']['window'], recipe))
        print(name, x.shape, '->', y32.shape, 'max|y32-y64|', np.abs(y32 - y64).max())
    out['meta'] = np.array(repr(meta))
    # the two persistent buffers, as the reference builds them (sr=24000 hann)
    ext = ref_feature.LogmelIV_Extractor(cfg(24000, 240))
    sd = ext.state_dict()
    out['buffers/keys'] = np.array(repr(sorted(sd.keys())))
    out['buffers/window'] = sd['stft_extractor.window'].numpy()
    out['buffers/fb_24k'] = sd['mel_scale.fb'].numpy()
    out['buffers/fb_32k'] = ref_feature.LogmelIV_Extractor(cfg(32000, 320)).state_dict()['mel_scale.fb'].numpy()
    np.savez_compressed(os.path.join(HERE, 'foa_small.npz'), **out)

    # BASELINE cfg1 at full size: 10 s, 4 ch, 24 kHz; keep a frame subsample + digests.
    x = synth.white(1234, 1, 4, 240000)
    ext = ref_feature.LogmelIV_Extractor(cfg(24000, 240))
    y32, y64 = run(ext, x)
    frames = np.array([0, 1, 2, 3, 250, 499, 500, 501, 750, 997, 998, 999, 1000])
    np.savez_compressed(
        os.path.join(HERE, 'foa_cfg1_full.npz'),
        recipe=np.array(repr(('white', 1234, 1, 4, 240000))), frames=frames,
        y32=y32[:, :, frames], y64=y64[:, :, frames],
        shape=np.array(y32.shape), sum64=y64.sum(axis=(2, 3)), sumsq64=(y64 ** 2).sum(axis=(2, 3)),
        absmax=np.abs(y64).max(axis=(2, 3)))
    print('cfg1', y32.shape, 'max|y32-y64| logmel', np.abs(y32 - y64)[:, :4].max(), 'iv', np.abs(y32 - y64)[:, 4:].max())


if __name__ == '__main__':
    main()
