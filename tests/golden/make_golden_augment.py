"""Golden vectors for the waveform-domain augmentation from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, read-only).  Imports
/root/reference/src/augment as-is (torch / numpy only) and runs augment.Rotation and augment.WavMix
on CPU with seeded generators on the deterministic inputs of oracle/synth.py.  For WavMix the draws are
replayed after re-seeding (same calls in the same order as wavmix.py:22-40) to record which clips were
mixed with which weights; the script asserts that wavmix.py:50 applied to those reproduces the
reference's output before storing them.

    python tests/golden/make_golden_augment.py     # rewrites tests/golden/augment.npz
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/src')
import augment as ref_augment  # noqa: E402  (the reference, unmodified)

from oracle import synth  # noqa: E402


def seed_all(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


def targets(kind, seed, B, T=6, K=5):
    if kind == 'accdoa_label':
        return {kind: torch.from_numpy(synth.uniform(seed, (B, T, 3 * K)))}
    if kind == 'doa_label':
        return {kind: torch.from_numpy(synth.uniform(seed, (B, T, 2, 3)))}
    return {kind: torch.from_numpy(synth.uniform(seed, (B, T, 6, 4, K)))}          # adpit_label


# name -> (rotation_type, p, label kind, seed, B, C, L)
ROT = {
    'rot48_accdoa': (48, 0.8, 'accdoa_label', 201, 8, 4, 1000),
    'rot16_doa':    (16, 0.6, 'doa_label', 202, 6, 4, 403),       # odd length: scalar path
    'rot48_adpit':  (48, 1.0, 'adpit_label', 203, 5, 4, 64),
}
# name -> (alpha, seed, ov list, C, L)
MIX = {
    'mix_a': (0.5, 301, ['1', '2', '1', '1', '2', '3', '1', '2'], 4, 1000),
    'mix_b': (0.5, 305, ['1', '2', '1', '1', '2', '3', '1', '2'], 4, 1000),
    'mix_c': (0.5, 303, ['1', '1', '1', '1', '1', '1'], 4, 257),            # add_ov '1': a 5-cycle and a fixed point
    'mix_d': (0.5, 316, ['1', '2', '2', '1', '3', '1', '1'], 2, 64),       # add_ov '1': a 2-cycle and two fixed points
}


def main():
    out = {}
    for name, (rtype, p, kind, seed, B, C, L) in ROT.items():
        x = torch.from_numpy(synth.white(seed, B, C, L))
        tgt = targets(kind, seed + 1, B)
        seed_all(seed)
        rx, rt = ref_augment.Rotation(p, rtype)(x.clone(), {k: v.clone() for k, v in tgt.items()})
        out[name + '/recipe'] = np.array([rtype, int(p * 100), seed, B, C, L])
        out[name + '/kind'] = np.array(kind)
        out[name + '/x'] = rx.numpy()
        out[name + '/label'] = rt[kind].numpy()
        assert not np.array_equal(rx.numpy(), x.numpy())
    for name, (alpha, seed, ov, C, L) in MIX.items():
        B = len(ov)
        x = torch.from_numpy(synth.white(seed, B, C, L))
        tgt = {'ov': list(ov), 'accdoa_label': torch.from_numpy(synth.uniform(seed + 1, (B, 6, 15)))}
        seed_all(seed)
        rx, _ = ref_augment.WavMix(alpha, 1.0)(x.clone(), {'ov': list(ov), 'accdoa_label': tgt['accdoa_label'].clone()})
        # replay the draws of wavmix.py:22-40
        seed_all(seed)
        assert not random.random() > 1.0
        idx1 = [n for n in range(B) if ov[n] == '1']
        idx2 = [n for n in range(B) if ov[n] == '2']
        add_ov = random.choice(['1', '2'])
        new_idx = np.random.permutation(idx1 if add_ov == '1' else idx2)
        N = min(len(idx1), len(new_idx))
        lambs = torch.distributions.beta.Beta(alpha, alpha).sample((N,))
        dst, src = np.array(idx1[:N]), np.array(new_idx[:N])
        chk = x.clone()
        lx = lambs.reshape(N, 1, 1)
        chk[dst] = lx * chk[dst] + (1. - lx) * chk[src]                     # wavmix.py:50
        assert torch.equal(chk, rx), name
        out[name + '/recipe'] = np.array([seed, B, C, L])
        out[name + '/add_ov'] = np.array(int(add_ov))
        out[name + '/dst'] = dst
        out[name + '/src'] = src
        out[name + '/lambs'] = lambs.numpy()
        out[name + '/x'] = rx.numpy()
        print(name, 'add_ov', add_ov, 'dst', dst, 'src', src)
    np.savez_compressed(os.path.join(HERE, 'augment.npz'), **out)
    print('wrote augment.npz', sum(v.nbytes for v in out.values()) // 1024, 'KiB raw')


if __name__ == '__main__':
    main()
