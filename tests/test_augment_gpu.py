"""Waveform-domain augmentation on the GPU, through the C ABI, against goldens made by
running the reference's own augment.Rotation / augment.WavMix on CPU (tests/golden/make_golden_augment.py).
Bar: bit-exact (sign flips, channel moves and the mul-mul-add of wavmix.py:50 in the reference's order)."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def golden():
    return np.load(os.path.join(HERE, 'golden', 'augment.npz'))


def seed_all(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


def targets(kind, seed, B, T=6, K=5):
    shape = {'accdoa_label': (B, T, 3 * K), 'doa_label': (B, T, 2, 3), 'adpit_label': (B, T, 6, 4, K)}[kind]
    return {kind: torch.from_numpy(synth.uniform(seed, shape))}


@pytest.mark.parametrize('name', ['rot48_accdoa', 'rot16_doa', 'rot48_adpit'])
def test_rotation_dropin_matches_reference(golden, name):
    import pseldnets_b200.augment as aug
    rtype, p100, seed, B, C, L = (int(v) for v in golden[name + '/recipe'])
    kind = str(golden[name + '/kind'])
    x = torch.from_numpy(synth.white(seed, B, C, L)).cuda()
    tgt = {k: v.cuda() for k, v in targets(kind, seed + 1, B).items()}
    seed_all(seed)
    rx, rt = aug.Rotation(p100 / 100.0, rtype)(x, tgt)
    assert rx is x
    assert np.array_equal(rx.cpu().numpy(), golden[name + '/x'])
    assert np.array_equal(rt[kind].cpu().numpy(), golden[name + '/label'])


@pytest.mark.parametrize('name', ['mix_a', 'mix_b', 'mix_c', 'mix_d'])
def test_wavmix_waveforms_match_reference(golden, name):
    import pseldnets_b200.augment as aug
    seed, B, C, L = (int(v) for v in golden[name + '/recipe'])
    x = torch.from_numpy(synth.white(seed, B, C, L)).cuda()
    out = aug.wavmix_waveforms(x, golden[name + '/dst'], golden[name + '/src'], torch.from_numpy(golden[name + '/lambs']))
    assert out is x
    assert np.array_equal(x.cpu().numpy(), golden[name + '/x'])


def torch_mix(x, dst, src, lam):
    lx = lam.reshape(-1, 1, 1)
    y = x.clone()
    y[dst] = lx * x[dst] + (1. - lx) * x[src]
    return y


@pytest.mark.parametrize('B,C,L,seed', [(64, 4, 24000, 1), (9, 4, 1001, 2), (5, 1, 7, 3), (16, 7, 4096, 4)])
def test_wavmix_random_chains_against_torch(B, C, L, seed):
    """Random partial permutations (open chains, cycles, fixed points) against the torch expression on the same GPU."""
    import pseldnets_b200.augment as aug
    rng = np.random.default_rng(seed)
    for trial in range(6):
        n = int(rng.integers(1, B + 1))
        dst = rng.permutation(B)[:n]
        src = rng.permutation(B)[:n] if trial % 2 else rng.permutation(dst)        # odd: open chains too; even: pure cycles
        lam = torch.from_numpy(rng.beta(0.5, 0.5, size=n).astype(np.float32)).cuda()
        x = torch.from_numpy(synth.white(seed * 10 + trial, B, C, L)).cuda()
        want = torch_mix(x, torch.from_numpy(dst).cuda(), torch.from_numpy(src).cuda(), lam)
        aug.wavmix_waveforms(x, dst, src, lam)
        assert torch.equal(x, want), (trial, dst, src)


def test_rotate_all_48_against_torch():
    import pseldnets_b200.augment as aug
    B, C, L = 48, 6, 2052                                   # two extra channels ride along untouched
    x = torch.from_numpy(synth.white(77, B, C, L)).cuda()
    want = x.clone()
    codes = []
    combos = [(axes, ch, (sx, sy, sz)) for axes, ch in aug.TRANS_48.items()
              for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
    assert len(combos) == 48
    for b, (axes, (s_x, s_y, s_z), (sx, sy, sz)) in enumerate(combos):
        codes.append(aug.rotation_code((s_x, s_y, s_z), (sy, sz, sx)))
        want[b, :4] = torch.stack((x[b, 0], sy * x[b, s_x], sz * x[b, s_y], sx * x[b, s_z]))   # rotate.py:72
    aug.rotate_waveforms(x, codes)
    assert torch.equal(x, want)
    y = x.clone()
    aug.rotate_waveforms(x, [aug.ROT_IDENTITY] * B)
    assert torch.equal(x, y)


def test_strided_and_unaligned_views():
    import pseldnets_b200.augment as aug
    base = torch.from_numpy(synth.white(5, 6, 4, 1003)).cuda()
    view = base[1:5, :, 1:1000]                              # odd offset: scalar path, batch stride != C * L
    want = view.clone()
    want[2] = torch.stack((view[2, 0], -view[2, 3], view[2, 2], -view[2, 1]))
    keep = base.clone()
    aug.rotate_waveforms(view, [aug.ROT_IDENTITY, aug.ROT_IDENTITY, aug.rotation_code((3, 2, 1), (-1, 1, -1)), aug.ROT_IDENTITY])
    assert torch.equal(view, want)
    keep[1:5, :, 1:1000] = want
    assert torch.equal(base, keep), 'nothing outside the view may change'
    lam = torch.tensor([0.25, 0.7], device='cuda')
    want = torch_mix(view, torch.tensor([0, 3]), torch.tensor([3, 1]), lam)
    aug.wavmix_waveforms(view, [0, 3], [3, 1], lam)
    assert torch.equal(view, want)


def test_errors():
    import pseldnets_b200.augment as aug
    x = torch.zeros((4, 4, 100), device='cuda')
    with pytest.raises(RuntimeError):
        aug.rotate_waveforms(x.cpu(), [aug.ROT_IDENTITY] * 4)
    with pytest.raises(ValueError):
        aug.rotate_waveforms(x, [aug.ROT_IDENTITY] * 3)
    with pytest.raises(ValueError):
        aug.rotate_waveforms(x[:, :3], [aug.ROT_IDENTITY] * 4)
    with pytest.raises(ValueError):
        aug.rotation_code((0, 1, 2), (1, 1, 1))
    with pytest.raises(ValueError):
        aug.wavmix_waveforms(x, [0, 0], [1, 2], [0.5, 0.5])          # repeated destination
    with pytest.raises(ValueError):
        aug.wavmix_waveforms(x, [0, 1], [2, 9], [0.5, 0.5])          # out of range
    with pytest.raises(ValueError):
        aug.Rotation(0.5, 48)(torch.zeros((2, 8, 10), device='cuda'), {'doa_label': torch.zeros(2, 3, 2, 3)})
    assert aug.wavmix_waveforms(x, [], [], []) is x


def test_augment_then_extract_pipeline():
    """Rotation -> WavMix -> extractor as in model_module.py:53-58: same features as extracting the
    torch-augmented batch."""
    import pseldnets_b200 as pb
    import pseldnets_b200.augment as aug
    cfg = {'data': {'sample_rate': 24000, 'nfft': 1024, 'hoplen': 240, 'n_mels': 64, 'window': 'hann',
                    'audio_feature': 'logmelIV'}}
    ext = pb.get_afextractor(cfg).cuda()
    x = torch.from_numpy(synth.white(91, 4, 4, 12000)).cuda()
    ref = x.clone()
    ref[1] = torch.stack((ref[1, 0], -ref[1, 2], ref[1, 1], ref[1, 3]))
    ref = torch_mix(ref, torch.tensor([0, 1]), torch.tensor([1, 3]), torch.tensor([0.3, 0.9], device='cuda'))
    aug.rotate_waveforms(x, [aug.ROT_IDENTITY, aug.rotation_code((2, 1, 3), (-1, 1, 1)), aug.ROT_IDENTITY, aug.ROT_IDENTITY])
    aug.wavmix_waveforms(x, [0, 1], [1, 3], [0.3, 0.9])
    assert torch.equal(x, ref)
    assert torch.equal(ext(x), ext(ref))
