"""GPU: the inference-side callers against what the reference does -- calling
the extractor on every transformed / overlapping waveform."""
import numpy as np
import pytest
import torch

from conftest import make_cfg
import pseldnets_b200 as pb
from pseldnets_b200 import inference as inf

pytestmark = pytest.mark.gpu


def _ext():
    return pb.LogmelIV_Extractor(make_cfg()).cuda()


def test_acs_variants_from_one_extraction():
    ext = _ext()
    g = torch.Generator(device='cuda').manual_seed(5)
    x = 0.1 * torch.randn(3, 4, 24000, device='cuda', generator=g)
    feat = ext(x)
    n = 0
    for xv, fv in zip(inf.acs_waveform_variants(x), inf.acs_feature_variants(feat)):
        ref = ext(xv.contiguous())
        assert torch.equal(fv[:, :4], ref[:, :4]) or (fv[:, :4] - ref[:, :4]).abs().max().item() < 1e-4 * ref[:, :4].abs().max().item()
        assert (fv[:, 4:] - ref[:, 4:]).abs().max().item() < 1e-5 * ref[:, 4:].abs().max().item()
        n += 1
    assert n == 16


def test_overlapped_chunks_equal_per_chunk_extraction():
    """evalMA setting: 10-s chunks every 0.5 s; a 17.3-s recording (ragged tail)."""
    ext = _ext()
    g = torch.Generator(device='cuda').manual_seed(6)
    for L in (415200, 420000, 300000 + 7):
        x = 0.1 * torch.randn(4, L, device='cuda', generator=g)
        feats, idx = inf.extract_overlapped(ext, x, 240000, 12000)
        chunks = torch.stack([torch.nn.functional.pad(x[:, b:e], (0, 240000 - (e - b))) for b, e in idx])
        ref = ext(chunks)
        assert feats.shape == ref.shape
        assert torch.equal(feats, ref), L
    before = pb._abi.lib().seld_launch_count()
    inf.extract_overlapped(ext, x, 240000, 12000)
    assert pb._abi.lib().seld_launch_count() - before <= 4          # recording, heads, tails (+ ragged last chunk)
