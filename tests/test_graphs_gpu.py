"""CUDA-graph replay of the front-end (pseldnets_b200/graphs.py): bit-identical to the eager calls."""
import numpy as np
import pytest
import torch

from conftest import make_cfg
from oracle import synth

pytestmark = pytest.mark.gpu


def _scalar(C, M=64):
    mean, var, weight, bias = synth.scalar_params(3, C, M)
    s = torch.nn.ModuleList([torch.nn.BatchNorm2d(M) for _ in range(C)])
    for c in range(C):
        s[c].running_mean.copy_(torch.from_numpy(mean[c]))
        s[c].running_var.copy_(torch.from_numpy(var[c]))
        s[c].weight.data.copy_(torch.from_numpy(weight[c]))
        s[c].bias.data.copy_(torch.from_numpy(bias[c]))
    return s.cuda().eval()


def test_graphed_foa_to_image_matches_eager():
    import pseldnets_b200 as pb
    from pseldnets_b200.graphs import GraphedFrontEnd
    ext = pb.get_afextractor(make_cfg(24000, 240, 'hann', 'logmelIV')).cuda()
    scalar = _scalar(7)
    g = GraphedFrontEnd(ext, (2, 4, 24000), scalar=scalar, spec_size=256)
    outs = []
    for seed in (1, 2, 3):
        x = torch.from_numpy(synth.white(seed, 2, 4, 24000)).cuda()
        want = pb.scalar_wav2img(ext(x), scalar, 256)
        got = g(x)
        assert got.shape == (2, 7, 256, 256) and torch.equal(got, want)
        outs.append(got)
    assert not torch.equal(outs[0], outs[1])                 # clones: earlier results survive later replays
    view = g(torch.from_numpy(synth.white(1, 2, 4, 24000)).cuda(), clone=False)
    assert view.data_ptr() == g.static_out.data_ptr() and torch.equal(view, outs[0])
    with pytest.raises(ValueError):
        g(torch.zeros(1, 4, 24000, device='cuda'))


def test_graphed_feature_map_variants():
    import pseldnets_b200 as pb
    from pseldnets_b200.graphs import GraphedFrontEnd
    x = torch.from_numpy(synth.white(7, 1, 4, 12000)).cuda()
    ext = pb.get_afextractor(make_cfg(24000, 240, 'hann', 'logmelIV')).cuda()
    assert torch.equal(GraphedFrontEnd(ext, x.shape)(x), ext(x))
    scalar = _scalar(7)
    assert torch.equal(GraphedFrontEnd(ext, x.shape, scalar=scalar)(x), pb.apply_scalar(ext(x), scalar))
    pcm = (x * 32767).round().to(torch.int16)
    assert torch.equal(GraphedFrontEnd(ext, pcm.shape, dtype=torch.int16)(pcm), ext(pcm))
    mic = pb.get_afextractor(make_cfg(24000, 240, 'hann', 'logmelgcc')).cuda()
    gm = GraphedFrontEnd(mic, x.shape)
    for seed in (8, 9):                                      # the top_db maxima are re-initialised inside the graph
        xm = torch.from_numpy(synth.white(seed, 1, 4, 12000)).cuda()
        assert torch.equal(gm(xm), mic(xm))
    with pytest.raises(RuntimeError):
        GraphedFrontEnd(pb.get_afextractor(make_cfg(24000, 240, 'hann', 'logmelIV')), (1, 4, 2400))
