"""GPU: MIC path (log-mel + GCC-PHAT) through LogmelGCC_Extractor -> C ABI -> CUDA against (a) the fixtures of
tests/golden/mic.npz -- an evaluation of the reference's definitions with an independent library stack (torch.stft /
torchaudio Slaney bank / torch.fft.irfft, tests/golden/make_golden_mic.py; librosa itself is not installable
offline) -- and (b) the numpy oracle (oracle/seld_oracle.py:logmel_gcc, pinned to the same fixtures by
tests/test_oracle.py).  Tolerances (north_star): log-mel 1e-4 of the block maximum, GCC-PHAT 1e-4 absolute."""
import numpy as np
import pytest
import torch

from conftest import block_err, make_cfg
import pseldnets_b200 as pb
from pseldnets_b200 import _abi

pytestmark = pytest.mark.gpu


def _mic(sr=24000, hop=240):
    return pb.get_afextractor(make_cfg(sr, hop, 'hann', 'logmelgcc')).cuda()


def _oracle(ext, x, dtype=np.float64, top_db=80.0):
    from oracle import seld_oracle as so
    return so.logmel_gcc(x, ext.stft_extractor.window.cpu().numpy(), ext.mel_scale.fb.cpu().numpy(),
                         1024, ext.hop, top_db=top_db, dtype=dtype)


def _check(y, ref, what):
    assert y.shape == ref.shape, (y.shape, ref.shape)
    assert np.isfinite(y).all()
    e = block_err(y, ref, slice(0, 4))
    assert e <= 1e-4, '%s: log-mel block error %.3e' % (what, e)
    g = float(np.abs(y[:, 4:].astype(np.float64) - ref[:, 4:]).max())
    assert g <= 1e-4, '%s: GCC abs error %.3e' % (what, g)


def _delta_at_lag0(g):
    """GCC planes (..., T, 64) that are a unit pulse at lag 0 (column 32): irfft of all-ones phasors"""
    return np.abs(g[..., 32] - 1.0).max() < 1e-5 and np.abs(np.delete(g, 32, axis=-1)).max() < 1e-5


def test_mic_against_independent_fixtures():
    """Every case of tests/golden/mic.npz: white, full scale, pure delays, 120 dB quiet tail, zero-filled tail, dead
    microphone, 32 kHz, ragged length -- kernel vs the torch/torchaudio evaluation in fp64."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'mic.npz'))
    for name in g['names']:
        sr, hop = (int(v) for v in g[name + '/sr_hop'])
        ext = _mic(sr, hop)
        x, ref = g[name + '/x'], g[name + '/y64'].copy()
        y = ext(torch.from_numpy(x).cuda()).cpu().numpy()
        sz = [4 + int(p) for p in g[name + '/signed_zero_planes']]
        if sz:
            # Pairs with a digitally silent microphone: R = conj(X_m) X_n is an exact zero and angle(R) hangs on the
            # SIGNS of that zero (atan2(+0, -0) = pi).  numpy and torch disagree with each other on these planes
            # (tests/test_oracle.py asserts > 1e-2 between them), i.e. the reference is implementation-defined; the
            # kernel takes angle(0) = 0, phasor 1: a unit pulse at lag 0, as every implementation gives for an
            # all-silent frame.  Those planes are checked for that; all others against the fixture.
            assert _delta_at_lag0(y[:, sz]), name
            keep = [p for p in range(10) if p not in sz]
            y, ref = y[:, keep], ref[:, keep]
        _check_planes(y, ref, name)
        ext.top_db = None
        y2 = ext(torch.from_numpy(x).cuda()).cpu().numpy()
        e = block_err(y2[:, :4], g[name + '/y64_notopdb'], slice(0, 4))
        assert e <= 1e-4, '%s (top_db=None): log-mel block error %.3e' % (name, e)


def _check_planes(y, ref, what):
    """like _check for a subset of planes: the first four are log-mel, the rest GCC"""
    assert y.shape == ref.shape and np.isfinite(y).all(), what
    e = block_err(y, ref, slice(0, 4))
    assert e <= 1e-4, '%s: log-mel block error %.3e' % (what, e)
    gerr = float(np.abs(y[:, 4:].astype(np.float64) - ref[:, 4:]).max())
    assert gerr <= 1e-4, '%s: GCC abs error %.3e' % (what, gerr)


def test_mic_against_oracle_small_and_ragged():
    from oracle import synth
    ext = _mic()
    for seed, B, L in ((31, 2, 4800), (32, 1, 5003), (33, 3, 1200)):
        x = synth.white(seed, B, 4, L)
        y = ext(torch.from_numpy(x).cuda()).cpu().numpy()
        assert y.shape == (B, 10, L // 240, 64)
        _check(y, _oracle(ext, x), 'L=%d' % L)


def test_mic_32k():
    from oracle import synth
    ext = _mic(32000, 320)
    x = synth.white(34, 2, 4, 9600)
    _check(ext(torch.from_numpy(x).cuda()).cpu().numpy(), _oracle(ext, x), 'sr=32k')


def test_mic_pure_delay_peaks_at_lag():
    from oracle import synth
    ext = _mic()
    s = synth.white(35, 1, 1, 24000 + 32)[0, 0]
    d = 7
    x = np.stack([s[16:16 + 24000], s[16 - d:16 - d + 24000], s[16 + 3:16 + 3 + 24000], s[16:16 + 24000]])[None]
    y = ext(torch.from_numpy(np.ascontiguousarray(x)).cuda()).cpu().numpy()
    mid = y[0, :, 50]
    assert int(np.argmax(mid[4])) == 32 + d          # pair (0,1): mic1 lags mic0 by d
    assert int(np.argmax(mid[5])) == 32 - 3          # pair (0,2): mic2 leads by 3
    assert int(np.argmax(mid[6])) == 32              # pair (0,3): identical
    _check(y, _oracle(ext, x), 'delay')


def test_mic_top_db_floor_and_disable():
    from oracle import synth
    ext = _mic()
    x = synth.white(36, 2, 4, 9600)
    x[:, :, 4800:] *= 1e-6                            # 120 dB quieter tail -> floor at max - 80 dB
    xt = torch.from_numpy(x).cuda()
    y = ext(xt).cpu().numpy()
    ref = _oracle(ext, x)
    _check(y, ref, 'top_db=80')
    for b in range(2):
        for c in range(4):
            assert abs(y[b, c].min() - (y[b, c].max() - 80.0)) < 1e-3
    ext.top_db = None
    y2 = ext(xt).cpu().numpy()
    _check(y2, _oracle(ext, x, top_db=None), 'top_db=None')
    assert y2[:, :4].min() < y[:, :4].min() - 10.0     # without the floor the quiet tail sits at the amin clamp


def test_mic_silence_and_errors():
    ext = _mic()
    y = ext(torch.zeros(1, 4, 2400, device='cuda')).cpu().numpy()
    assert np.abs(y[:, :4] + 100.0).max() < 1e-4
    # phasor of 0 is 1 -> irfft of all-ones = delta at lag 0 (column 32)
    g = y[0, 4:, 5]
    assert np.abs(g[:, 32] - 1.0).max() < 1e-5 and np.abs(np.delete(g, 32, axis=1)).max() < 1e-5
    with pytest.raises(_abi.SeldError):
        ext(torch.zeros(1, 3, 2400, device='cuda'))
    with pytest.raises(ValueError):
        ext(torch.zeros(4, 2400, device='cuda'))
    assert ext(torch.zeros(0, 4, 2400, device='cuda')).shape == (0, 10, 10, 64)
    assert ext(torch.zeros(2, 4, 100, device='cuda')).shape == (2, 10, 0, 64)


def test_mic_dead_channel_and_silent_tail():
    """Vanishing cross-spectrum bins (angle(0) = 0 -> phasor 1) next to ordinary ones: a clip whose second half
    is zero fill (frames with and without vanishing bins in one launch), and one digitally silent microphone.

    For the pairs that involve the silent microphone the reference's own output is not well defined: numpy's
    complex product keeps signed zeros and angle(-0 + 0j) = pi, so its phasors there are +-1 in a pattern set by
    the signed zeros its FFT library happens to leave in an all-zero spectrum.  The kernel takes the documented
    convention angle(0) = 0 -- a delta at lag 0, exactly as in the all-silent case, where the reference agrees --
    and those planes are checked for that; everything else has to match the oracle."""
    from oracle import synth
    ext = _mic()
    for loud in (False, True):
        x = synth.uniform(42, (2, 4, 4800)).astype(np.float32) if loud else synth.white(41, 2, 4, 4800)
        x[0, 2] = 0.0                                        # dead microphone in clip 0
        x[1, :, 2400:] = 0.0                                 # clip 1: silent tail
        y = ext(torch.from_numpy(x).cuda()).cpu().numpy()
        ref = _oracle(ext, x)
        dead_planes = [4 + 1, 4 + 3, 4 + 5]                  # pairs (0,2) (1,2) (2,3)
        live_planes = [p for p in range(10) if p not in dead_planes]
        assert _delta_at_lag0(y[0][dead_planes])             # the documented convention (see test_mic_against_independent_fixtures)
        _check_planes(y[:1, live_planes], ref[:1, live_planes], 'dead channel, planes without it (loud=%s)' % loud)
        _check(y[1:], ref[1:], 'silent tail (loud=%s)' % loud)
        assert np.abs(y[0, 2] + 100.0).max() < 1e-4          # its log-mel is the amin clamp
        g = y[1, 4:, -1]                                     # all four silent: every phasor is 1 -> delta at lag 0
        assert np.abs(g[:, 32] - 1.0).max() < 1e-5 and np.abs(np.delete(g, 32, axis=1)).max() < 1e-5


def test_mic_numpy_front():
    from oracle import synth
    cfg = make_cfg(24000, 240, 'hann', 'logmelgcc')
    front = pb.Features_Extractor_MIC(cfg)
    audio = synth.white(37, 1, 4, 4800)[0].T.copy()             # (L, C) soundfile layout
    feat = front.extract_logmelgcc(audio)
    assert feat.shape == (10, 20, 64) and feat.dtype == np.float32
    _check(feat[None], _oracle(front._ext, audio.T[None]), 'numpy front')


def test_mic_cfg3_full_size():
    """BASELINE cfg3: B=64 x 10 s x 4 mics -> (64, 10, 1000, 64); clips 0-1 against the oracle,
    batch independence and gain invariance of GCC at full size."""
    ext = _mic()
    g = torch.Generator(device='cuda').manual_seed(1235)
    x = 0.1 * torch.randn(64, 4, 240000, device='cuda', generator=g)
    y = ext(x)
    assert y.shape == (64, 10, 1000, 64) and torch.isfinite(y).all()
    for b in (0, 40, 63):
        assert torch.equal(ext(x[b:b + 1]), y[b:b + 1])
    y3 = ext(3.0 * x[:2])
    assert (y3[:, 4:] - y[:2, 4:]).abs().max().item() < 1e-4
    assert (y3[:, :4] - y[:2, :4] - 20.0 * np.log10(3.0)).abs().max().item() < 1e-4 * y[:2, :4].abs().max().item()
    _check(y[:2].cpu().numpy(), _oracle(ext, x[:2].cpu().numpy()), 'cfg3 clips 0-1')


def test_mic_reference_stage_methods_run_preprocess_unchanged():
    """Features_Extractor_MIC with the reference's three stages: the body of Preprocess.extract_mic_features
    (preprocess.py:546-556) runs against the drop-in verbatim and yields the fixture features; the spectrogram has
    the reference's layout and values ((T, 513, C) complex64: torch.stft with zero centre padding)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'mic.npz'))
    for name in ('white_24k', 'delay_24k', 'quiet_tail_24k', 'ragged_24k', 'white_32k'):
        sr, hop = (int(v) for v in g[name + '/sr_hop'])
        af_extractor_mic = pb.Features_Extractor_MIC(make_cfg(sr, hop, 'hann', 'logmelgcc'))
        hoplen = hop
        for b in range(g[name + '/x'].shape[0]):
            waveform = np.ascontiguousarray(g[name + '/x'][b].T)                 # sf.read layout: (L, C) float32
            # ---- preprocess.py:546-556, unchanged
            nb_feat_frams = int(len(waveform) / hoplen)
            spect = af_extractor_mic._spectrogram(waveform, nb_feat_frams)
            logmel_spec = af_extractor_mic._get_logmel_spectrogram(spect)
            gcc = af_extractor_mic._get_gcc(spect)
            feature = np.concatenate((logmel_spec, gcc), axis=-1).transpose((2, 0, 1))
            # ----
            assert spect.shape == (nb_feat_frams, 513, 4) and spect.dtype == np.complex64
            assert logmel_spec.shape == (nb_feat_frams, 64, 4) and gcc.shape == (nb_feat_frams, 64, 6)
            _check(feature[None].astype(np.float32), g[name + '/y64'][b:b + 1], name + ' via the three stages')
            fused = af_extractor_mic.extract_logmelgcc(waveform)
            _check(fused[None], g[name + '/y64'][b:b + 1], name + ' fused')
            # the spectrogram itself against torch.stft (fp64)
            xt = torch.from_numpy(g[name + '/x'][b]).double()
            X = torch.stft(xt, 1024, hop, 1024, torch.hann_window(1024, dtype=torch.float64), center=True,
                           pad_mode='constant', return_complex=True)[..., :nb_feat_frams].permute(2, 1, 0).numpy()
            assert np.abs(spect - X).max() <= 2e-6 * np.abs(X).max(), name
    # a spectrogram that did not come from _spectrogram (any array of that layout works), and fewer frames than available
    front = pb.Features_Extractor_MIC(make_cfg(24000, 240, 'hann', 'logmelgcc'))
    wav = np.ascontiguousarray(g['white_24k/x'][0].T)
    sp = front._spectrogram(wav, 9)
    assert sp.shape == (9, 513, 4)
    lm = front._get_logmel_spectrogram(sp.copy())
    assert lm.shape == (9, 64, 4) and np.isfinite(lm).all()


@pytest.mark.parametrize('quiet', [0, 1, 2, 3])
@pytest.mark.parametrize('db', [20, 40, 60, 80, 100])
def test_mic_microphone_attenuated_against_its_partner(quiet, db):
    """Finite level imbalance inside a microphone pair that shares a packed transform (0/1 and 2/3): PHAT keeps only
    phases, so the partner's rounding noise in a quiet microphone's spectrum shows in every GCC plane of that
    microphone.  Frames whose microphones are too far apart are transformed again with each microphone alone
    (as the FOA kernels do)."""
    from oracle import synth
    ext = _mic()
    x = synth.uniform(700 + 10 * quiet + db, (2, 4, 4800)).astype(np.float32)
    x[:, quiet] *= np.float32(10.0 ** (-db / 20.0))
    y = ext(torch.from_numpy(x).cuda()).cpu().numpy()
    _check(y, _oracle(ext, x), 'microphone %d at -%d dB' % (quiet, db))
