"""Drop-in extractor modules: same names, constructor and forward contract as the reference's
/root/reference/src/utils/feature.py, with the arithmetic done by the fused sm_100a kernels of
libseldfeat.so instead of ~45 torchaudio / ATen / cuFFT / cuBLAS launches.

    LogmelIV_Extractor(cfg)(x)  x (B, C>=4, L) fp32 cuda -> (B, C+3, 1+L//hop, n_mels)   feature.py:20-56
    Logmel_Extractor(cfg)(x)    x (B, C,    L) fp32 cuda -> (B, C,   1+L//hop, n_mels)   feature.py:59-91

State: no parameters; the same two persistent buffers under the same names the reference's
checkpoints contain (`stft_extractor.window`, `mel_scale.fb`).  The kernels take their tables from
those buffers, so a loaded state_dict is honoured.  Output is a fresh tensor each call (callers
mutate it in place: accdoa.py:224-227, specaug.py:55-57).  Runs on the current CUDA stream; no
backward (the reference's input never requires grad).
"""
import ctypes

import torch
import torch.nn as nn

from . import _abi
from .filterbank import (librosa_mel_bank, make_window, melscale_fbanks_htk_slaney,  # noqa: F401
                         window_fn_dict)

eps = torch.finfo(torch.float32).eps  # feature.py:8
AMIN = 1e-10                          # torchaudio AmplitudeToDB


class _Buffer(nn.Module):
    """Holder giving a buffer the attribute path it has in the reference's state_dict."""

    def __init__(self, name, value):
        super().__init__()
        self.register_buffer(name, value)


class _Plan:
    """Owns one seld_plan (device tables) for a (device, buffer contents) pair."""

    def __init__(self, device_index, window, fb, n_fft, hop, n_mels):
        w = window.detach().to('cpu', torch.float32).contiguous()
        f = fb.detach().to('cpu', torch.float32).contiguous()
        handle = ctypes.c_void_p()
        code = _abi.lib().seld_plan_create(
            ctypes.byref(handle), device_index,
            ctypes.cast(w.data_ptr(), ctypes.POINTER(ctypes.c_float)),
            ctypes.cast(f.data_ptr(), ctypes.POINTER(ctypes.c_float)),
            n_fft, hop, n_mels, AMIN, eps)
        _abi.check(code, 'seld_plan_create')
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                _abi.lib().seld_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _ExtractorBase(nn.Module):
    _entry = None       # C-ABI symbol
    _entry_i16 = None   # C-ABI symbol for int16 PCM input, if the path has one
    _extra_ch = 0

    def __init__(self, cfg):
        super().__init__()
        data = cfg['data']
        assert data['window'] in window_fn_dict.keys(), \
            "window must be in {}, but got {}".format(window_fn_dict.keys(), data['window'])
        self.n_fft = int(data['nfft'])
        self.hop = int(data['hoplen'])
        self.n_mels = int(data['n_mels'])
        self.sample_rate = data['sample_rate']
        self.stft_extractor = _Buffer('window', make_window(data['window'], self.n_fft))
        self.mel_scale = _Buffer('fb', self._make_bank())
        self._plans = {}

    def __getstate__(self):
        # Plans own raw device handles (ctypes pointers): they are per-process caches, rebuilt on first use.
        # Keeping them out of the copied / pickled state is what makes copy.deepcopy(model), pickle and
        # torch.save(model) work after a forward, as they do for the reference module (feature.py:20-37),
        # and keeps two module copies from sharing (and double-freeing) one handle.
        state = self.__dict__.copy()
        state['_plans'] = {}
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._plans = {}

    def _make_bank(self):
        # MelScale(norm='slaney', f_min=20, f_max=sr/2, mel_scale='htk' default) -- feature.py:32-34
        return melscale_fbanks_htk_slaney(self.n_fft // 2 + 1, 20, self.sample_rate / 2, self.n_mels,
                                          self.sample_rate)

    def _plan(self, device):
        win, fb = self.stft_extractor.window, self.mel_scale.fb
        key = (device.index, win.data_ptr(), win._version, fb.data_ptr(), fb._version)
        plan = self._plans.get(device.index)
        if plan is None or plan[0] != key:
            plan = (key, _Plan(device.index, win, fb, self.n_fft, self.hop, self.n_mels))
            self._plans[device.index] = plan
        return plan[1]

    def forward(self, x):
        """
        input:
            (batch_size, channels, data_length)
        output:
            (batch_size, channels(+3), time_steps, mel_bins)
        """
        if x.ndim != 3:
            raise ValueError("x shape must be (batch_size, num_channels, data_length)\n \
                            Now it is {}".format(x.shape))
        if not x.is_cuda:
            raise RuntimeError('pseldnets_b200 extractors run on CUDA tensors only (no CPU path); '
                               'got a tensor on %s' % x.device)
        entry = self._entry
        if x.dtype == torch.int16 and self._entry_i16 is not None:
            entry = self._entry_i16          # 16-bit PCM as decoded from wav/flac: converted in-kernel as s / 32768
        elif x.dtype != torch.float32:
            raise TypeError('expected float32 waveform, got %s' % x.dtype)
        if x.stride(2) != 1:
            x = x.contiguous()
        B, C, L = x.shape
        dev = x.device
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        plan = self._plan(dev)
        T = 1 + L // self.hop
        out = torch.empty((B, C + self._extra_ch, T, self.n_mels), dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        fn = getattr(_abi.lib(), entry)
        with _abi.device_guard(x.device):
            code = fn(plan.handle, x.data_ptr(), B, C, L, x.stride(0), x.stride(1), out.data_ptr(), stream)
        _abi.check(code, entry)
        return out


class LogmelIV_Extractor(_ExtractorBase):
    """log-mel of every channel + mel-projected normalised intensity vector of channels 1..3
    against channel 0 (FOA: W, Y, Z, X).  feature.py:20-56, 93-117."""
    _entry = 'seld_logmel_iv_f32'
    _entry_i16 = 'seld_logmel_iv_i16'
    _extra_ch = 3


    def forward_host(self, x, out=None, device=None, chunk_clips=0, synchronize=True):
        """Host-buffer form of forward(): x is a CPU float32 tensor (B, C, L) (page-locked for
        full overlap); returns a page-locked CPU tensor (B, C+3, T, n_mels).  Chunks of the batch
        are copied in, transformed and copied out on three overlapping streams inside
        libseldfeat.so (seld_logmel_iv_f32_host).  With synchronize=False the call only enqueues:
        the result is complete once the current stream of `device` has been synchronised."""
        if x.ndim != 3:
            raise ValueError("x shape must be (batch_size, num_channels, data_length)\n \
                            Now it is {}".format(x.shape))
        if x.is_cuda or x.dtype not in (torch.float32, torch.int16):
            raise TypeError('forward_host expects a CPU float32 (or int16 PCM) tensor')
        x = x.contiguous()
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        B, C, L = x.shape
        T = 1 + L // self.hop
        if out is None:
            out = torch.empty((B, C + 3, T, self.n_mels), dtype=torch.float32, pin_memory=True)
        plan = self._plan(dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        name = 'seld_logmel_iv_i16_host' if x.dtype == torch.int16 else 'seld_logmel_iv_f32_host'
        code = getattr(_abi.lib(), name)(plan.handle, x.data_ptr(), B, C, L, out.data_ptr(), int(chunk_clips), stream)
        _abi.check(code, name)
        if synchronize:
            torch.cuda.current_stream(dev).synchronize()
        return out


class Logmel_Extractor(_ExtractorBase):
    """log-mel of every channel.  feature.py:59-91."""
    _entry = 'seld_logmel_f32'
    _extra_ch = 0


class LogmelGCC_Extractor(_ExtractorBase):
    """MIC-format features as an nn.Module (the reference only has the numpy/librosa class
    `Features_Extractor_MIC`, feature.py:119-175, driven by preprocess.py:546-556):
    x (B, 4, L) -> (B, 4 + 6, int(L/hop), n_mels): per-mic log-mel (librosa Slaney bank,
    power_to_db with top_db=80 per plane) + GCC-PHAT of the 6 mic pairs, lags [-n_mels/2, n_mels/2).
    `mel_scale.fb` holds `librosa.filters.mel(sr, n_fft, n_mels).T`."""
    _entry = 'seld_logmel_gcc_f32'
    top_db = 80.0            # librosa.power_to_db default

    def _make_bank(self):
        return librosa_mel_bank(self.sample_rate, self.n_fft, self.n_mels)      # feature.py:126

    def forward(self, x):
        if x.ndim != 3:
            raise ValueError("x shape must be (batch_size, num_channels, data_length)\n \
                            Now it is {}".format(x.shape))
        if not x.is_cuda:
            raise RuntimeError('pseldnets_b200 extractors run on CUDA tensors only (no CPU path); '
                               'got a tensor on %s' % x.device)
        if x.dtype != torch.float32:
            raise TypeError('expected float32 waveform, got %s' % x.dtype)
        if x.stride(2) != 1:
            x = x.contiguous()
        B, C, L = x.shape
        dev = x.device
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        plan = self._plan(dev)
        lib = _abi.lib()
        T = L // self.hop
        out = torch.empty((B, C + C * (C - 1) // 2, T, self.n_mels), dtype=torch.float32, device=x.device)
        ws = torch.empty((max(1, lib.seld_workspace_bytes(plan.handle, B, C) // 4),), dtype=torch.int32,
                         device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        top_db = -1.0 if self.top_db is None else float(self.top_db)
        with _abi.device_guard(x.device):
            code = lib.seld_logmel_gcc_f32(plan.handle, x.data_ptr(), B, C, L, x.stride(0), x.stride(1), top_db,
                                           out.data_ptr(), ws.data_ptr(), ws.numel() * 4, stream)
        _abi.check(code, self._entry)
        return out


class Features_Extractor_MIC():
    """Drop-in for the reference's MIC class (feature.py:119-175) with its three stages, so that
    Preprocess.extract_mic_features (preprocess.py:546-556) runs against it unchanged:

        spect  = ext._spectrogram(waveform, nb_frames)        # (T, n_fft/2+1, C) complex64
        logmel = ext._get_logmel_spectrogram(spect)           # (T, n_mels, C)
        gcc    = ext._get_gcc(spect)                          # (T, n_mels, C(C-1)/2)

    Each stage is one CUDA launch (seld_mic_spectrogram_f32 / seld_logmel_gcc_from_spectra_f32); the two feature
    stages share one launch when they are given the same spectrogram array.  `extract_logmelgcc(waveform)` is the
    fused path (one launch, the spectrogram never leaves the SM) and returns preprocess.py's final
    (C + C(C-1)/2, T, n_mels) float32 array directly.  `mel_bank` is `librosa.filters.mel(...).T` as in the
    reference.  SALSA-Lite (`_get_salsalite`) is outside this path (it crashes in the reference on numpy >= 1.24).
    """

    def __init__(self, cfg, device='cuda'):
        self.fs = cfg['data']['sample_rate']
        self.n_fft = cfg['data']['nfft']
        self.n_mels = cfg['data']['n_mels']
        self.hoplen = cfg['data']['hoplen']
        self.window = cfg['data']['window']
        self._ext = LogmelGCC_Extractor(cfg).to(device)
        self.mel_bank = self._ext.mel_scale.fb.cpu().numpy()
        self._last = None                      # (spectrogram array, features) of the latest from-spectra launch

    def _device(self):
        dev = self._ext.mel_scale.fb.device
        return dev if dev.index is not None else torch.device('cuda', torch.cuda.current_device())

    def _spectrogram(self, audio_input, _nb_frames):
        """(L, C) float waveform in soundfile layout -> (min(_nb_frames, L // hop), n_fft/2+1, C) complex64"""
        import numpy as np
        ext, dev = self._ext, self._device()
        x = torch.as_tensor(np.ascontiguousarray(np.asarray(audio_input, dtype=np.float32).T)).to(dev)   # (C, L)
        C, L = x.shape
        T = L // ext.hop
        spec = torch.empty((1, T, ext.n_fft // 2 + 1, C), dtype=torch.complex64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with _abi.device_guard(dev):
            code = _abi.lib().seld_mic_spectrogram_f32(ext._plan(dev).handle, x.data_ptr(), 1, C, L, C * L, L,
                                                       spec.data_ptr(), stream)
        _abi.check(code, 'seld_mic_spectrogram_f32')
        return spec[0, :_nb_frames].cpu().numpy()

    def _features_from_spectra(self, linear_spectra):
        import numpy as np
        if self._last is not None and self._last[0] is linear_spectra:
            return self._last[1]
        ext, dev = self._ext, self._device()
        sp = np.ascontiguousarray(np.asarray(linear_spectra, dtype=np.complex64))
        if sp.ndim != 3 or sp.shape[1] != ext.n_fft // 2 + 1:
            raise ValueError('linear_spectra must be (n_frames, n_fft/2+1, n_channels), got %s' % (sp.shape,))
        T, _, C = sp.shape
        spec = torch.from_numpy(sp).to(dev)
        out = torch.empty((1, C + C * (C - 1) // 2, T, ext.n_mels), dtype=torch.float32, device=dev)
        lib = _abi.lib()
        plan = ext._plan(dev)
        ws = torch.empty((max(1, lib.seld_workspace_bytes(plan.handle, 1, C) // 4),), dtype=torch.int32, device=dev)
        top_db = -1.0 if ext.top_db is None else float(ext.top_db)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with _abi.device_guard(dev):
            code = lib.seld_logmel_gcc_from_spectra_f32(plan.handle, spec.data_ptr(), 1, C, T, top_db, out.data_ptr(),
                                                        ws.data_ptr(), ws.numel() * 4, stream)
        _abi.check(code, 'seld_logmel_gcc_from_spectra_f32')
        feat = out[0].cpu().numpy()
        self._last = (linear_spectra, feat)
        return feat

    def _get_logmel_spectrogram(self, linear_spectra):
        """(T, F, C) complex -> (T, n_mels, C) float64 container, as the reference's np.zeros(...) gives"""
        import numpy as np
        C = linear_spectra.shape[-1]
        return np.ascontiguousarray(self._features_from_spectra(linear_spectra)[:C].transpose(1, 2, 0)).astype(np.float64)

    def _get_gcc(self, linear_spectra):
        """(T, F, C) complex -> (T, n_mels, C(C-1)/2) float64 container"""
        import numpy as np
        C = linear_spectra.shape[-1]
        return np.ascontiguousarray(self._features_from_spectra(linear_spectra)[C:].transpose(1, 2, 0)).astype(np.float64)

    def extract_logmelgcc(self, waveform):
        x = torch.as_tensor(waveform, dtype=torch.float32).t().unsqueeze(0)        # (1, C, L)
        return self._ext(x.to(self._ext.mel_scale.fb.device))[0].cpu().numpy()
