"""CUDA-graph replay of the front-end for fixed shapes (small-batch / serving latency).

One call of an extractor costs one kernel launch plus the Python / ctypes path around it
(argument checks, output allocation, plan lookup) -- tens of microseconds, more than the kernel
itself takes for a single 10-s clip.  For inference, where shapes repeat (the reference evaluates
fixed 10-s chunks, `components/model_module.py:304-330`), `GraphedFrontEnd` records

    waveform -> extractor [-> scalar -> reshape_wav2img]

once into a CUDA graph on static buffers and replays it per call: one `cudaGraphLaunch`, no
allocation, nothing traced or compiled -- the recorded nodes are the library's own kernels.
"""
import torch

from . import epilogue


class GraphedFrontEnd:
    """Record `extractor(x)` (and optionally the backbone-input stage) for one input shape.

    extractor : LogmelIV_Extractor / Logmel_Extractor / LogmelGCC_Extractor on a CUDA device
    shape     : (B, C, L) of every batch that will be fed
    scalar    : None, or the backbone's `scalar` ModuleList / ScalarParams (eval mode)
    spec_size : None -> output is the feature map (B, C', T, M) (after the scalar if given);
                int  -> output is HTS-AT's image (B, C', spec_size, spec_size)
    dtype     : torch.float32, or torch.int16 for PCM input (LogmelIV_Extractor only)

    __call__(x) copies x into the static input, replays, and returns the result -- by default a
    fresh clone (the reference's extractor returns a tensor the caller may mutate and keep);
    pass clone=False to get the static output buffer itself, valid until the next call.
    """

    def __init__(self, extractor, shape, scalar=None, spec_size=None, dtype=torch.float32, warmup=3):
        dev = next(extractor.buffers()).device
        if dev.type != 'cuda':
            raise RuntimeError('GraphedFrontEnd needs the extractor on a CUDA device; there is no CPU path')
        self.extractor = extractor
        self.shape = tuple(int(s) for s in shape)
        self.static_in = torch.zeros(self.shape, dtype=dtype, device=dev)
        self._scalar = None if scalar is None else (
            scalar if isinstance(scalar, epilogue.ScalarParams) else epilogue.ScalarParams(scalar, dev))
        self._spec_size = spec_size
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                     # warm up outside the capture: plans, lazy CUDA state
            for _ in range(max(1, warmup)):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self._run()
        # The recorded kernel nodes hold pointers into the plan's device tables: keep that plan alive for as long as
        # the graph exists, and remember which buffer contents it was built from.
        self._plan = extractor._plan(dev)
        self._plan_key = self._buffers_key()

    def _buffers_key(self):
        win, fb = self.extractor.stft_extractor.window, self.extractor.mel_scale.fb
        return (win.data_ptr(), win._version, fb.data_ptr(), fb._version)

    def _check_buffers(self):
        if self._buffers_key() != self._plan_key:
            raise RuntimeError('the extractor\'s window / mel bank changed after the graph was recorded '
                               '(load_state_dict, .to(), in-place edit): record a new GraphedFrontEnd')

    def _run(self):
        y = self.extractor(self.static_in)
        if self._spec_size is not None:
            return epilogue.scalar_wav2img(y, self._scalar, self._spec_size)
        if self._scalar is not None:
            epilogue.apply_scalar(y, self._scalar)
        return y

    def replay(self):
        """Replay on whatever `static_in` holds (fill it directly, e.g. as the target of the host->device copy,
        to skip the staging copy of __call__); returns the static output buffer."""
        self._check_buffers()
        self.graph.replay()
        return self.static_out

    def __call__(self, x, clone=True):
        if tuple(x.shape) != self.shape or x.dtype != self.static_in.dtype:
            raise ValueError('GraphedFrontEnd was recorded for %s %s, got %s %s'
                             % (self.shape, self.static_in.dtype, tuple(x.shape), x.dtype))
        self._check_buffers()
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out.clone() if clone else self.static_out
