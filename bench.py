#!/usr/bin/env python
"""Benchmark of the SELD feature front-end hot path (BASELINE.json metric: audio-seconds/sec,
log-mel + IV, 4 ch 24 kHz; % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the fused extractor over one batch of synthetic clips (BASELINE cfg2:
B=64 x 10 s x 4 ch @ 24 kHz per GPU; weak scaling: every rank owns its own 64 clips, no
data-path collective).  Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR, HOP, NFFT, NMELS = 24000, 240, 1024, 64
CLIP_S = 10
C = 4
L = SR * CLIP_S
T = 1 + L // HOP
ALGO_BYTES_PER_CLIP = C * L * 4 + (C + 3) * T * NMELS * 4      # 5,633,792
CFG = {'data': {'sample_rate': SR, 'nfft': NFFT, 'hoplen': HOP, 'n_mels': NMELS,
                'window': 'hann', 'audio_feature': 'logmelIV'}}
METRIC = 'audio-seconds/sec (log-mel+IV, 4ch 24 kHz)'
UNIT = 'audio-s/s'


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a background thread (one
    query every ~2 ms); falls back to polling `nvidia-smi` (the B200_PROFILING.md clocks line)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = False
        self.thread = None
        self.source = 'nvml'

    def _loop_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = int(vis.split(',')[self.idx]) if vis and all(v.strip().isdigit() for v in vis.split(',')) else self.idx
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = {'hw_slowdown': nv.nvmlClocksEventReasonHwSlowdown, 'hw_thermal_slowdown': nv.nvmlClocksEventReasonHwThermalSlowdown,
                'sw_thermal_slowdown': nv.nvmlClocksEventReasonSwThermalSlowdown, 'sw_power_cap': nv.nvmlClocksEventReasonSwPowerCap}
        while not self.stop_flag:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            self.mx.append(float(mx))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            for name, bit in bits.items():
                if r & bit:
                    self.reasons.add(name)
            time.sleep(0.002)

    def _loop_smi(self):
        self.source = 'nvidia-smi'
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    r = [v.strip() for v in line.split(',')]
                    self.sm.append(float(r[1])); self.mx.append(float(r[2]))
                    for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                        if v.lower().startswith('active'):
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def _loop(self):
        try:
            self._loop_nvml()
        except Exception:
            self._loop_smi()

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def mark(self):
        """Samples taken from now on belong to the timed region."""
        self.first = len(self.sm)

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=10)
        first = getattr(self, 'first', 0)
        sm = self.sm[first:] or self.sm
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(self.mx) if self.mx else None,
                'reasons': sorted(self.reasons), 'samples': len(sm), 'source': self.source}


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads to the CPUs closest to its GPU (NVML's ideal affinity), so page-locked
    buffers are first-touched on that NUMA node and host<->device copies do not cross sockets."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = int(vis.split(',')[local_rank]) if vis and all(v.strip().isdigit() for v in vis.split(',')) else local_rank
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(phys))
        return True
    except Exception:
        return False


def library_ops_same_gpu(x, window, fb):
    """The reference's op sequence on whatever device x lives on -- feature.py:46-56 + 101-115 as the library calls
    they resolve to (torch.stft with reflect centre padding, |X|^2, (T,F)@(F,M), 10*log10(clamp), three real cross
    terms / norm, three more matmuls, cat).  Used for the same-GPU library baseline only."""
    import torch
    B, Cn, Ln = x.shape
    X = torch.stft(x.reshape(-1, Ln), n_fft=NFFT, hop_length=HOP, win_length=NFFT, window=window, center=True,
                   pad_mode='reflect', normalized=False, onesided=True, return_complex=True)
    X = X.reshape(B, Cn, X.shape[-2], X.shape[-1])                              # (B, C, F, T)
    mel = torch.matmul((torch.abs(X) ** 2).transpose(-1, -2), fb)                # (B, C, T, M)
    logmel = 10.0 * torch.log10(torch.clamp(mel, min=1e-10))
    re, im = X.real.transpose(-1, -2), X.imag.transpose(-1, -2)
    cross = [re[:, 0] * re[:, j] + im[:, 0] * im[:, j] for j in (1, 2, 3)]
    norm = torch.sqrt(cross[0] ** 2 + cross[1] ** 2 + cross[2] ** 2) + torch.finfo(torch.float32).eps
    iv = torch.stack([torch.matmul(c / norm, fb) for c in cross], dim=1)
    return torch.cat((logmel, iv), dim=1)


def cpu_reference_callable():
    """The reference's CPU implementation of the path as a callable x (B, 4, L) -> features, and what it is:
    'reference' = the UNMODIFIED /root/reference/src/utils/feature.py, vendored by __graft_entry__.build() into the
    git-ignored baseline/_ref/ (it travels to the GPU box with the snapshot; only `librosa`, which the FOA extractor
    never touches, is stubbed); 'port' = oracle/torch_port.py (the same torch.stft / matmul calls, pinned to the
    reference's goldens) when that file is not there."""
    import torch
    path = os.path.join(ROOT, 'baseline', '_ref', 'utils', 'feature.py')
    if os.path.exists(path):
        try:
            import importlib.util
            import types
            sys.modules.setdefault('librosa', types.ModuleType('librosa'))
            spec = importlib.util.spec_from_file_location('pseldnets_reference_feature', path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            ext = mod.LogmelIV_Extractor(CFG).eval()

            def call(x):
                with torch.no_grad():
                    return ext(x)
            return call, 'reference', 'unmodified reference feature.py (LogmelIV_Extractor, torchaudio) from baseline/_ref, torch CPU fp32'
        except Exception as e:                                   # e.g. torchaudio missing on the box
            sys.stderr.write('reference module unusable (%s); timing the torch port\n' % (e,))
    from oracle import torch_port
    from pseldnets_b200 import filterbank as fbk
    win = fbk.make_window('hann', NFFT)
    fb = fbk.melscale_fbanks_htk_slaney(NFFT // 2 + 1, 20, SR / 2, NMELS, SR)
    return (lambda x: torch_port.logmel_iv(x, win, fb, NFFT, HOP)), 'port', 'torch CPU fp32 port of feature.py (oracle/torch_port.py)'


def cpu_baseline_throughput(batch, min_seconds, max_calls, threads):
    """The reference's CPU path on the host cores, bounded sample.  Returns (audio-s/s, calls, seconds, kind, what)."""
    import torch
    torch.set_num_threads(threads)
    call, kind, what = cpu_reference_callable()
    g = torch.Generator().manual_seed(1234)
    x = 0.1 * torch.randn(batch, C, L, generator=g)
    call(x)                                                # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < max_calls and (time.perf_counter() - t_start < min_seconds or len(times) < 3):
        t0 = time.perf_counter()
        call(x)
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return batch * CLIP_S / med, len(times), sum(times), kind, what


def cfg2_config(B, world):
    return {'workload': 'cfg2: FOA log-mel+IV, batch %d x 10 s x 4 ch @ 24 kHz per GPU -> (%d,7,1001,64)' % (B, B),
            'n_fft': NFFT, 'hop': HOP, 'n_mels': NMELS, 'per_gpu_batch': B, 'global_batch': B * world,
            'sharding': 'by clip, no collective on the data path',
            'l2': 'inputs 245.8 MB + outputs 114.8 MB per step exceed the 126 MB L2 (no flush needed)'}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (the unmodified
    feature.py from baseline/_ref when present, else the torch port), all host threads, on OUR arm's config: every
    step is one call on the full 64-clip batch.  Rank 0 only."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    call, kind, what = cpu_reference_callable()
    B = args.batch
    g = torch.Generator().manual_seed(1234)
    x = 0.1 * torch.randn(B, C, L, generator=g)
    for _ in range(args.warmup):
        call(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        call(x)
    dt = time.perf_counter() - t0
    value = args.steps * B * CLIP_S / dt
    # the same on an 8-clip sample (large CPU batches are allocator-bound: the smaller call is the reference's better case)
    x8 = x[:8].contiguous()
    call(x8)
    t8 = []
    for _ in range(5):
        t0 = time.perf_counter()
        call(x8)
        t8.append(time.perf_counter() - t0)
    # ONE host thread, one clip (B = 1), median of 5 after 2 warm-ups; and the host CPU
    torch.set_num_threads(1)
    x1 = x[:1].contiguous()
    t1 = []
    for i in range(7):
        t0 = time.perf_counter()
        call(x1)
        if i >= 2:
            t1.append(time.perf_counter() - t0)
    torch.set_num_threads(threads)
    cpu_model = None
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                cpu_model = line.split(':', 1)[1].strip()
                break
    except OSError:
        pass
    sample = 'every step = one call on the full batch of %d clips (10 s, 4 ch, 24 kHz); %s' % (B, what)
    emit_json(({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': cfg2_config(B, 1),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'sample_8_clips': {'value': 8 * CLIP_S / statistics.median(t8), 'unit': UNIT, 'cores': threads,
                           'sample': 'calls of 8 clips, median of 5'},
        'cfg1_one_thread': {'value': CLIP_S / statistics.median(t1), 'unit': UNIT, 'ms_per_clip': 1e3 * statistics.median(t1),
                            'cores': 1, 'sample': 'one 10-s clip (B = 1), median of 5'},
        'host_cpu': cpu_model,
    }))


def extra_record(workload, steps, warmup, dev, rank, world, dist=None):
    """cfg3 / cfg4 of BASELINE.json on this rank's GPU: resident-input time per step, roofline fraction and the SM
    clocks sampled DURING its own timed region.  Returns the record on rank 0 (None elsewhere)."""
    import torch
    import pseldnets_b200 as pb
    from pseldnets_b200 import _abi, shard
    if workload == 'cfg3':
        sr, hop, Lw, Cin, name = 24000, 240, 240000, 4, 'cfg3: MIC log-mel+GCC-PHAT, batch 64 x 10 s x 4 mics @ 24 kHz per GPU -> (64,10,1000,64)'
        ext = pb.get_afextractor({'data': dict(CFG['data'], audio_feature='logmelgcc')}).to(dev)
        nb = 64
        step = lambda x: ext(x)
        out_bytes = 10 * 1000 * 64 * 4
        scaling = 'weak'
    else:
        sr, hop, Lw, Cin, name = 32000, 320, 320000, 8, 'cfg4: L3DAS22 dual-FOA 8 ch @ 32 kHz, global batch 128 x 10 s sharded by clip -> (128,14,1001,64)'
        ext = pb.get_afextractor({'data': dict(CFG['data'], sample_rate=sr, hoplen=hop)}).to(dev)
        a, b = shard.clip_shard(128, rank, world)
        nb = b - a
        # two FOA arrays per clip: view (nb, 8, L) as (2 nb, 4, L) -> (2 nb, 7, T, 64) == (nb, 14, T, 64)
        step = lambda x: ext(x.view(2 * x.shape[0], 4, x.shape[2])).view(x.shape[0], 14, 1001, 64)
        out_bytes = 14 * 1001 * 64 * 4
        scaling = 'strong'
    g = torch.Generator(device=dev).manual_seed(1235 + rank)
    x = 0.1 * torch.randn(nb, Cin, Lw, device=dev, generator=g)
    for _ in range(max(warmup, 3)):
        y = step(x)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start()
        time.sleep(0.01)
        sampler.mark()
    l0 = _abi.lib().seld_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        y = step(x)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / steps
    finite = bool(torch.isfinite(y).all())
    del x, y
    if rank != 0:
        return None
    peak, peak_src = measured_peaks()
    n_global = nb * world if workload == 'cfg3' else 128
    algo = nb * (Cin * Lw * 4 + out_bytes)
    return {'metric': 'audio-seconds/sec (%s)' % workload, 'value': n_global * CLIP_S / (ms * 1e-3), 'unit': UNIT,
            'n_gpus': world, 'steps': steps, 'warmup': max(warmup, 3), 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': name},
            'roofline': {'bound': 'hbm', 'achieved': algo / (ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                         'frac': algo / (ms * 1e-3) / 1e9 / peak, 'traffic': None, 'peak_source': peak_src,
                         'algorithmic_bytes_per_step_per_gpu': algo},
            'gpu_launches': int(_abi.lib().seld_launch_count() - l0), 'clocks': clocks, 'outputs_finite': finite}


def run_extra(args):
    """--workload cfg3 / cfg4: the extra workloads of BASELINE.json as their own JSON line (any N)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)
    rec = extra_record(args.workload, args.steps, args.warmup, dev, rank, world, dist)
    if rank == 0:
        emit_json(rec)
    if world > 1:
        dist.destroy_process_group()


def run_epoch(args):
    """--workload cfg5: the whole epoch sweep as its own JSON line (any N)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=dev)
    rec = epoch_record(402000, args.batch, dev, rank, world, dist)
    if rank == 0:
        emit_json(rec)
    if world > 1:
        dist.destroy_process_group()


def epoch_record(n_chunks, B, dev, rank, world, dist=None):
    """cfg5 of BASELINE.json / SURVEY 8d: one epoch sweep of the reference's training set -- 67,000 one-minute clips
    = 402,000 ten-second chunks (configs/data/default.yaml:15-21; chunked before extraction, preprocess.py:464-521)
    -- sharded by chunk over the ranks, batches of 64 from a resident pool of 4 synthetic batches per rank (984 MB,
    cycled; generation excluded from the time).  Time = max over ranks for the whole sweep.  `n_chunks` < 402,000 is
    the SCALED sweep the default bench line carries (same code path, a fixed fraction of the epoch)."""
    import torch
    import pseldnets_b200 as pb
    from pseldnets_b200 import _abi, shard
    local_rank = dev.index or 0
    lo, hi = shard.clip_shard(n_chunks, rank, world)
    n_local = hi - lo
    steps = (n_local + B - 1) // B
    ext = pb.get_afextractor(CFG).to(dev)
    n_pool = 4
    g = torch.Generator(device=dev).manual_seed(1237 + rank)
    pool = [0.1 * torch.randn(B, C, L, device=dev, generator=g) for _ in range(n_pool)]
    first = [shard.clip_checksums(ext(p)) for p in pool]            # also the warm-up (>= 3 calls)
    last = [None] * n_pool
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.mark()
    l0 = _abi.lib().seld_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        nb = min(B, n_local - i * B)
        y = ext(pool[i % n_pool][:nb])
        if i >= steps - n_pool and nb == B:
            last[i % n_pool] = y
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = _abi.lib().seld_launch_count() - l0
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    ok = all(torch.equal(shard.clip_checksums(v), first[j]) for j, v in enumerate(last) if v is not None)
    okt = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        table = shard.gather_clip_checksums(last[(steps - 2) % n_pool] if steps > 1 else y, B * world)   # cross-rank check: NCCL all_gather
        finite = bool(torch.isfinite(table).all())
    else:
        finite = bool(torch.isfinite(first[0]).all())
    del pool, first, last
    if rank != 0:
        return None
    sec = float(t[0]) * 1e-3
    peak, peak_src = measured_peaks()
    algo = n_chunks * ALGO_BYTES_PER_CLIP
    full = n_chunks == 402000
    return {
        'metric': 'audio-seconds/sec (cfg5 epoch sweep%s)' % ('' if full else ', scaled'), 'value': n_chunks * CLIP_S / sec, 'unit': UNIT,
        'n_gpus': world, 'steps': steps, 'warmup': n_pool, 'ms_per_step': sec * 1e3 / steps, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'cfg5: epoch sweep, %d ten-second 4-ch chunks @ 24 kHz (%.3g M audio-s%s) sharded by chunk, '
                               'batches of %d from a resident pool of %d batches per rank (984 MB > L2)'
                               % (n_chunks, n_chunks * CLIP_S / 1e6, '' if full else ': %.4g of the 402000-chunk epoch' % (n_chunks / 402000.0), B, n_pool)},
        'sweep_seconds': sec, 'epoch_seconds_extrapolated': sec * 402000.0 / n_chunks,
        'roofline': {'bound': 'hbm', 'achieved': algo / sec / 1e9 / world, 'peak': peak, 'unit': 'GB/s',
                     'frac': algo / sec / 1e9 / world / peak, 'traffic': None, 'peak_source': peak_src,
                     'algorithmic_bytes_total': algo, 'note': 'per-GPU rate'},
        'gpu_launches': int(launches), 'clocks': clocks,
        'pool_results_reproduced_bitwise': bool(okt[0]), 'checksums_finite': finite}


def run_epilogue(args):
    """SURVEY 8f-1 (extra measurement, 1 GPU): the backbone-input stage on the cfg2 feature map
    (64, 7, 1001, 64) -> (64, 7, 256, 256): fused scalar + fold, the in-place scalar alone, and the
    reference's own op chain (accdoa.py:222-227 + htsat.py:493-511) run by torch on the same GPU."""
    import torch
    import pseldnets_b200 as pb
    from pseldnets_b200 import _abi
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    B, C, T, M, S = args.batch, 7, 1001, 64, 256
    g = torch.Generator(device=dev).manual_seed(1238)
    n_maps = 4                                   # 4 x 115 MB of input in rotation: each step reads a map that left L2
    maps = [40.0 * torch.rand(B, C, T, M, device=dev, generator=g) - 50.0 for _ in range(n_maps)]
    scalar = torch.nn.ModuleList([torch.nn.BatchNorm2d(M) for _ in range(C)]).to(dev).eval()
    for bn in scalar:
        bn.running_mean.copy_(30.0 * torch.rand(M, device=dev, generator=g) - 40.0)
        bn.running_var.copy_(50.0 * torch.rand(M, device=dev, generator=g) + 1.0)
    sp = pb.ScalarParams(scalar)

    def reference_chain(x):
        with torch.no_grad():
            x = x.transpose(1, 3)
            for nch in range(x.shape[-1]):
                x[..., [nch]] = scalar[nch](x[..., [nch]])
            x = x.transpose(1, 3)
            x = torch.nn.functional.pad(x, (0, 0, 0, 4 * S - T))
            x = x.permute(0, 1, 3, 2).contiguous()
            x = x.reshape(B, C, M, 4, S).permute(0, 1, 3, 2, 4).contiguous()
            return x.reshape(B, C, 4 * M, S)

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(maps[i % n_maps])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(maps[i % n_maps])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    warm = max(args.warmup, 3)
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    l0 = _abi.lib().seld_launch_count()
    sampler.mark()
    ms_fused = timed(lambda x: pb.scalar_wav2img(x, sp, S), args.steps, warm)
    n_launch = int(_abi.lib().seld_launch_count() - l0)
    ms_scalar = timed(lambda x: pb.apply_scalar(x, sp), args.steps, warm)
    clocks = sampler.stop()
    ms_ref = timed(reference_chain, max(3, args.steps // 10), 3)
    peak, peak_src = measured_peaks()
    in_b, out_b = B * C * T * M * 4, B * C * S * S * 4
    algo = in_b + out_b
    emit_json(({
        'metric': 'audio-seconds/sec (backbone-input stage: scalar + reshape_wav2img)', 'value': B * CLIP_S / (ms_fused * 1e-3),
        'unit': UNIT, 'n_gpus': 1, 'steps': args.steps, 'warmup': warm, 'ms_per_step': ms_fused, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': '8f-1: eval BatchNorm scalar + HTS-AT fold of the cfg2 feature map (%d,7,1001,64) -> (%d,7,256,256); '
                               'inputs rotate over %d maps (%d MB > L2)' % (B, B, n_maps, n_maps * in_b // 2**20)},
        'roofline': {'bound': 'hbm', 'achieved': algo / (ms_fused * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                     'frac': algo / (ms_fused * 1e-3) / 1e9 / peak, 'traffic': None, 'peak_source': peak_src,
                     'algorithmic_bytes_per_step': algo, 'kernel': 'seld::epi::scalar_wav2img_kernel'},
        'scalar_in_place': {'ms_per_step': ms_scalar, 'achieved_gbs': 2 * in_b / (ms_scalar * 1e-3) / 1e9,
                            'frac': 2 * in_b / (ms_scalar * 1e-3) / 1e9 / peak, 'kernel': 'seld::epi::scalar_kernel'},
        'torch_op_chain_same_gpu': {'ms_per_step': ms_ref, 'speedup': ms_ref / ms_fused,
                                    'what': 'accdoa.py:222-227 loop + htsat.py:493-511 in torch eager'},
        'gpu_launches': n_launch, 'clocks': clocks}))


def run_augment(args):
    """SURVEY 8f-4 (extra measurement, 1 GPU): Rotation + WavMix waveform work on the cfg2 batch
    (64, 4, 240000), ours (one launch each, in place) against the reference's own torch expressions
    (rotate.py:72 per clip, wavmix.py:50) on the same GPU.  Draws are fixed: p = 0.8 of the clips rotated,
    the 32 even clips mixed with a permutation of themselves."""
    import numpy as np
    import torch
    import pseldnets_b200.augment as aug
    from pseldnets_b200 import _abi
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    B, C, L = args.batch, 4, SR * CLIP_S
    g = torch.Generator(device=dev).manual_seed(1239)
    x = 0.1 * torch.randn(B, C, L, device=dev, generator=g)
    rng = np.random.default_rng(7)
    table = list(aug.TRANS_48.values())
    rot = [(table[rng.integers(6)], rng.choice([-1, 1], size=3)) if rng.random() < 0.8 else None for _ in range(B)]
    codes = torch.tensor([aug.ROT_IDENTITY if r is None else aug.rotation_code(r[0], (r[1][1], r[1][2], r[1][0])) for r in rot],
                         dtype=torch.int32, device=dev)
    dst = np.arange(0, B, 2)
    src = rng.permutation(dst)
    lam = torch.from_numpy(rng.beta(0.5, 0.5, size=len(dst)).astype(np.float32)).to(dev)
    dst_t, src_t = torch.from_numpy(dst).to(dev), torch.from_numpy(src).to(dev)

    def ours(x):
        aug.rotate_waveforms(x, codes)
        aug.wavmix_waveforms(x, dst, src, lam)

    def reference(x):
        for n, r in enumerate(rot):                       # rotate.py:13-36, waveform part
            if r is None:
                continue
            (s_x, s_y, s_z), (sx, sy, sz) = r
            d = x[n]
            x[n] = torch.stack((d[0], sy * d[s_x], sz * d[s_y], sx * d[s_z]), axis=0)
        lx = lam.reshape(-1, 1, 1)
        x[dst_t] = lx * x[dst_t] + (1. - lx) * x[src_t]   # wavmix.py:50

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn(x)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    warm = max(args.warmup, 3)
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    l0 = _abi.lib().seld_launch_count()
    sampler.mark()
    ms = timed(ours, args.steps, warm)
    n_launch = int(_abi.lib().seld_launch_count() - l0)
    clocks = sampler.stop()
    ms_ref = timed(reference, max(3, args.steps // 10), 3)
    n_rot = sum(r is not None for r in rot)
    algo = n_rot * 3 * L * 4 * 2 + len(dst) * C * L * 4 * 2     # rotated: 3 channels read + written; mixed (cycles): read + written once
    peak, peak_src = measured_peaks()
    emit_json(({
        'metric': 'audio-seconds/sec (waveform augmentation: Rotation + WavMix)', 'value': B * CLIP_S / (ms * 1e-3), 'unit': UNIT,
        'n_gpus': 1, 'steps': args.steps, 'warmup': warm, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': '8f-4: FOA Rotation (%d of %d clips) + WavMix (%d clips, cyclic) in place on the cfg2 batch (%d,4,240000); '
                               'batch 246 MB > L2' % (n_rot, B, len(dst), B)},
        'roofline': {'bound': 'hbm', 'achieved': algo / (ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                     'frac': algo / (ms * 1e-3) / 1e9 / peak, 'traffic': None, 'peak_source': peak_src,
                     'algorithmic_bytes_per_step': algo, 'kernel': 'seld::aug::foa_rotate_kernel + seld::aug::wavmix_kernel'},
        'torch_expressions_same_gpu': {'ms_per_step': ms_ref, 'speedup': ms_ref / ms,
                                       'what': 'rotate.py:72 per clip + wavmix.py:50 in torch eager'},
        'gpu_launches': n_launch, 'clocks': clocks}))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import pseldnets_b200 as pb
    from pseldnets_b200 import _abi

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the hot path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    bind_to_gpu_numa_node(local_rank)       # pinned host buffers (e2e) then live on the GPU's own NUMA node
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')    # keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=dev)
    B = args.batch
    ext = pb.get_afextractor(CFG).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # synthetic clips, resident in HBM (each rank its own shard of the global batch: weak scaling)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = 0.1 * torch.randn(B, C, L, device=dev, generator=g)
    y = None
    sampler = ClockSampler(local_rank)          # samples from the warm-up to the end of the timed steps
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        y = ext(x)
    barrier()

    # ---- timed region: K steps, device timing on the launching (current) stream
    launches0 = _abi.lib().seld_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    if rank == 0:
        sampler.mark()
    ev[0].record()
    for i in range(args.steps):
        y = ext(x)
        ev[i + 1].record()
    barrier()
    launches = _abi.lib().seld_launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public API with HOST buffers: every step copies that step's
    # inputs from pinned host memory, runs the kernels and reads the whole feature map back into
    # pinned host memory (LogmelIV_Extractor.forward_host -> seld_logmel_iv_f32_host: chunked,
    # three streams, H2D / kernel / D2H of neighbouring chunks overlap).
    xh = x.cpu().pin_memory()
    yh = torch.empty(y.shape, dtype=y.dtype).pin_memory()
    for _ in range(2):
        ext.forward_host(xh, out=yh, device=dev, chunk_clips=args.e2e_chunk, synchronize=False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 10))
    e0.record()
    for _ in range(e2e_steps):
        ext.forward_host(xh, out=yh, device=dev, chunk_clips=args.e2e_chunk, synchronize=False)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    e2e_ok = bool(torch.equal(yh[:2], y[:2].cpu()))          # host path returns the same bits as the resident path

    # ---- the roofline of e2e: the step's copies alone -- the same 245.76 MB host->device and 114.8 MB device->host,
    # from / to the same pinned buffers, on two streams at once (PCIe is full duplex), all ranks at the same time
    xd = torch.empty_like(x)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def copies(n):
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        s_in.wait_stream(torch.cuda.current_stream(dev)); s_out.wait_stream(torch.cuda.current_stream(dev))
        for _ in range(n):
            with torch.cuda.stream(s_in):
                xd.copy_(xh, non_blocking=True)
            with torch.cuda.stream(s_out):
                yh.copy_(y, non_blocking=True)
        torch.cuda.current_stream(dev).wait_stream(s_in); torch.cuda.current_stream(dev).wait_stream(s_out)
        c1.record()
        barrier()
        return c0.elapsed_time(c1) / n
    copies(1)
    copy_ms = copies(e2e_steps)
    del xd

    # ---- and the same pipeline fed with 16-bit PCM (what a wav / flac decoder yields): half the host->device bytes
    xi = (x * 32767.0).round().clamp(-32768, 32767).to(torch.int16).cpu().pin_memory()
    for _ in range(2):
        ext.forward_host(xi, out=yh, device=dev, chunk_clips=args.e2e_chunk, synchronize=False)
    barrier()
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record()
    for _ in range(e2e_steps):
        ext.forward_host(xi, out=yh, device=dev, chunk_clips=args.e2e_chunk, synchronize=False)
    i1.record()
    barrier()
    e2e_i16_ms = i0.elapsed_time(i1)
    del xi

    t = torch.tensor([total_ms, e2e_ms, copy_ms, e2e_i16_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # result check only (after timing): one gather of per-clip checksums, 16 B per clip, over NCCL
    from pseldnets_b200 import shard
    table = shard.gather_clip_checksums(y, world * B)
    finite = bool(torch.isfinite(table).all()) and table.shape[0] == world * B
    total_ms, e2e_ms, copy_ms, e2e_i16_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])

    if rank == 0:
        peak, peak_src = measured_peaks()
        audio_s = world * B * CLIP_S
        value = audio_s * args.steps / (total_ms * 1e-3)
        launch_ms = statistics.mean(per_launch_ms)
        achieved = B * ALGO_BYTES_PER_CLIP / (launch_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get('foa_iv2_kernel_bytes_per_launch')
            except Exception:
                traffic = None
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))       # the CPU baseline gets every host core back
        except Exception:
            pass
        cpu_threads = os.cpu_count() or 1
        cpu_val, cpu_calls, cpu_secs, cpu_kind, cpu_what = cpu_baseline_throughput(8, args.cpu_seconds, 200, cpu_threads)
        # FP32 side of the roofline: algorithmic flops (FFT 1.0e8 + pointwise 2.3e7 + sparse mel 1.4e7 per clip)
        # against the FP32 FMA rate measured here with a cuBLAS SGEMM (TF32 off), next to the nominal 148 SM x 128 lanes x 2
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        ga = torch.randn(8192, 8192, device=dev); gb = torch.randn(8192, 8192, device=dev)
        torch.matmul(ga, gb)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(3):
            torch.matmul(ga, gb)
        f1.record()
        torch.cuda.synchronize()
        fp32_peak = 3 * 2 * 8192.0 ** 3 / (f0.elapsed_time(f1) * 1e-3) / 1e12
        torch.backends.cuda.matmul.allow_tf32 = tf32
        del ga, gb
        flops_per_clip = 1.0e8 + 2.3e7 + 1.4e7
        fp32_achieved = B * flops_per_clip / (launch_ms * 1e-3) / 1e12
        library = None
        if world == 1 and args.cpu_seconds > 0:
            # the reference's op sequence (torch.stft -> |X|^2 -> matmul -> log10, IV arithmetic: cuFFT, cuBLAS
            # and ~45 ATen launches) on the SAME GPU and batch -- the library baseline the fused kernel replaces
            win_d, fb_d = ext.stft_extractor.window, ext.mel_scale.fb
            with torch.no_grad():
                for _ in range(2):
                    y_lib = library_ops_same_gpu(x, win_d, fb_d)
            torch.cuda.synchronize()
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            with torch.no_grad():
                for _ in range(5):
                    y_lib = library_ops_same_gpu(x, win_d, fb_d)
            l1.record()
            torch.cuda.synchronize()
            lib_ms = l0.elapsed_time(l1) / 5
            library = {'ms_per_step': lib_ms, 'value': audio_s / (lib_ms * 1e-3), 'unit': UNIT,
                       'speedup_of_value': lib_ms / (total_ms / args.steps),
                       'max_abs_diff_vs_ours': float((y_lib - y).abs().max()),
                       'what': 'feature.py op sequence in torch eager on the same B200 (cuFFT / cuBLAS / ATen), same resident batch'}
            del y_lib
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': cfg2_config(B, world),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'peak_source': peak_src, 'kernel': 'seld::foa_iv2_kernel<8,float,true,false,true> (item form of the mel step; one launch per step)',
                         'algorithmic_bytes_per_launch': B * ALGO_BYTES_PER_CLIP, 'launch_ms': launch_ms,
                         'fp32_frac': fp32_achieved / fp32_peak, 'fp32_achieved_tflops': fp32_achieved,
                         'fp32_peak_tflops': fp32_peak, 'fp32_peak_source': 'measured here: cuBLAS SGEMM 8192^3, TF32 off (nominal 148 x 128 x 2 x 1.965 GHz = 74.4)',
                         'algorithmic_flops_per_launch': B * flops_per_clip},
            'cpu_baseline': {'value': cpu_val, 'unit': UNIT, 'cores': cpu_threads, 'kind': cpu_kind,
                             'sample': '%d calls of 8 clips (10 s, 4 ch, 24 kHz) in %.1f s; %s' % (cpu_calls, cpu_secs, cpu_what)},
            'e2e': {'value': audio_s * e2e_steps / (e2e_ms * 1e-3), 'unit': UNIT,
                    'h2d_bytes_per_step': B * C * L * 4, 'd2h_bytes_per_step': B * (C + 3) * T * NMELS * 4,
                    'steps': e2e_steps, 'matches_resident_path': e2e_ok,
                    'path': 'pinned host -> LogmelIV_Extractor.forward_host (seld_logmel_iv_f32_host: %d-clip chunks, H2D / kernel / D2H on 3 streams) -> pinned host' % args.e2e_chunk,
                    # roofline of e2e: the two copies of a step on their own, concurrently, all ranks at once
                    'copies_only_ms_per_step': copy_ms, 'ms_per_step': e2e_ms / e2e_steps, 'frac': copy_ms / (e2e_ms / e2e_steps),
                    'copies_only_h2d_gbs': B * C * L * 4 / (copy_ms * 1e-3) / 1e9,
                    'bound': 'host<->device copies (PCIe / host memory), measured in this run',
                    'int16_input': {'value': audio_s * e2e_steps / (e2e_i16_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': B * C * L * 2,
                                    'what': 'same pipeline fed with 16-bit PCM (seld_logmel_iv_i16_host), same features bit for bit'}},
            'gpu_launches': int(launches), 'clocks': clocks, 'outputs_finite': finite,
        }
        if library is not None:
            out['library_baseline_same_gpu'] = library
    # ---- the other workloads of BASELINE.json, each with its own time, roofline fraction and clock record (N = 1; at
    # N > 1 run them with --workload cfg3|cfg4|cfg5 under the same launcher)
    if world == 1 and not args.no_extra:
        del x, y, xh, yh
        torch.cuda.empty_cache()
        extra = {}
        for w in ('cfg3', 'cfg4'):
            extra[w] = extra_record(w, 30, 5, dev, rank, world)
        extra['cfg5_scaled'] = epoch_record(4020, 64, dev, rank, world)
        for rec in extra.values():
            for k in ('higher_is_better', 'vs_baseline', 'dtype', 'data', 'unit', 'n_gpus', 'scaling'):
                rec.pop(k, None)
        out['extra'] = extra
    if rank == 0:
        emit_json((out))
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries below us may print there too (NCCL's version banner does
    on some hosts), so the real stdout is set aside for the JSON line and file descriptor 1 is pointed at stderr."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit_json(obj):
    line = (json.dumps(obj) + '\n').encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=400)
    ap.add_argument('--warmup', type=int, default=100)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='clips per GPU per step (cfg2: 64)')
    ap.add_argument('--workload', default='cfg2', choices=['cfg2', 'cfg3', 'cfg4', 'cfg5', 'wav2img', 'augment'],
                    help='cfg2 = BASELINE metric (default); cfg3 = MIC log-mel+GCC B=64; cfg4 = L3DAS22 dual-FOA '
                         '8 ch 32 kHz, global batch 128 sharded by clip (extra measurements, not the headline)')
    ap.add_argument('--cpu-seconds', type=float, default=10.0, help='bound on the cpu_baseline sample')
    ap.add_argument('--e2e-chunk', type=int, default=4, help='clips per chunk of the host-buffer pipeline (e2e)')
    ap.add_argument('--no-extra', action='store_true', help='skip the cfg3 / cfg4 / cfg5-scaled sub-records of the default line')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload == 'cfg5':
        run_epoch(args)
    elif args.workload == 'wav2img':
        run_epilogue(args)
    elif args.workload == 'augment':
        run_augment(args)
    elif args.workload != 'cfg2':
        run_extra(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
