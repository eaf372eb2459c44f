#!/usr/bin/env python
"""Benchmark of the CP-ALS hot path (BASELINE.json: "MTTKRP GB/s (% HBM roofline) and CP-ALS iters/sec at 1/2/4/8 B200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload dense|sparse|cfg3|cfg5]

One "step" is one full CP-ALS iteration on the workload -- the MTTKRP of every mode (two dimension-tree passes over
the tensor plus the small second-stage reductions), the Hadamard-of-Grams pseudo-inverse, the factor
update/normalisation, and the fit -- i.e. the loop body of the reference's cp.als (sktensor/cp.py:138-169).

Workload at N = 1: BASELINE.json configs[1], "cp_als dense 3-way 512x512x512 rank 32, FP64".  For N > 1 (launched by
torch.distributed.run, one rank per GPU, NCCL) the tensor is (512 N) x 512 x 512, cut along mode 0 so that every GPU
keeps a 512^3 slab: weak scaling.

`value` (GB/s) = algorithmic MTTKRP bytes of one iteration -- the reference's three MTTKRPs, sum over modes of
8 (prod I + R sum I), summed over all ranks' slabs -- divided by the time of the WHOLE iteration (solves and fit
included), tensor resident in HBM.  `als_iters_per_sec` is the same clock.  `e2e` is the same metric through the
public API, cp_als(dtensor(host ndarray), ...): host tensor upload, initial factor upload, per-iteration fit read-back
and final factor download all inside the timed region (median of three whole calls).  `roofline` describes the
dominant kernel (the dense MTTKRP pass), timed with CUDA events around each of its launches inside the timed
iterations; at rank 32 it is bound by the FP64 tensor pipe, whose peak is probed live (MEASURED_PEAKS.json has no
FP64 figure), and the HBM fraction is reported beside it.

`--impl reference` times the reference's own algorithm (the NumPy restatement in oracle/, the reference being pure
Python that cannot travel to the GPU box) on the host cores, on a bounded 256^3 sub-cube of the same workload,
reported in the same unit.  `--workload cfg3|cfg5` run the sharded sparse / 4-way configurations of BASELINE.json
(secondary lines, see run_config); `--workload sparse` times the three sparse MTTKRPs alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, 'scikit-tensor_b200'), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = 'cp_als_mttkrp_GBps'
UNIT = 'GB/s'
SIDE, RANK = 512, 32                       # BASELINE.json configs[1]
SAMPLE_SIDE = 256                          # CPU sample: 1/8 of the tensor
SPARSE_SHAPE = (10000000, 1000000, 100000)  # BASELINE.json configs[2]


def alg_bytes_dense(shape, R):
    """Algorithmic bytes of one dense MTTKRP: X once + the N-1 factors + the output."""
    return 8.0 * (float(np.prod([float(s) for s in shape])) + R * float(sum(shape)))


def alg_bytes_sparse(shape, nnz, R):
    """int32 subscripts + FP64 value per non-zero, every factor row once, output once."""
    return float(nnz) * (4 * len(shape) + 8) + 8.0 * R * float(sum(shape))


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thr = index, [], None, None
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '25'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        inside = [r for (t, r) in self.rows if self.t_begin is not None and self.t_begin <= t <= self.t_end + 0.05]
        window = 'timed region'
        if len(inside) < 3:      # a very short timed region: fall back to every sample taken under load
            inside = [r for (t, r) in self.rows]
            window = 'warm-up + timed region (timed region shorter than 3 sampling periods)'
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(names, r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'window': window, 'reasons': sorted(reasons)}


def cpu_sample(steps=1):
    """The reference algorithm (oracle port) on a 256^3 sub-cube: seconds per ALS iteration."""
    from oracle import sktensor_oracle as orc
    rng = np.random.RandomState(0)
    X = rng.rand(SAMPLE_SIDE, SAMPLE_SIDE, SAMPLE_SIDE)
    np.random.seed(1)
    # all the host threads BLAS can use (torchrun exports OMP_NUM_THREADS=1 for N > 1: lift that for this leg)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        with threadpool_limits(limits=os.cpu_count()):
            _, _, _, _, ex = orc.cp_als(X, RANK, init='random', max_iter=steps, conv=0)
            cores = max([p.get('num_threads', 1) for p in threadpool_info()] + [1])
    except ImportError:
        _, _, _, _, ex = orc.cp_als(X, RANK, init='random', max_iter=steps, conv=0)
        cores = os.cpu_count()
    return np.asarray(ex), cores


def cpu_baseline_full():
    """cpu_baseline of the GPU arm: 2 iterations of the unmodified reference on the full 512^3 workload (~25 s of host
    time), run as `bench.py --impl reference` in its own process (the reference's package is also called `sktensor`);
    the oracle port on a 256^3 sample when that fails."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '2', '--warmup', '0'],
                           capture_output=True, text=True, timeout=900,
                           env={k: v for k, v in os.environ.items() if k not in ('OMP_NUM_THREADS', 'RANK', 'WORLD_SIZE')})
        lines = [l for l in r.stdout.splitlines() if l.strip().startswith('{')]
        cb = json.loads(lines[-1])['cpu_baseline']
        if cb.get('value'):
            return cb
    except Exception:
        pass
    ex, cores = cpu_sample(2)
    return cpu_baseline_obj(ex, cores)


def cpu_baseline_obj(ex, cores):
    per = float(np.mean(ex))
    b = 3 * alg_bytes_dense((SAMPLE_SIDE,) * 3, RANK)
    return {'value': b / per / 1e9, 'unit': UNIT, 'cores': int(cores), 'kind': 'port',
            'als_iters_per_sec_on_sample': 1.0 / per,
            'sample': 'oracle/sktensor_oracle.cp_als (NumPy restatement of sktensor/cp.py:46-171), %d iteration(s) '
                      'on a %d^3 rank-%d sub-cube (1/8 of the 512^3 workload); GB/s uses the same algorithmic '
                      'byte count per iteration; BLAS threads = cores, the rest of the algorithm is single-threaded'
                      % (len(ex), SAMPLE_SIDE, RANK)}


def reference_sample(side, iters):
    """The UNMODIFIED reference (baseline/_ref/sktensor, imported through the alias-only shim oracle/ref_shim.py) on a
    side^3 rank-32 tensor of the bench's generator: wall seconds per ALS iteration.  Returns None when baseline/_ref is
    absent (it is installed by __graft_entry__.build() where /root/reference exists and travels to the GPU box)."""
    ref_root = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_root, 'sktensor')):
        return None
    import warnings
    warnings.simplefilter('ignore')
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ref_shim
    ref_shim.import_reference(ref_root)
    time.clock = time.perf_counter            # the reference brackets every iteration with time.clock() (cp.py:139,163)
    from sktensor import cp_als as ref_cp_als, dtensor as ref_dtensor
    rng = np.random.RandomState(1000)         # rank 0's slab of the GPU arm
    X = rng.rand(side, side, side)
    np.random.seed(1)

    def call():
        return ref_cp_als(ref_dtensor(X), RANK, init='random', max_iter=iters, conv=0)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        with threadpool_limits(limits=os.cpu_count()):
            P, fit, itr, ex = call()
            cores = max([p.get('num_threads', 1) for p in threadpool_info()] + [1])
    except ImportError:
        P, fit, itr, ex = call()
        cores = os.cpu_count()
    return np.asarray(ex, dtype=float), int(cores), float(np.ravel(fit)[0])


def run_reference(args):
    """The reference arm: the unmodified reference's cp_als on the host cores, on the GPU arm's full 512^3 rank-32
    configuration and step counts (capped so that the run ends within a few minutes: an iteration takes ~11 s)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    warm = max(0, min(args.warmup, 5))
    steps = max(1, min(args.steps, 20))
    side = int(os.environ.get('SKT_BENCH_REF_SIDE', SIDE))     # (the CPU contract test shrinks the tensor)
    got = reference_sample(side, warm + steps)
    if got is not None:
        ex, cores, fit = got
        ex = ex[warm:]
        per = float(np.mean(ex))
        b = 3 * alg_bytes_dense((side,) * 3, RANK)
        cb = {'value': b / per / 1e9, 'unit': UNIT, 'cores': cores, 'kind': 'reference',
              'als_iters_per_sec_on_sample': 1.0 / per,
              'sample': 'the unmodified reference (baseline/_ref/sktensor via oracle/ref_shim.py): sktensor.cp_als('
                        "dtensor(X), 32, init='random', max_iter=%d, conv=0) on the full %d^3 tensor; per-iteration wall "
                        'time from its own exectimes (time.clock aliased to perf_counter), first %d iteration(s) dropped as '
                        'warm-up; BLAS threads = cores, the rest of the reference is single-threaded Python'
                        % (warm + steps, side, warm)}
        note = 'full configuration, unmodified reference' if side == SIDE else 'SKT_BENCH_REF_SIDE=%d' % side
        its = 1.0 / per
    else:
        ex, cores = cpu_sample(min(warm, 1) + min(steps, 8))
        warm = min(warm, 1)
        ex = ex[warm:]
        cb = cpu_baseline_obj(ex, cores)
        note = 'baseline/_ref absent: CPU arm runs the oracle port on a bounded 256^3 sample; GB/s is size-normalised'
        its = cb['als_iters_per_sec_on_sample'] / 8.0
        fit = None
    line = {
        'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'impl': 'reference', 'n_gpus': args.gpus,
        'steps': int(len(ex)), 'warmup': int(warm), 'requested_steps': args.steps, 'requested_warmup': args.warmup,
        'ms_per_step': float(np.mean(ex)) * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'cp_als dense 3-way 512x512x512 rank 32 FP64 (BASELINE.json configs[1])',
                   'shape': [SIDE, SIDE, SIDE] if got is None else [side] * 3, 'rank': RANK, 'init': 'random',
                   'fit_method': 'full', 'note': note},
        'als_iters_per_sec': its, 'final_fit': fit,
        'cpu_baseline': cb,
        'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def run_gpu(args):
    # keep stdout clean for the single JSON line: libraries (NCCL prints its version banner) go to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from sktensor import _lib, _parallel, cp_als, dtensor
    from sktensor import _device as dev

    lib = _lib.load()
    R = RANK
    gshape = (SIDE * world, SIDE, SIDE)
    lshape = (SIDE, SIDE, SIDE)
    offsets = np.arange(world + 1, dtype=np.int64) * SIDE
    rng = np.random.RandomState(1000 + rank)
    # this rank's slab in page-locked host memory (the e2e contract: "host->device copy ... from pinned host memory"):
    # skt_upload then DMAs it in place.  SKT_BENCH_PAGEABLE=1 uses an ordinary NumPy array (multi-threaded staging).
    if os.environ.get('SKT_BENCH_PAGEABLE', '0') == '1':
        Xh = rng.rand(*lshape)
    else:
        Xpin = torch.empty(lshape, dtype=torch.float64, pin_memory=True)
        Xh = Xpin.numpy()
        Xh[...] = rng.rand(*lshape)

    def make_input():
        if world == 1:
            return dtensor(Xh)
        return _parallel.LocalShard(dtensor(Xh), 0, gshape, offsets)

    def init_factors():
        np.random.seed(1)                                    # same stream on every rank (cp.py:197-199)
        return [None] + [np.random.rand(gshape[n], R) for n in range(1, 3)]

    # ---- device-resident arm: the engine behind cp_als, X uploaded once, outside the timed region
    class TimedBackend(_parallel.CudaBackend):
        def __init__(self):
            _parallel.CudaBackend.__init__(self)
            self.events, self.record = [], False

        def mttkrp_dense(self, Xd, shape, U, R_, n, out):
            if not self.record:
                return _parallel.CudaBackend.mttkrp_dense(self, Xd, shape, U, R_, n, out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = _parallel.CudaBackend.mttkrp_dense(self, Xd, shape, U, R_, n, out)
            e1.record()
            self.events.append((n, e0, e1))
            return r

        def planes_ttm(self, planes, mode, U, R_, out):
            if not self.record:
                return _parallel.CudaBackend.planes_ttm(self, planes, mode, U, R_, out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = _parallel.CudaBackend.planes_ttm(self, planes, mode, U, R_, out)
            e1.record()
            self.events.append((mode, e0, e1))
            return r

    be = TimedBackend()
    eng = _parallel.CpAlsEngine(make_input(), R, init_factors(), backend=be)
    W = max(args.warmup, 3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)          # let nvidia-smi come up so that samples exist from the warm-up on
    # device / host pre-conditioning: the first process on a fresh box runs its first ~0.2 s of sweeps 5 % slower (power state,
    # cold host path: 0.728 ms per step with 5 warm-up steps against 0.688 from the second process on, or with 300 warm-up
    # steps).  The same count on every rank (the sweeps contain collectives); then the W warm-up steps the caller asked for.
    PRE = max(0, int(os.environ.get('SKT_BENCH_PRECONDITION', '300')) - W)
    for w in range(PRE + W):
        eng.sweep(first_iter=(w == 0))
        eng.begin_read()
        eng.prefetch_next_sweep()
        eng.read_fit()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    be.record = True
    launches0 = lib.skt_launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    sampler.mark_begin()
    t0.record()
    fit = 0.0
    for _ in range(args.steps):              # the loop body of sktensor.cp.als
        eng.sweep(first_iter=False)
        eng.begin_read()
        eng.prefetch_next_sweep()
        fit = eng.read_fit()
    t1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    launches = lib.skt_launch_count() - launches0
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    be.record = False
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    iter_bytes = sum(alg_bytes_dense(gshape, R) for _ in range(3))      # whole job, all ranks' slabs: the 3 MTTKRPs of cp.py:142-143
    value = iter_bytes / (ms_step * 1e-3) / 1e9

    # ---- dominant kernel: per-launch durations of the dense MTTKRP inside the timed iterations
    durs = {}
    for n, e0, e1 in be.events:
        durs.setdefault(n, []).append(e0.elapsed_time(e1))
    all_d = [d for v in durs.values() for d in v]
    kbytes = alg_bytes_dense(lshape, R)                                   # one launch handles one slab
    kern_ms = float(np.mean(all_d))
    hbm_peak, peak_src = measured_peaks()
    dmma_tf, dfma_tf = (0.0, 0.0)
    if rank == 0:
        from sktensor import _kernels as K
        dmma_tf, dfma_tf = K.probe_fp64_peaks()
    kflops = 2.0 * R * float(np.prod([float(s) for s in lshape]))
    t_roof = max(kbytes / (hbm_peak * 1e9), kflops / (dmma_tf * 1e12)) if dmma_tf else kbytes / (hbm_peak * 1e9)
    traffic = None
    prof = os.path.join(ROOT, 'profiles', 'dense_mttkrp_traffic.json')
    if os.path.exists(prof):
        with open(prof) as f:
            traffic = json.load(f).get('dram_bytes_per_launch')
    ach_tf = kflops / (kern_ms * 1e-3) / 1e12
    ach_gb = kbytes / (kern_ms * 1e-3) / 1e9
    roofline = {
        'kernel': 'mttkrp_dense_tma_kernel (dense FP64 MTTKRP pass over the 512^3 slab: TMA-staged tiles, FP64 '
                  'tensor-core MMA; two launches per iteration -- the dimension-tree pass that serves modes 0 and 1, '
                  'and the mode-2 MTTKRP)',
        # rank 32 -> 8 flop per algorithmic byte, above the FP64 ridge (peak_dmma / peak_hbm ~ 5.7 flop/B): the FP64
        # tensor pipe is the binding roof.  MEASURED_PEAKS.json carries no FP64 figure (its tensor number is bf16),
        # so the denominator is measured live by skt_probe_fp64_peaks (DMMA issue rate; cuBLAS DGEMM 8192^3 reaches
        # 35.5 TFLOP/s on the same parts, profiles/r1_quick_perf.log).
        'bound': 'tensor', 'achieved': ach_tf, 'peak': dmma_tf, 'unit': 'TFLOP/s',
        'frac': (ach_tf / dmma_tf) if dmma_tf else None, 'traffic': traffic,
        'peak_source': 'FP64 DMMA peak measured live by bench.py (skt_probe_fp64_peaks); MEASURED_PEAKS.json has no FP64 entry',
        'alg_flops_per_launch': kflops, 'alg_bytes_per_launch': kbytes, 'avg_launch_ms': kern_ms,
        'per_pass_ms': {('tree pass A (modes 0,1)' if n == 0 else 'mode %d' % n): float(np.mean(v))
                        for n, v in sorted(durs.items())},
        'launches_timed': len(all_d),
        'share_of_step': float(sum(all_d)) / ms_total if world == 1 else None,
        'hbm': {'achieved': ach_gb, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach_gb / hbm_peak, 'peak_source': peak_src,
                'note': 'algorithmic bytes 8(prod I + R sum I) per launch over the same launch time'},
        'peak_tflops_dfma': dfma_tf,
        'frac_of_binding_roofline': t_roof / (kern_ms * 1e-3),
    }
    del eng
    torch.cuda.empty_cache()

    # ---- end-to-end arm: the public call a user makes, host buffers in, host factors out
    K2 = max(2, min(args.steps, 10))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_runs = []
    for rep in range(3):                      # every run is the whole call (upload included); the median is reported
        if world > 1:
            dist.barrier()
        np.random.seed(1)
        tic = time.perf_counter()
        P, fit2, itr, _ = cp_als(make_input(), R, init='random', max_iter=K2, conv=0)
        torch.cuda.synchronize()
        e2e_runs.append(time.perf_counter() - tic)
    e2e_s = float(np.median(e2e_runs))
    e2e = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e.item())
    h2d = (Xh.nbytes + 8 * R * (gshape[1] + gshape[2])) / float(K2)
    d2h = (8 * R * sum(gshape) + 8 * R) / float(K2) + 16 + 4 * 3
    e2e_obj = {'value': iter_bytes * K2 / e2e_s / 1e9, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': d2h, 'als_iters_per_sec': K2 / e2e_s, 'iterations': K2,
               'call': "cp_als(dtensor(X_host), 32, init='random', max_iter=%d, conv=0)" % K2,
               'runs_s': [float(x) for x in e2e_runs],
               'note': 'median of 3 whole calls, each: tensor upload from %s host memory (once per call), factor upload, per-iteration ' % ('pageable' if os.environ.get('SKT_BENCH_PAGEABLE', '0') == '1' else 'page-locked') +
                       'fit read-back, final factor download; bytes are per iteration (upload amortised over '
                       '%d iterations)' % K2}

    # ---- opt-in arm (not the headline): the same workload with both tree passes as sliced contractions on the INT8
    # tensor cores (SKT_DENSE_I8=6: 48-bit fixed point per operand, csrc/slice_gemm.cu)
    sliced = None
    if not args.no_extras:
        os.environ['SKT_DENSE_I8'] = '6'
        try:
            be2 = TimedBackend()
            eng2 = _parallel.CpAlsEngine(make_input(), R, init_factors(), backend=be2)
            if eng2.i8 is not None:
                plane_bytes = eng2.i8.device_bytes()
                be2.record = True
                ms2, fit2_ = timed_sweeps(torch, dist, world, eng2, args.steps, PRE + W)     # same iteration count as the default arm
                be2.record = False
                ev = be2.events[-2 * args.steps:]
                pass_ms = {}
                for mode, e0, e1 in ev:
                    pass_ms.setdefault(mode, []).append(e0.elapsed_time(e1))
                eng2.release()
                del eng2
                torch.cuda.empty_cache()
                runs = []
                for rep in range(3):
                    if world > 1:
                        dist.barrier()
                    np.random.seed(1)
                    tic = time.perf_counter()
                    cp_als(make_input(), R, init='random', max_iter=K2, conv=0)
                    torch.cuda.synchronize()
                    runs.append(time.perf_counter() - tic)
                e2 = torch.tensor([float(np.median(runs))], dtype=torch.float64, device='cuda')
                if world > 1:
                    dist.all_reduce(e2, op=dist.ReduceOp.MAX)
                avg_pass = float(np.mean([d for v in pass_ms.values() for d in v]))
                nelem = float(np.prod([float(d) for d in lshape]))
                sliced = {
                    'how': 'SKT_DENSE_I8=6 (opt-in; the headline above is the default FP64 DMMA path)',
                    'ms_per_step': ms2, 'als_iters_per_sec': 1e3 / ms2, 'value_GBps': iter_bytes / (ms2 * 1e-3) / 1e9,
                    'speedup_vs_default': ms_step / ms2, 'final_fit': fit2_, 'fit_delta_vs_default': abs(fit2_ - fit),
                    'per_pass_ms': {('last-mode contraction (serves modes 0,1)' if m == 2 else
                                     'first-mode contraction (+ second-stage reduction -> mode 2)'): float(np.mean(v))
                                    for m, v in sorted(pass_ms.items())},
                    'plane_bytes_per_gpu': plane_bytes,
                    'e2e_als_iters_per_sec': K2 / float(e2.item()),
                    'roofline': {'kernel': 'slice_gemm_kernel (tcgen05.mma.kind::i8, INT32 accumulators in tensor memory, TMA-fed; '
                                           'streams 6 byte planes per element instead of 8 bytes of FP64)',
                                 'bound': 'hbm', 'achieved': 6.0 * nelem / (avg_pass * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                                 'frac': 6.0 * nelem / (avg_pass * 1e-3) / 1e9 / hbm_peak, 'avg_launch_ms': avg_pass,
                                 'note': 'per launch: the factor slicing kernel + the contraction; bytes = 6 per tensor element'},
                    'accuracy': 'max|dT|/max|T| 4e-15 against the FP64 kernel on this tensor (tools/slices_perf.py); tests/test_gpu_slices.py',
                }
        except Exception as e:
            sliced = {'error': '%s: %s' % (type(e).__name__, e)}
            if world > 1:
                raise
        finally:
            os.environ.pop('SKT_DENSE_I8', None)

    extras = {}
    if sliced is not None:
        extras['opt_in_int8_slices'] = sliced
    if not args.no_extras:
        try:
            extras.update(secondary_configs(torch, dist, world, rank, args.steps))
        except Exception as e:                      # the headline line must survive a failure of a secondary run
            extras['extras_error'] = '%s: %s' % (type(e).__name__, e)
            if world > 1:
                raise
    if rank == 0:
        # the CPU baseline is taken at N = 1 only (torchrun pins OMP_NUM_THREADS=1 for N > 1, which would misstate it)
        cpu_obj = None
        if world == 1:
            cpu_obj = cpu_baseline_full()
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': W,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'cp_als dense 3-way 512x512x512 rank 32 FP64 (BASELINE.json configs[1])'
                                   + ('' if world == 1 else '; weak scaling: %dx512x512 cut along mode 0, one 512^3 slab per GPU' % gshape[0]),
                       'shape': list(gshape), 'rank': R, 'init': 'random', 'fit_method': 'full',
                       'parallelism': 'mode0-shard x%d' % world,
                       'algorithm': 'dimension tree: X is streamed twice per iteration (pass A = X x_2 U_2 serves the '
                                    'MTTKRPs of modes 0 and 1, pass B is the mode-2 MTTKRP); value counts the '
                                    "reference's three MTTKRPs per iteration as the algorithmic bytes",
                       'cache': 'tensor slab 1.07 GB per GPU >> 126 MB L2: every pass streams it from HBM',
                       'preconditioning_steps': PRE},
            'als_iters_per_sec': 1e3 / ms_step, 'final_fit': fit,
            'clocks': clocks, 'e2e': e2e_obj, 'gpu_launches': int(launches),
            'roofline': roofline, 'cpu_baseline': cpu_obj,
        }
        line.update(extras)
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def device_powerlaw_coo(torch, shape, nnz, seed=0, lo_hi=None):
    """BASELINE configs[2] generator (subscript = min(int(I * u^3), I - 1), u uniform) drawn on the device
    with torch's generator -- same distribution as the host generator of the parity tests, no 6.4 GB host round trip."""
    g = torch.Generator(device='cuda')
    g.manual_seed(seed)
    subs = []
    for m, I in enumerate(shape):
        u = torch.rand(nnz, dtype=torch.float64, device='cuda', generator=g)
        if m == 0 and lo_hi is not None:                 # a rank's quantile range of the sharded mode
            u = lo_hi[0] + u * (lo_hi[1] - lo_hi[0])
        subs.append(torch.clamp((I * u ** 3.0).to(torch.int64), max=I - 1))
        del u
    vals = torch.rand(nnz, dtype=torch.float64, device='cuda', generator=g)
    return subs, vals


def time_sparse_mttkrp(torch, K, dev, coo, U, R, n, steps):
    M = dev.empty((coo.shape[n], R))
    for _ in range(3):
        K.mttkrp_coo(coo, U, R, n, out=M)
    ts = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K.mttkrp_coo(coo, U, R, n, out=M)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def sparse_roofline(torch, K, dev, nnz, steps, variants=False):
    """The three MTTKRPs of configs[2] (10M x 1M x 100K power law, rank 16) at `nnz` non-zeros: ms per mode, algorithmic
    GB/s (nnz (4N + 8) + 8 R sum I bytes per mode) against measured HBM, and against the L2->SM gather
    bound (every non-zero pulls two 128-byte factor rows through L2: 99 % of the index pairs are distinct)."""
    R = 16
    hbm_peak, peak_src = measured_peaks()
    subs, vals = device_powerlaw_coo(torch, SPARSE_SHAPE, nnz)
    torch.cuda.synchronize()
    tic = time.perf_counter()
    coo = K.CooHandle.from_device(subs, vals, SPARSE_SHAPE)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - tic
    g = torch.Generator(device='cuda')
    g.manual_seed(1)
    U = [torch.rand(s, R, dtype=torch.float64, device='cuda', generator=g) for s in SPARSE_SHAPE]
    b = alg_bytes_sparse(SPARSE_SHAPE, nnz, R)
    gather_bytes = float(nnz) * 2 * 8 * R + float(nnz) * (4 * 3 + 8)
    traffic = {}
    prof = os.path.join(ROOT, 'profiles', 'sparse_mttkrp_traffic.json')
    if os.path.exists(prof):
        with open(prof) as f:
            traffic = json.load(f)
    out = {'kernel': 'mttkrp_coo_kernel (segmented reduction over the mode-sorted copy, cp.async-staged factor-row gathers)',
           'bound': 'hbm', 'unit': UNIT, 'peak': hbm_peak, 'peak_source': peak_src, 'nnz': nnz, 'rank': R,
           'shape': list(SPARSE_SHAPE), 'alg_bytes_per_launch': b, 'build_seconds_device_arrays': build_s,
           'device_bytes': coo.device_bytes(), 'modes': {},
           'l2_gather_bound': {'bytes_per_launch': gather_bytes,
                               'note': 'bytes that must cross L2->SM per launch (two uncorrelated 128-byte factor rows + 20 '
                                       'bytes of subscripts/value per non-zero); at the measured L2->SM ceiling of ~10-12 TB/s '
                                       'this, not HBM, bounds the launch'}}
    for n in range(3):
        ms, best = time_sparse_mttkrp(torch, K, dev, coo, U, R, n, steps)
        out['modes'][str(n)] = {'ms': ms, 'ms_best': best, 'achieved': b / ms / 1e6, 'frac': b / ms / 1e6 / hbm_peak,
                                'l2_gather_TBps': gather_bytes / ms / 1e9,
                                'traffic': (traffic.get(str(nnz), {}) or {}).get(str(n))}
    if variants:
        os.environ['SKT_COO_CA'] = '1'
        out['variant_ca'] = {str(n): time_sparse_mttkrp(torch, K, dev, coo, U, R, n, steps)[0] for n in range(3)}
        os.environ['SKT_COO_CA'] = '0'
        coo.close()
        os.environ['SKT_COO_PANELS'] = '0'
        coo = K.CooHandle.from_device(subs, vals, SPARSE_SHAPE)
        out['variant_no_panels'] = {str(n): time_sparse_mttkrp(torch, K, dev, coo, U, R, n, steps)[0] for n in range(3)}
        os.environ['SKT_COO_PANELS'] = '1'
        for rows in (1 << 18, 1 << 20):
            coo.close()
            os.environ['SKT_COO_PANEL_ROWS'] = str(rows)
            coo = K.CooHandle.from_device(subs, vals, SPARSE_SHAPE)
            out['variant_panel_rows_%d' % rows] = {str(n): time_sparse_mttkrp(torch, K, dev, coo, U, R, n, steps)[0] for n in range(3)}
        os.environ.pop('SKT_COO_PANEL_ROWS')
    coo.close()
    del subs, vals, U
    torch.cuda.empty_cache()
    fr = [out['modes'][str(n)]['frac'] for n in range(3)]
    out['achieved'] = float(np.mean([out['modes'][str(n)]['achieved'] for n in range(3)]))
    out['frac'] = float(np.mean(fr))
    return out


def timed_sweeps(torch, dist, world, eng, steps, warm):
    """ms per ALS iteration of an engine (device events, max over ranks), after `warm` untimed iterations."""
    for w in range(warm):
        eng.sweep(first_iter=(w == 0))
        eng.begin_read()
        eng.prefetch_next_sweep()
        eng.read_fit()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    fit = 0.0
    for _ in range(steps):
        eng.sweep(first_iter=False)
        eng.begin_read()
        eng.prefetch_next_sweep()
        fit = eng.read_fit()
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if os.environ.get('SKT_SWEEP_TRACE', '0') == '1' and hasattr(eng, 'trace_begin'):
        eng.trace_begin()
        for _ in range(5):
            eng.sweep(first_iter=False)
            eng.begin_read()
            eng.prefetch_next_sweep()
            eng.read_fit()
        torch.cuda.synchronize()
        rep = eng.trace_report()
        if int(os.environ.get('RANK', '0')) == 0:
            sys.stderr.write('[sweep trace, ms per iteration on rank 0] ' +
                             ', '.join('%s %.3f' % (k, v / 5.0) for k, v in rep.items()) + '\n')
    return float(ms.item()) / steps, fit


def device_shard_cfg5(torch, _parallel, dtensor, world, rank):
    """This rank's slab of the dense 160^4 tensor (BASELINE configs[4]), generated on the device."""
    gshape = (160, 160, 160, 160)
    smode = _parallel.shard_mode(gshape, dense=True)
    offsets = _parallel.split_rows(gshape[smode], world)
    lsh = list(gshape)
    lsh[smode] = int(offsets[rank + 1] - offsets[rank])
    g = torch.Generator(device='cuda')
    g.manual_seed(2000 + rank)
    Xd = torch.rand(tuple(lsh), dtype=torch.float64, device='cuda', generator=g)
    Xl = dtensor(np.broadcast_to(np.zeros(1), tuple(lsh)))      # shape carrier only: the data lives in Xl._dev
    Xl._dev = Xd
    return _parallel.LocalShard(Xl, smode, gshape, offsets), gshape, smode


def device_shard_cfg3(torch, K, _parallel, sptensor, world, rank, nnz):
    """This rank's equal-nnz row range of the sparse configs[2] tensor, generated and sorted on the device."""
    gshape = SPARSE_SHAPE
    I0 = gshape[0]
    # the engine's cuts (_parallel.split_nnz) balance nnz + w * rows; with the generator's CDF F(x) = (x / I0)^(1/3) the
    # same cuts follow in closed form: solve nnz F(x) + w x = r / world * (nnz + w I0) for x
    w = float(os.environ.get('SKT_SHARD_ROW_WEIGHT', '2'))
    offsets = [0]
    for r in range(1, world):
        target = r / float(world) * (nnz + w * I0)
        lo_, hi_ = 0.0, float(I0)
        for _ in range(80):
            mid = 0.5 * (lo_ + hi_)
            if nnz * (mid / I0) ** (1.0 / 3.0) + w * mid < target:
                lo_ = mid
            else:
                hi_ = mid
        offsets.append(int(np.ceil(hi_)))
    offsets = np.array(offsets + [I0], dtype=np.int64)
    q_lo, q_hi = (offsets[rank] / float(I0)) ** (1.0 / 3.0), (offsets[rank + 1] / float(I0)) ** (1.0 / 3.0)
    nloc = int(round(nnz * (q_hi - q_lo)))
    subs, vals = device_powerlaw_coo(torch, gshape, nloc, seed=3000 + rank, lo_hi=(q_lo, q_hi))
    subs[0] = torch.clamp(subs[0], min=int(offsets[rank]), max=int(offsets[rank + 1]) - 1) - int(offsets[rank])
    lshape = (int(offsets[rank + 1] - offsets[rank]),) + tuple(gshape[1:])
    coo = K.CooHandle.from_device(subs, vals, lshape)
    del subs, vals
    torch.cuda.empty_cache()
    S = sptensor(tuple(np.zeros(0, dtype=np.int64) for _ in gshape), np.zeros(0), shape=lshape)
    S._coo = coo                                               # the engine uses the sorted device copies as they are
    tot = torch.tensor([nloc], dtype=torch.int64, device='cuda')
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tot)
    return _parallel.LocalShard(S, 0, gshape, offsets), gshape, int(tot.item())


def secondary_configs(torch, dist, world, rank, steps):
    """The other BASELINE.json configurations at this GPU count, as extra keys of the default line:
    strong scaling of configs[4] (dense 160^4 rank 64) and configs[2] (sparse 200 M nnz rank 16), and -- on one GPU --
    the sparse MTTKRP roofline, the whole tucker_hooi call of configs[3] and the launch-latency case configs[0]."""
    from sktensor import _kernels as K, _parallel, cp_als, dtensor, sptensor, tucker_hooi
    from sktensor import _device as dev
    out = {}
    hbm_peak, _ = measured_peaks()
    strong = {}
    # ---- configs[4]
    X, gshape, smode = device_shard_cfg5(torch, _parallel, dtensor, world, rank)
    np.random.seed(1)
    U = [None] + [np.random.rand(gshape[n], 64) for n in range(1, 4)]
    eng = _parallel.CpAlsEngine(X, 64, U)
    ms, fit = timed_sweeps(torch, dist, world, eng, max(3, min(steps, 10)), 3)
    alg = 4 * alg_bytes_dense(gshape, 64)
    strong['cfg5_dense_160^4_rank64'] = {'ms_per_iteration': ms, 'als_iters_per_sec': 1e3 / ms, 'GBps_algorithmic': alg / ms / 1e6,
                                         'sharded_mode': smode, 'final_fit': fit,
                                         'flops_per_iteration_reference': 4 * 2.0 * 64 * 160.0 ** 4}
    del eng
    # the same with the opt-in sliced passes (INT8 tensor cores): reported beside the default, never instead of it
    os.environ['SKT_DENSE_I8'] = '6'
    try:
        np.random.seed(1)
        U = [None] + [np.random.rand(gshape[n], 64) for n in range(1, 4)]
        eng = _parallel.CpAlsEngine(X, 64, U)
        if eng.i8 is not None:
            ms8, fit8 = timed_sweeps(torch, dist, world, eng, max(3, min(steps, 10)), 3)
            strong['cfg5_dense_160^4_rank64']['opt_in_int8_slices'] = {
                'ms_per_iteration': ms8, 'als_iters_per_sec': 1e3 / ms8, 'final_fit': fit8, 'fit_delta_vs_default': abs(fit8 - fit),
                'speedup_vs_default': ms / ms8}
        eng.release()
        del eng
    finally:
        os.environ.pop('SKT_DENSE_I8', None)
    del X
    torch.cuda.empty_cache()
    # ---- configs[2]
    X, gshape, nnz = device_shard_cfg3(torch, K, _parallel, sptensor, world, rank, 200000000)
    np.random.seed(1)
    U = [None] + [np.random.rand(gshape[n], 16) for n in range(1, 3)]
    eng = _parallel.CpAlsEngine(X, 16, U)
    ms, fit = timed_sweeps(torch, dist, world, eng, max(3, min(steps, 10)), 3)
    alg = 3 * alg_bytes_sparse(gshape, nnz, 16)
    strong['cfg3_sparse_200Mnnz_rank16'] = {'ms_per_iteration': ms, 'als_iters_per_sec': 1e3 / ms, 'GBps_algorithmic': alg / ms / 1e6,
                                            'frac_of_hbm_per_gpu': alg / ms / 1e6 / world / hbm_peak, 'final_fit': fit}
    coo = X.tensor._coo
    del eng, X
    coo.close()
    torch.cuda.empty_cache()
    out['strong_scaling'] = strong
    if world > 1 or rank != 0:
        return out
    # ---- one GPU only from here
    out['roofline_sparse'] = sparse_roofline(torch, K, dev, 200000000, 5)
    out['roofline_sparse_20M'] = {k: v for k, v in sparse_roofline(torch, K, dev, 20000000, 5).items()
                                  if k in ('modes', 'achieved', 'frac', 'nnz')}
    # configs[3]: tucker_hooi(384^3, [32, 32, 32]) -- the whole public call, host tensor in, host core/factors out
    Xh = dtensor(np.random.RandomState(0).rand(384, 384, 384))
    hooi = {}
    for init in ('random', 'nvecs'):
        secs = {}
        for it in (5, 20):
            np.random.seed(0)
            tucker_hooi(Xh, [32, 32, 32], init=init, maxIter=2, conv=0)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                np.random.seed(0)
                tic = time.perf_counter()
                tucker_hooi(Xh, [32, 32, 32], init=init, maxIter=it, conv=0)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - tic)
            secs[it] = float(np.median(ts))
        hooi[init] = {'whole_call_its_per_sec_20_iterations': 20 / secs[20], 'whole_call_seconds_20_iterations': secs[20],
                      'ms_per_iteration_steady': (secs[20] - secs[5]) / 15.0 * 1e3}
    # share of the symmetric eigensolver (own cluster kernel) in one iteration: 3 calls of n = 384, k = 32
    Yg = np.random.RandomState(1).rand(384, 1024)
    G = dev.to_device(Yg.dot(Yg.T))
    for _ in range(3):
        K.symeig_top(G, 32)
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K.symeig_top(G, 32)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    eig_ms = float(np.median(ts))
    hooi['eigensolver'] = {'kernel': 'symeig_cluster_kernel (own code, no cuSOLVER)', 'ms_per_call_n384_k32': eig_ms,
                           'share_of_iteration': 3 * eig_ms / hooi['nvecs']['ms_per_iteration_steady']}
    out['cfg4_hooi_384^3_rank32'] = hooi
    del Xh, G
    torch.cuda.empty_cache()
    # configs[0]: cp_als 10x11x8 rank 3 -- pure launch latency; next to the CPU port on the same input
    rng = np.random.RandomState(0)
    X1 = rng.rand(10, 11, 8)
    lat = []
    for _ in range(7):
        np.random.seed(42)
        tic = time.perf_counter()
        P, fit1, itr1, _ = cp_als(dtensor(X1), 3, init='random')
        lat.append(time.perf_counter() - tic)
    from oracle import sktensor_oracle as orc
    cpu = []
    for _ in range(5):
        np.random.seed(42)
        tic = time.perf_counter()
        _, _, fo, io, _ = orc.cp_als(X1, 3, init='random')
        cpu.append(time.perf_counter() - tic)
    out['cfg1_latency'] = {'gpu_whole_call_ms': float(np.median(lat[2:])) * 1e3, 'iterations': int(itr1) + 1, 'fit': float(fit1),
                           'cpu_port_whole_call_ms': float(np.median(cpu)) * 1e3, 'cpu_port_fit': float(np.ravel(fo)[0]),
                           'note': '880-element tensor: launch-latency bound on the GPU; reported, not a target'}
    return out


def run_sparse(args):
    """Secondary bench (not the driver's line): BASELINE.json configs[2] generator, scaled by --nnz."""
    import torch
    from sktensor import _kernels as K
    from sktensor import _device as dev
    out = sparse_roofline(torch, K, dev, args.nnz, args.steps if args.steps != 200 else 10, variants=args.variants)
    out['metric'] = 'sparse_mttkrp_GBps'
    print(json.dumps(out))


def run_config(args):
    """Secondary, multi-GPU-capable runs of the other BASELINE.json configurations (not the driver's line):

      --workload cfg3 : cp_als on the sparse 10M x 1M x 100K power-law tensor, --nnz non-zeros in total (200 M =
                        configs[2]), rank 16, cut along mode 0 into equal-nnz row ranges, one per GPU
      --workload cfg5 : cp_als on the dense 160^4 tensor, rank 64, cut along a middle mode (configs[4]; strong scaling)

    Every rank generates only its own slab (same distribution as SURVEY 8(d): mode-0 subscripts are drawn by
    inverse-CDF sampling inside the rank's quantile range).  Prints one JSON line on rank 0.
    """
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from sktensor import _lib, _parallel, dtensor, sptensor
    lib = _lib.load()
    rng = np.random.default_rng(1000 + rank)
    tic = time.perf_counter()
    if args.workload == 'cfg3':
        R, gshape = 16, SPARSE_SHAPE
        I0 = gshape[0]
        offsets = np.array([int(np.ceil(I0 * (r / float(world)) ** 3.0)) for r in range(world)] + [I0], dtype=np.int64)
        nloc = args.nnz // world
        u = (rank + rng.random(nloc)) / float(world)                          # this rank's quantile range
        i0 = np.clip((I0 * u ** 3.0).astype(np.int64), offsets[rank], offsets[rank + 1] - 1) - offsets[rank]
        subs = (i0,) + tuple(np.minimum((I * rng.random(nloc) ** 3.0).astype(np.int64), I - 1) for I in gshape[1:])
        lshape = (int(offsets[rank + 1] - offsets[rank]),) + tuple(gshape[1:])
        X = _parallel.LocalShard(sptensor(subs, rng.random(nloc), shape=lshape), 0, gshape, offsets)
        alg = 3 * alg_bytes_sparse(gshape, nloc * world, R)
        desc = 'cp_als sparse 10Mx1Mx100K power-law, %d nnz, rank 16 (BASELINE.json configs[2]), mode-0 equal-nnz cuts' % (nloc * world)
    else:
        R, gshape = 64, (160, 160, 160, 160)
        smode = _parallel.shard_mode(gshape, dense=True)          # a middle mode: both passes keep their full GEMM K
        offsets = _parallel.split_rows(gshape[smode], world)
        rows = int(offsets[rank + 1] - offsets[rank])
        lsh = list(gshape)
        lsh[smode] = rows
        X = _parallel.LocalShard(dtensor(rng.random(tuple(lsh))), smode, gshape, offsets)
        alg = 4 * alg_bytes_dense(gshape, R)
        desc = 'cp_als dense 4-way 160^4 rank 64 (BASELINE.json configs[4]), cut along mode %d (strong scaling)' % smode
    gen_s = time.perf_counter() - tic
    np.random.seed(1)
    U = [None] + [np.random.rand(gshape[n], R) for n in range(1, len(gshape))]
    tic = time.perf_counter()
    eng = _parallel.CpAlsEngine(X, R, U)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - tic
    W = max(args.warmup, 3)
    steps = args.steps if args.steps != 200 else 20
    for w in range(W):
        eng.sweep(first_iter=(w == 0))
        eng.begin_read()
        eng.prefetch_next_sweep()
        eng.read_fit()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = lib.skt_launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    fit = 0.0
    for _ in range(steps):
        eng.sweep(first_iter=False)
        eng.begin_read()
        eng.prefetch_next_sweep()
        fit = eng.read_fit()
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / steps
    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        line = {'metric': METRIC, 'value': alg / (ms_step * 1e-3) / 1e9, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': W,
                'ms_per_step': ms_step, 'als_iters_per_sec': 1e3 / ms_step, 'higher_is_better': True,
                'scaling': 'strong', 'dtype': 'f64', 'data': 'synthetic', 'final_fit': fit,
                'config': {'workload': desc, 'shape': list(gshape), 'rank': R},
                'frac_of_hbm_per_gpu': alg / (ms_step * 1e-3) / 1e9 / world / hbm_peak,
                'host_seconds': {'generate_slab': gen_s, 'upload_sort_setup': setup_s},
                'gpu_launches': int(lib.skt_launch_count() - launches0)}
        os.write(real_stdout, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='dense', choices=['dense', 'sparse', 'cfg3', 'cfg5'])
    ap.add_argument('--nnz', type=int, default=20000000)
    ap.add_argument('--no-extras', action='store_true', help='default workload: skip the secondary configurations')
    ap.add_argument('--variants', action='store_true', help='sparse workload: also time the experiment switches')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.workload == 'sparse':
        return run_sparse(args)
    if args.workload in ('cfg3', 'cfg5'):
        return run_config(args)
    return run_gpu(args)


if __name__ == '__main__':
    main()
