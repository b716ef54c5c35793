# -*- coding: utf-8 -*-
"""
bench.py -- rays traced per second, source -> detector (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rays R] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1]): XicsrtSourceDirected (Gaussian line) ->
XicsrtOpticSphericalCrystal (Bragg test, Gaussian rocking curve) -> XicsrtOpticDetector,
1e9 rays per iteration per GPU (weak scaling), history off, images on.  One "step" = one
iteration: the FP32 broad phase (k_cull32) + the fused FP64 kernel (k_trace) over this rank's
ray-id range and, for N > 1, one NCCL all-reduce of the packed counters + images (issued on a
second stream so that it overlaps the next step's kernels).

value     rays launched per second over all ranks, device time (CUDA events on the launch
          stream, max over ranks), scene already uploaded.
e2e       the same through the public API call ``xicsrt_b200.raytrace(config)`` per step:
          host config dict in, host result dict out (scene preparation + upload, launch,
          all-reduce, device->host copy of counters and images inside the timed region).
roofline  the kernels read no global memory beyond the id list: FP64-pipe bound by the
          counting rule of SURVEY.md section 8d (achieved = rays/s x F flop-equivalents per ray of the
          reference's algorithm, survival fractions from this run's own counters; peak =
          dependent-DFMA-chain microbenchmark measured in this process -- MEASURED_PEAKS.json
          has no FP64 entry).  The executed instruction mix of the committed ncu captures sits
          beside it (executed_ncu): the binding resource is issue slots, not the FP64 pipe.
          roofline_history is the HBM-bound history pass (57 B per ray per element).
configs   the other BASELINE.json configs (3 mosaic, 4 mesh, 5 plasma), each with device rate,
          e2e, roofline by the same counting rule and its own CPU baseline sample.
strong    BASELINE.json's config 2 as stated: 1e9 rays per iteration IN TOTAL, sharded over the N GPUs.
history   keep_history=True (the reference's default) through the public API at 1e8 rays.
target    north_star target: config 5 at 1.25e9 rays per GPU (1e10 over 8 GPUs) + history of a 1e6-ray subsample.
shard_parity  (N > 1, untimed) the reduced counters / images of one step equal a single-rank replay of the
          whole id range.
cpu_baseline / --impl reference
          the unmodified reference (xicsrt.raytrace_mp from the offline install baseline/_ref, kind "reference")
          over the host cores, one run per pool task; without that install the oracle port of the
          reference's NumPy path under the same scheme (kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'rays_traced_per_sec_source_to_detector'
UNIT = 'rays/s'
WORKLOAD = 'config2: XicsrtSourceDirected(Gaussian line) -> XicsrtOpticSphericalCrystal(Bragg, gaussian rocking) -> XicsrtOpticDetector'
WORKLOADS = {
    'config2': WORKLOAD,
    'config3': 'config3: XicsrtSourceDirected -> XicsrtOpticSphericalMosaicCrystal(depth 15, spread 0.4 deg, fwhm 200 urad) -> XicsrtOpticDetector',
    'config4': 'config4: XicsrtSourceDirected -> XicsrtOpticMeshToroidalCrystal(41x41 fine, 5x5 coarse, refine + interpolate, check_bragg off) -> XicsrtOpticDetector',
    'config5': 'config5: XicsrtPlasmaCubic(1e5 bundles, Poisson counts, 10 cm cube) -> XicsrtOpticSphericalCrystal -> XicsrtOpticDetector',
}


def spectrometer(n_rays, seed=0, history=False):
    """Geometry G of SURVEY.md section 8d (reference examples/example_01, testing/integrated_test_01)."""
    return {
        'general': {'number_of_iter': 1, 'number_of_runs': 1, 'random_seed': seed, 'print_results': False,
                    'keep_history': history, 'keep_images': True, 'keep_meta': True},
        'sources': {'source': {
            'class_name': 'XicsrtSourceDirected', 'intensity': n_rays, 'wavelength': 3.9492,
            'spread': float(np.radians(10.0)), 'temperature': 1000.0, 'mass_number': 39.948, 'linewidth': 0.0,
            'xsize': 0.0, 'ysize': 0.0, 'zsize': 0.0}},
        'optics': {
            'crystal': {'class_name': 'XicsrtOpticSphericalCrystal', 'check_size': True,
                        'origin': [0.0, 0.0, 0.80374151], 'zaxis': [0.0, 0.59497864, -0.80374151],
                        'xsize': 0.2, 'ysize': 0.2, 'radius': 1.0, 'crystal_spacing': 2.45676,
                        'rocking_type': 'gaussian', 'rocking_fwhm': 48.070e-6},
            'detector': {'class_name': 'XicsrtOpticDetector', 'origin': [0.0, 0.76871290, 0.56904832],
                         'zaxis': [0.0, -0.95641806, 0.29200084], 'xsize': 0.4, 'ysize': 0.2}},
    }


def workload_config(name, n_rays, seed=0, history=False, bundle_count=100000):
    """The BASELINE.json configs: config2 is the headline; the others are measured with --workload."""
    cfg = spectrometer(n_rays, seed=seed, history=history)
    crystal = cfg['optics']['crystal']
    if name == 'config2':
        return cfg
    if name == 'config3':          # mosaic HOPG crystal, random mosaic normals, per-ray reflectivity mask
        crystal.update({'class_name': 'XicsrtOpticSphericalMosaicCrystal', 'mosaic_spread': float(np.radians(0.4)),
                        'mosaic_depth': 15, 'rocking_fwhm': 200e-6})
        return cfg
    if name == 'config4':          # mesh-defined toroidal crystal, coarse/fine refinement + interpolation
        crystal.pop('radius')
        crystal.update({'class_name': 'XicsrtOpticMeshToroidalCrystal', 'radius_major': 1.0, 'radius_minor': 0.2,
                        'mesh_size': (41, 41), 'mesh_coarse_size': (5, 5), 'check_bragg': False})
        return cfg
    if name == 'config5':          # extended plasma source -> crystal -> detector
        # E[rays] = emissivity * dt * bundle_volume * Omega/4pi * volume / (bundle_count * bundle_volume)
        spread = float(np.radians(2.0))
        omega = np.sin(spread / 2)**2
        volume = 0.1**3
        cfg['sources']['source'] = {
            'class_name': 'XicsrtPlasmaCubic', 'origin': [0.0, 0.0, 0.0], 'xsize': 0.1, 'ysize': 0.1, 'zsize': 0.1,
            'target': [0.0, 0.0, 0.80374151], 'spread': spread, 'bundle_type': 'voxel', 'bundle_volume': 1e-9,
            'bundle_count': bundle_count, 'use_poisson': True, 'time_resolution': 1.0,
            'emissivity': float(n_rays) / (omega * volume), 'temperature': 1000.0, 'mass_number': 39.948,
            'wavelength': 3.9492, 'linewidth': 0.0, 'max_rays': int(4 * n_rays) + 1000}
        return cfg
    raise KeyError(name)


# ---------------------------------------------------------------------------
# clocks

class ClockSampler:
    """
    nvidia-smi sampling in the background (one sample per 20 ms with its own timestamp).  It is started well before the
    timed region (the tool needs 0.1 - 0.3 s to come up, longer than a short timed region lasts); only the samples whose
    timestamps fall inside the region that mark_begin / mark_end bracket are reported.
    """

    QUERY = ('timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.file = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.proc = None
        self.t_begin = self.t_end = None

    def start(self):
        if self.proc is not None:
            return
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits', '-lms', '20',
                 '-i', str(self.gpu_index)], stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    @staticmethod
    def _stamp(text):
        import datetime
        try:
            return datetime.datetime.strptime(text, '%Y/%m/%d %H:%M:%S.%f').timestamp()
        except ValueError:
            return None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        time.sleep(0.03)            # the sample that covers the end of the region
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        rows = []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.file.read().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 10:
                continue
            try:
                row = (self._stamp(f[0]), float(f[2]), float(f[3]), float(f[4]),
                       {name for name, val in zip(names, f[6:10]) if val.lower().startswith('active')})
            except ValueError:
                continue
            rows.append(row)
        self.file.close()
        os.unlink(self.file.name)
        if not rows:
            return out
        inside = [r for r in rows if r[0] is not None and self.t_begin is not None and self.t_end is not None
                  and self.t_begin - 0.005 <= r[0] <= self.t_end + 0.015]
        where = 'inside the timed region'
        if not inside:
            # a region shorter than the sampling period: the samples under load next to it (warm-up, the steps themselves)
            pmax = max(r[3] for r in rows)
            inside = [r for r in rows if r[3] > 0.5 * pmax] or rows
            where = 'under load next to the timed region (none fell inside it)'
        reasons = set()
        for r in inside:
            reasons |= r[4]
        out.update({'sm_mhz': float(np.median([r[1] for r in inside])), 'sm_max_mhz': float(max(r[2] for r in inside)),
                    'power_w_max': float(max(r[3] for r in inside)), 'reasons': sorted(reasons), 'samples': len(inside),
                    'samples_total': len(rows), 'sampled': where})
        return out


# ---------------------------------------------------------------------------
# flop-equivalents per launched ray of the REFERENCE's algorithm (SURVEY.md section 8d counting rule: add / mul /
# compare = 1, FMA = 2, transcendental = 20; stage constants of that section, survival fractions from the run)

def flops_config2(f_bounds, f_reflect):
    """Directed source 133 + sphere distance/location 25 + normal 12 + bounds 13, then Bragg 78 on
    the rays inside the crystal bounds and reflect + detector plane + bounds + bin (12+20+13+6) on the
    reflected ones."""
    return 133.0 + 25.0 + 12.0 + 13.0 + f_bounds * 78.0 + f_reflect * (12.0 + 20.0 + 13.0 + 6.0)


def flops_config3(f_bounds, f_reflect, depth=15):
    """As config 2 up to the bounds test; then per ray inside the bounds the crystallite layers it visits at 230
    flop-equivalents each (2 normals + unit vector, basis + rotation, Bragg test, reflect).  With a per-layer
    reflection probability p, f_reflect / f_bounds = 1 - (1 - p)^depth and the mean number of layers visited is
    (1 - (1 - p)^depth) / p."""
    q = min(max(f_reflect / max(f_bounds, 1e-300), 0.0), 1.0 - 1e-12)
    p = 1.0 - (1.0 - q) ** (1.0 / depth)
    layers = depth if p <= 0.0 else q / p
    return 133.0 + 25.0 + 12.0 + 13.0 + f_bounds * layers * 230.0 + f_reflect * (20.0 + 13.0 + 6.0), layers


def flops_config4(f_coarse, f_crystal, n_coarse_faces=32):
    """Directed source 133; Moeller-Trumbore against every coarse face (50 each); for the rays that hit the coarse
    mesh: the reference's 8 candidate faces (50 each), barycentric lookup 20 and four Clough-Tocher cubics (120 each),
    bounds 13; mirror-like reflection 12 (check_bragg off) and detector plane + bounds + bin (20+13+6) for the rays
    that leave the crystal."""
    return 133.0 + 50.0 * n_coarse_faces + f_coarse * (8 * 50.0 + 20.0 + 4 * 120.0 + 13.0) + f_crystal * (12.0 + 20.0 + 13.0 + 6.0)


def flops_config5(f_bounds, f_reflect):
    """config 2 with the focused voxel source of a plasma bundle (+60: per-ray cone axis and basis)."""
    return flops_config2(f_bounds, f_reflect) + 60.0


# ---------------------------------------------------------------------------
# CPU arm on the host cores: the UNMODIFIED reference (xicsrt.raytrace_mp) when its offline install travelled with the
# repository (baseline/_ref, git-ignored: `python __graft_entry__.py` installs it where /root/reference exists),
# else the oracle port of the reference's NumPy path (same scheme: one run per pool task)

REF_DIR = os.path.join(ROOT, 'baseline', '_ref')
_REF = {}


def reference_module():
    """The reference package from baseline/_ref, or None (then the oracle port is timed)."""
    if 'mod' not in _REF:
        mod = None
        if os.environ.get('XRT_BENCH_CPU_PORT') != '1' and os.path.isfile(os.path.join(REF_DIR, 'xicsrt', '__init__.py')):
            sys.path.insert(0, REF_DIR)
            try:
                import logging
                import xicsrt as mod
                if os.path.realpath(os.path.dirname(mod.__file__)) != os.path.realpath(os.path.join(REF_DIR, 'xicsrt')):
                    mod = None
                else:
                    logging.getLogger('xicsrt').setLevel(logging.WARNING)
            except Exception as e:      # noqa: BLE001  (missing dependency of the reference on this box)
                print(f'[bench] reference install in {REF_DIR} not importable ({e!r}): timing the oracle port', file=sys.stderr)
                mod = None
            finally:
                sys.path.remove(REF_DIR)
        _REF['mod'] = mod
    return _REF['mod']


def cpu_kind():
    return 'reference' if reference_module() is not None else 'port'


def cpu_what():
    return ('the unmodified reference from baseline/_ref: xicsrt.raytrace_mp' if reference_module() is not None
            else 'oracle port of the NumPy path; reference scheme xicsrt_multiprocessing')


class _StdoutToStderr:
    """The reference and its pool workers log to fd 1; the bench's stdout is one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def _warm_parent(ref, workload):
    """The reference imports its object classes (and scipy) lazily inside the first raytrace of a process, ~1.7 s; its
    pool forks a fresh set of workers per raytrace_mp call.  One tiny in-process run per workload lets the workers
    inherit the imported modules, so the timed sample holds raytracing only."""
    if ref is None or workload in _REF.setdefault('warm', set()):
        return
    _REF['warm'].add(workload)
    cfg = workload_config(workload, 2000, seed=1, bundle_count=20) if workload == 'config5' \
        else workload_config(workload, 2000, seed=1, history=False)
    cfg['general']['print_results'] = False
    with _StdoutToStderr():
        ref.raytrace(cfg)


def cpu_reference_step(rays_per_run, runs, processes, seed, workload='config2'):
    ref = reference_module()
    _warm_parent(ref, workload)
    if workload == 'config5':
        # the reference builds one Python source object per bundle (0.5 ms each): 2000 bundles keep the sample bounded
        cfg = workload_config('config5', rays_per_run, seed=seed, bundle_count=2000)
    else:
        cfg = workload_config(workload, rays_per_run, seed=seed, history=False)
    cfg['general']['number_of_runs'] = runs
    cfg['general']['print_results'] = False
    if ref is not None:
        with _StdoutToStderr():
            t0 = time.perf_counter()
            res = ref.raytrace_mp(cfg, processes=processes)
            dt = time.perf_counter() - t0
    else:
        import oracle
        t0 = time.perf_counter()
        res = oracle.raytrace_mp(cfg, processes=processes)
        dt = time.perf_counter() - t0
    n = int(res['total']['meta']['source']['num_out'])
    return n, dt


def cpu_baseline_for(workload, procs):
    """A bounded CPU sample of one workload (a few seconds of host time)."""
    per_run = {'config2': 1_000_000, 'config3': 200_000, 'config4': 100_000, 'config5': 500_000}[workload]
    runs = {'config2': 8 * procs, 'config3': 2 * procs, 'config4': 2 * procs, 'config5': 2 * procs}[workload]
    if workload == 'config2':
        cpu_reference_step(200000, procs, procs, seed=7)          # warm the pool / imports
    n, dt = cpu_reference_step(per_run, runs, procs, seed=8, workload=workload)
    note = ' (2000 bundles instead of 1e5: the reference spends 0.5 ms of Python per bundle)' if workload == 'config5' else ''
    return {'value': n / dt, 'unit': UNIT, 'cores': procs, 'kind': cpu_kind(),
            'sample': f'{runs} runs x {per_run} rays over {procs} processes, {dt:.1f} s{note} ({cpu_what()})'}


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = min(cores, 64)
    rays_per_run = 500000
    runs = procs
    for w in range(args.warmup):
        cpu_reference_step(rays_per_run, runs, procs, seed=100 + w)
    total_rays, total_t = 0, 0.0
    for k in range(args.steps):
        n, dt = cpu_reference_step(rays_per_run, runs, procs, seed=k)
        total_rays += n
        total_t += dt
    value = total_rays / total_t
    sample = f'{runs} runs x {rays_per_run} rays per step over {procs} processes ({cpu_what()})'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_t / max(args.steps, 1),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'rays_per_step': runs * rays_per_run, 'history': False},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': procs, 'kind': cpu_kind(), 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm

class Ctx:
    """Process-wide state of the GPU arm."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise RuntimeError('bench.py needs a CUDA device: xicsrt_b200 has no CPU path')
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device('cuda', self.local_rank)
        if self.world > 1:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl', device_id=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)   # > 126 MB L2
        self.side = torch.cuda.Stream(device=self.dev)      # all-reduce stream

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]


def make_tracer(ctx, workload, total_rays, seed=0):
    from xicsrt_b200 import config as xconfig
    from xicsrt_b200 import _driver as xrt
    cfg = workload_config(workload, total_rays, seed=seed, history=False)
    return xrt.Tracer(xconfig.get_config(xconfig.to_numpy(cfg)), seed=seed, rank=ctx.rank, world=ctx.world)


def timed_steps(ctx, tracer, steps, warmup, sampler=None):
    """
    W warm-up + K timed steps.  A step = (plasma: new bundle table) + broad phase + fused kernel into one of two
    packed buffers on the launch stream, and for N > 1 the all-reduce of that buffer on a second stream, so that it
    overlaps the kernels of the next step.  Returns (seconds for K steps incl. the last all-reduce, seconds of kernel
    time only, rays launched per step): device times from CUDA events, max over ranks.
    """
    torch = ctx.torch
    bufs = [tracer.packed, torch.zeros_like(tracer.packed)]
    main = torch.cuda.current_stream(ctx.dev)
    reduced = [None, None]

    def step(it, k, ev_kernel=None):
        buf = bufs[k % 2]
        if reduced[k % 2] is not None:
            main.wait_event(reduced[k % 2])          # the all-reduce that last used this buffer
        tracer.begin_iteration(it)
        tracer.trace(it, keep_images=True, packed=buf)
        if ev_kernel is not None:
            ev_kernel.record(main)
        if ctx.world > 1:
            traced = torch.cuda.Event()
            traced.record(main)
            with torch.cuda.stream(ctx.side):
                ctx.side.wait_event(traced)
                tracer.allreduce(packed=buf)
                done = torch.cuda.Event(enable_timing=True)
                done.record(ctx.side)
            reduced[k % 2] = done
            return done
        return None

    for w in range(warmup):
        step(1000 + w, w)
    ctx.barrier()
    if sampler is not None:
        sampler.start()
    start = torch.cuda.Event(enable_timing=True)
    stop = torch.cuda.Event(enable_timing=True)
    k0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    k1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ctx.barrier()
    t_wall0 = time.perf_counter()
    if sampler is not None:
        sampler.mark_begin()
    start.record(main)
    launched = 0
    for k in range(steps):
        ctx.flush.zero_()                   # L2 flush between timed iterations
        k0[k].record(main)
        last = step(k, k, ev_kernel=k1[k])
        launched += tracer.n_rays
    if last is not None:
        main.wait_event(last)               # the step is complete when its reduced counters are
    stop.record(main)
    ctx.barrier()
    t_wall = time.perf_counter() - t_wall0
    if sampler is not None:
        sampler.mark_end()
    # the flushes are outside what a step is: subtract their device time (measured separately, same stream)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(main)
    for _ in range(steps):
        ctx.flush.zero_()
    f1.record(main)
    torch.cuda.synchronize()
    t_flush = f0.elapsed_time(f1) * 1e-3
    t_all = start.elapsed_time(stop) * 1e-3 - t_flush
    t_kern = sum(k0[k].elapsed_time(k1[k]) for k in range(steps)) * 1e-3
    t_all, t_kern = ctx.max_over_ranks([t_all, t_kern])
    tracer.packed = bufs[(steps - 1) % 2]   # counters / images of the last step
    return t_all, t_kern, launched // steps, t_wall


def e2e_through_api(ctx, workload, total_rays, steps, warmup, history=False):
    """The same metric through xicsrt_b200.raytrace(config): host dict in, host dict out, per step."""
    import xicsrt_b200
    res = None
    for w in range(warmup):      # the same `res = raytrace(cfg)` loop a user writes: the previous result is still alive
        res = xicsrt_b200.raytrace(workload_config(workload, total_rays, seed=50 + w, history=history))
    ctx.barrier()
    t0 = time.perf_counter()
    rays, found = 0, 0
    for k in range(steps):
        res = xicsrt_b200.raytrace(workload_config(workload, total_rays, seed=k, history=history))
        rays += int(res['total']['meta']['source']['num_out'])
        if history and ctx.rank == 0:
            found += len(res['found']['history']['detector']['mask'])
    ctx.barrier()
    (t,) = ctx.max_over_ranks([time.perf_counter() - t0])
    return rays / t, t / steps, found // max(steps, 1)


def fp64_peak(ctx):
    import ctypes as C
    from xicsrt_b200 import _lib as L
    torch = ctx.torch
    lib = L.load()
    sink = torch.zeros(1, dtype=torch.float64, device=ctx.dev)
    flops = C.c_double()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream_ptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(lib.xrt_fp64_burn(20000, sink.data_ptr(), C.byref(flops), stream_ptr))
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(3):
        e0.record()
        L.check(lib.xrt_fp64_burn(200000, sink.data_ptr(), C.byref(flops), stream_ptr))
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def bounds_fraction(workload, seed=1):
    """Fraction of the launched rays inside the crystal bounds: the same scene with the Bragg test off (1e7 rays, untimed)."""
    from xicsrt_b200 import config as xconfig
    from xicsrt_b200 import _driver as xrt
    cfg = workload_config(workload, 10_000_000, seed=seed)
    cfg['optics']['crystal']['check_bragg'] = False
    tb = xrt.Tracer(xconfig.get_config(xconfig.to_numpy(cfg)), seed=seed)
    tb.begin_iteration(0)
    tb.trace(0)
    mb, _ = tb.counts_and_images(False)
    tb.close()
    return mb['crystal'] / mb['source']


def plan_kernels(info):
    """The kernels of one step under the launch plan xrt_scene_create chose for the scene."""
    if info.get('broad_phase'):
        return 'k_cull32 + k_trace'
    if info.get('mosaic_broad_phase'):
        return 'k_mosaic32 + k_trace'
    if info.get('mesh_sort'):
        return 'k_mesh_coarse + k_mesh_scan + k_mesh_scatter + k_mesh_refine'
    return 'k_trace'


def roofline_for(workload, meta, rays_per_s_kernel, peak, tracer):
    f_reflect = meta['crystal'] / meta['source']
    extra = {}
    if workload == 'config2':
        f_b = bounds_fraction(workload)
        F = flops_config2(f_b, f_reflect)
        extra = {'f_bounds': f_b, 'f_reflect': f_reflect}
    elif workload == 'config3':
        f_b = bounds_fraction(workload)
        F, layers = flops_config3(f_b, f_reflect)
        extra = {'f_bounds': f_b, 'f_reflect': f_reflect, 'mean_layers_visited': layers}
    elif workload == 'config4':
        # check_bragg is off: every ray that hits the mesh inside the bounds leaves the crystal; the coarse-mesh hit
        # fraction is that of the bounds-free scene, taken from the crystal counter (coarse hit ~ crystal footprint)
        f_c = f_reflect
        F = flops_config4(f_c, f_reflect)
        extra = {'f_coarse_hit': f_c, 'f_crystal': f_reflect}
    else:
        f_b = bounds_fraction(workload)
        F = flops_config5(f_b, f_reflect)
        extra = {'f_bounds': f_b, 'f_reflect': f_reflect}
    achieved = rays_per_s_kernel * F / 1e12
    out = {'bound': 'fp64', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
           'traffic': None, 'flop_equiv_per_ray': F,
           'counting': 'algorithmic flop-equivalents of the reference algorithm (SURVEY 8d), not pipe utilisation',
           'peak_source': 'DFMA-chain microbenchmark (xrt_fp64_burn) measured in this run'}
    out.update(extra)
    try:        # DRAM bytes of one step from the committed ncu captures of the plan's kernels (profiles/r02_traffic.json)
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))['configs'][workload]
        out['traffic'] = prof['dram_bytes_per_launch']
        out['traffic_rays_per_launch'] = prof['rays_per_launch']
    except (OSError, KeyError, ValueError):
        pass
    return out


def table_bytes(tracer):
    """Host -> device bytes of one scene upload: the descriptor plus every table xrt_scene_create copies."""
    import ctypes as C
    from xicsrt_b200 import _lib as L
    return C.sizeof(L.XrtSceneDesc) + int(getattr(tracer, 'upload_bytes', 0))


def run_config(ctx, workload, rays_per_gpu, steps, warmup, peak, with_cpu):
    """One of the non-headline configs: device rate, e2e, roofline, CPU sample."""
    total = rays_per_gpu * ctx.world
    tracer = make_tracer(ctx, workload, total)
    info = tracer.scene.launch_info()
    t_all, t_kern, launched, _ = timed_steps(ctx, tracer, steps, warmup)
    meta, _ = tracer.counts_and_images(True)
    h2d, d2h = table_bytes(tracer), 8 * int(tracer.packed.numel())
    line = {'workload': WORKLOADS[workload], 'value': launched * steps / t_all, 'unit': UNIT,
            'rays_per_step': launched, 'steps': steps, 'ms_per_step': 1e3 * t_all / steps,
            'detected_per_step': meta['detector'], 'launch': info}
    if ctx.rank == 0:
        line['roofline'] = roofline_for(workload, meta, (launched / ctx.world) * steps / t_kern, peak, tracer)
        line['roofline']['kernel'] = plan_kernels(info)
        line['roofline']['kernel_ms_per_step'] = 1e3 * t_kern / steps
    tracer.close()
    e2e, _, _ = e2e_through_api(ctx, workload, total, max(2, steps // 2), 1)
    line['e2e'] = {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                   'api': 'xicsrt_b200.raytrace(config)'}
    line['cpu_baseline'] = None
    if with_cpu and ctx.rank == 0 and ctx.world == 1:
        line['cpu_baseline'] = cpu_baseline_for(workload, min(os.cpu_count() or 1, 64))
    return line


def shard_parity(ctx, workload='config2', n=20_000_000):
    """Untimed: reduced counters / images of one sharded step == a single-rank trace of the whole id range."""
    from xicsrt_b200 import config as xconfig
    from xicsrt_b200 import _driver as xrt
    torch = ctx.torch
    cfg = xconfig.get_config(xconfig.to_numpy(workload_config(workload, n, seed=5)))
    tr = xrt.Tracer(cfg, seed=5, rank=ctx.rank, world=ctx.world)
    tr.begin_iteration(0)
    tr.trace(3)
    tr.allreduce()
    sharded = tr.packed.clone()
    tr.trace(3, ray_begin=0, ray_count=tr.n_rays)        # every rank replays the whole range on its own
    same = bool(torch.equal(tr.packed, sharded))
    tr.close()
    flag = torch.tensor([1 if same else 0], dtype=torch.int64, device=ctx.dev)
    if ctx.world > 1:
        ctx.dist.all_reduce(flag, op=ctx.dist.ReduceOp.MIN)
    return bool(int(flag.cpu()[0]))


def history_replay_roofline(ctx, tracer):
    torch = ctx.torch
    n_h = 1 << 24
    ids = torch.arange(n_h, dtype=torch.int64, device=ctx.dev)
    bufs = tracer.history(0, ids)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        tracer.history(0, ids, out=bufs)
    e1.record()
    torch.cuda.synchronize()
    t_h = e0.elapsed_time(e1) * 1e-3 / reps
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        hbm_peak, src = float(peaks['hbm_gbs']), 'measured'
    except (OSError, KeyError, ValueError):
        hbm_peak, src = 6650.0, 'fallback'
    bytes_h = 57.0 * n_h * tracer.n_elem
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))['k_record']['dram_bytes_per_launch']
    except (OSError, KeyError, ValueError):
        pass
    return {'bound': 'hbm', 'achieved': bytes_h / t_h / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
            'frac': bytes_h / t_h / 1e9 / hbm_peak, 'traffic': traffic, 'peak_source': src,
            'rays': n_h, 'elements': tracer.n_elem, 'ms': t_h * 1e3, 'algorithmic_bytes': bytes_h,
            'kernel': 'k_record<0, PHILOX>: full replay of every ray + SoA stores'}


def target_config5(ctx, rays_per_gpu=1_250_000_000, n_hist=1_000_000, steps=2):
    """north_star target: config 5 at 1e10 rays over 8 GPUs (1.25e9 per GPU here) + history of a 1e6-ray subsample."""
    torch = ctx.torch
    tracer = make_tracer(ctx, 'config5', rays_per_gpu * ctx.world)
    t_all, t_kern, launched, _ = timed_steps(ctx, tracer, steps, 1)
    meta, _ = tracer.counts_and_images(True)
    out = {'workload': WORKLOADS['config5'], 'rays_per_step': launched, 'value': launched * steps / t_all, 'unit': UNIT,
           'ms_per_step': 1e3 * t_all / steps, 'detected_per_step': meta['detector']}
    if ctx.rank == 0:
        ids = torch.arange(n_hist, dtype=torch.int64, device=ctx.dev) * (tracer.n_rays // n_hist)
        bufs = tracer.history(steps - 1, ids)
        torch.cuda.synchronize()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        rays, mask = tracer.history(steps - 1, ids, out=bufs)
        h1.record()
        torch.cuda.synchronize()
        out['history'] = {'rays': n_hist, 'elements': tracer.n_elem, 'found': int(mask[-1].sum()),
                          'ms': h0.elapsed_time(h1), 'GBps': 57.0 * n_hist * tracer.n_elem / (h0.elapsed_time(h1) * 1e-3) / 1e9}
    tracer.close()
    return out


def guarded(ctx, what, fn):
    """A side measurement must never cost the headline line: report the failure instead."""
    try:
        return fn()
    except Exception as exc:      # noqa: BLE001 -- reported in the JSON line
        import traceback
        sys.stderr.write(f'[bench] {what} failed:\n{traceback.format_exc()}\n')
        ctx.torch.cuda.synchronize()
        return {'error': f'{type(exc).__name__}: {exc}'}


def run_gpu_arm(args):
    ctx = Ctx()
    torch = ctx.torch
    rank, world = ctx.rank, ctx.world
    single = args.workload != 'all'
    headline_wl = 'config2' if not single or args.workload == 'config2' else args.workload

    rays_per_gpu = int(args.rays)
    total_rays = rays_per_gpu * world if args.scaling == 'weak' else rays_per_gpu
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.start()             # comes up while the scene is prepared and the FP64 peak is measured
    tracer = make_tracer(ctx, headline_wl, total_rays)
    info = tracer.scene.launch_info()
    peak = fp64_peak(ctx)

    # ---- headline: device-timed steps
    t_steps, t_kernel, launched, t_wall = timed_steps(ctx, tracer, args.steps, args.warmup, sampler)
    clocks = sampler.stop() if rank == 0 else None
    value = launched * args.steps / t_steps
    meta, _ = tracer.counts_and_images(True)
    h2d, d2h = table_bytes(tracer), 8 * int(tracer.packed.numel())
    roof = None
    if rank == 0:
        roof = roofline_for(headline_wl, meta, (launched / world) * args.steps / t_kernel, peak, tracer)
        roof['kernel'] = plan_kernels(info)
        roof['kernel_ms_per_step'] = 1e3 * t_kernel / args.steps
        try:
            prof = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))
            roof['traffic'] = prof['step']['dram_bytes_per_launch']
            roof['executed_ncu'] = prof['step'].get('executed')
        except (OSError, KeyError, ValueError):
            pass
    hist_line = None
    if rank == 0 and headline_wl == 'config2':
        hist_line = guarded(ctx, 'history replay', lambda: history_replay_roofline(ctx, tracer))
    tracer.close()

    # ---- end to end through the public API (host dict in, host dict out)
    e2e_value, _, _ = e2e_through_api(ctx, headline_wl, total_rays, args.steps, min(args.warmup, 2))

    extras = {}
    if not single and not args.quick:
        # ---- the other BASELINE.json configs
        extras['configs'] = {}
        for wl, rays, steps in (('config3', 100_000_000, 5), ('config4', 100_000_000, 3), ('config5', 1_000_000_000, 5)):
            extras['configs'][wl] = guarded(ctx, wl, lambda wl=wl, rays=rays, steps=steps: run_config(
                ctx, wl, rays, steps, 3, peak, with_cpu=not args.no_cpu))
        # ---- strong scaling of config 2: 1e9 rays in total over the N GPUs
        def strong():
            tr = make_tracer(ctx, 'config2', 1_000_000_000)
            t_all, t_kern, n, _ = timed_steps(ctx, tr, args.steps, args.warmup)
            tr.close()
            v = n * args.steps / t_all
            return {'rays_total': n, 'value': v, 'unit': UNIT, 'ms_per_step': 1e3 * t_all / args.steps,
                    'kernel_ms_per_step': 1e3 * t_kern / args.steps,
                    'efficiency_vs_weak': v / value if args.scaling == 'weak' else None,
                    'note': 'ideal strong scaling = N x the single-GPU rate on 1e9 rays = the weak-scaling value of this line'}
        extras['strong'] = guarded(ctx, 'strong scaling', strong)
        # ---- keep_history=True (the reference's default) through the public API, 1e8 rays per GPU
        def hist_e2e():
            n_h = 100_000_000 * world
            # two warm-up calls: the pinned host buffers of the result arrays come from torch's caching host allocator
            v, t, found = e2e_through_api(ctx, 'config2', n_h, 3, 2, history=True)
            return {'rays_per_step': n_h, 'e2e_value': v, 'unit': UNIT, 'ms_per_step': 1e3 * t, 'found_rays_per_step': found,
                    'api': 'xicsrt_b200.raytrace(config) with keep_history=True, history_max_lost=10000',
                    'd2h_bytes_per_step': 57 * 3 * (found + 10000)}
        extras['history'] = guarded(ctx, 'history e2e', hist_e2e)
        extras['target'] = guarded(ctx, 'config5 target', lambda: target_config5(ctx))
    extras['shard_parity'] = guarded(ctx, 'shard parity', lambda: shard_parity(ctx)) if world > 1 else None
    if world > 1 and not single and not args.quick:
        # the same check for the other launch plans (mosaic broad phase, sorted mesh path, plasma bundles)
        extras['shard_parity_configs'] = {wl: guarded(ctx, 'shard parity ' + wl, lambda wl=wl: shard_parity(ctx, wl, 24_000_000))
                                          for wl in ('config3', 'config4', 'config5')}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = guarded(ctx, 'cpu baseline', lambda: cpu_baseline_for(headline_wl, min(os.cpu_count() or 1, 64)))

    if rank == 0:
        n_kernels = 2 if info.get('broad_phase') else 1
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * t_steps / args.steps, 'higher_is_better': True,
            'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOADS[headline_wl], 'rays_per_step': launched,
                       'rays_per_gpu_per_step': launched // world, 'history': False, 'images': True,
                       'parallelism': f'ray-id ranges over {world} GPU(s)',
                       'l2': 'flushed between timed steps (256 MiB memset; its device time is measured and subtracted); '
                             'the only global input of the kernels is the id list written by the same step',
                       'launch': info},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'api': 'xicsrt_b200.raytrace(config)'},
            'gpu_launches': n_kernels * args.steps,
            'roofline': roof,
            'roofline_history': hist_line,
            'cpu_baseline': cpu,
            'detected_per_step': meta['detector'],
            'wall_s_timed_region': t_wall,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--rays', type=float, default=1e9, help='rays per GPU per step')
    ap.add_argument('--scaling', choices=['weak', 'strong'], default='weak')
    ap.add_argument('--impl', choices=['b200', 'reference'], default='b200')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline legs')
    ap.add_argument('--quick', action='store_true', help='headline only: skip configs 3-5, strong scaling, history, target')
    ap.add_argument('--workload', choices=['all', 'config2', 'config3', 'config4', 'config5'], default='all',
                    help='all (default): headline config2 + the other configs as sub-objects; or one config alone')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
