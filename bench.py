# -*- coding: utf-8 -*-
"""
bench.py -- rays traced per second, source -> detector (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rays R] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): XicsrtSourceDirected (Gaussian line) ->
XicsrtOpticSphericalCrystal (Bragg test, Gaussian rocking curve) -> XicsrtOpticDetector,
1e9 rays per iteration per GPU, history off, images on.  One "step" = one iteration:
one fused generate -> trace -> bin kernel launch over this rank's ray-id range and, for
N > 1, one NCCL all-reduce of the packed counters + images.

value     rays launched per second over all ranks, device time (CUDA events on the launch
          stream, max over ranks), scene already uploaded.
e2e       the same through the public API call ``xicsrt_b200.raytrace(config)`` per step:
          host config dict in, host result dict out (scene preparation + upload, launch,
          all-reduce, device->host copy of counters and images inside the timed region).
roofline  dominant kernel k_trace is FP64-pipe bound (it reads no global memory);
          achieved = rays/s x F flop-equivalents per ray (SURVEY.md section 8d counting rule,
          survival fractions from this run's own counters), peak = dependent-DFMA-chain
          microbenchmark measured in this process (MEASURED_PEAKS.json has no FP64 entry).
          roofline_history is the HBM-bound history pass (57 B per ray per element).
cpu_baseline / --impl reference
          the oracle port of the reference's NumPy path, run with multiprocessing over the
          host cores (the reference's xicsrt_multiprocessing scheme: one run per task).
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


def workload_config(name, n_rays, seed=0, history=False):
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
        volume, bundle_count = 0.1**3, 100000
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
    """nvidia-smi sampling in the background while the timed region runs."""

    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.file = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits', '-lms', '20',
                 '-i', str(self.gpu_index)], stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, smax, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.file.read().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        self.file.close()
        os.unlink(self.file.name)
        if sm:
            busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
            out.update({'sm_mhz': float(np.median(busy)), 'sm_max_mhz': float(max(smax)),
                        'power_w_max': float(max(power)), 'reasons': sorted(reasons), 'samples': len(sm)})
        return out


# ---------------------------------------------------------------------------
# flop-equivalents per ray

def flops_per_ray(f_bounds, f_reflect):
    """Directed source 133 + sphere distance/location 25 + normal 12 + bounds 13, then Bragg 78 on
    the rays inside the crystal bounds and reflect + detector plane + bounds + bin (12+20+13+6) on the
    reflected ones."""
    return 133.0 + 25.0 + 12.0 + 13.0 + f_bounds * 78.0 + f_reflect * (12.0 + 20.0 + 13.0 + 6.0)


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference NumPy path on the host cores

def cpu_reference_step(rays_per_run, runs, processes, seed):
    import oracle
    cfg = spectrometer(rays_per_run, seed=seed, history=False)
    cfg['general']['number_of_runs'] = runs
    t0 = time.perf_counter()
    res = oracle.raytrace_mp(cfg, processes=processes)
    dt = time.perf_counter() - t0
    n = int(res['total']['meta']['source']['num_out'])
    return n, dt


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
    sample = f'{runs} runs x {rays_per_run} rays per step over {procs} processes (oracle port of the NumPy path)'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_t / max(args.steps, 1),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'rays_per_step': runs * rays_per_run, 'history': False},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': procs, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm

def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: xicsrt_b200 has no CPU path')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL writes its version / debug lines to stdout by default: keep stdout for the one JSON line
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=dev)

    import ctypes as C
    import xicsrt_b200
    from xicsrt_b200 import _lib as L
    from xicsrt_b200 import config as xconfig
    from xicsrt_b200 import _driver as xrt

    rays_per_gpu = int(args.rays)
    total_rays = rays_per_gpu * world if args.scaling == 'weak' else rays_per_gpu
    cfg = workload_config(args.workload, total_rays, seed=0, history=False)
    full = xconfig.get_config(xconfig.to_numpy(cfg))
    tracer = xrt.Tracer(full, seed=0, rank=rank, world=world)
    headline = args.workload == 'config2'
    info = tracer.scene.launch_info()
    lib = tracer.lib

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(it):
        tracer.begin_iteration(it)
        tracer.trace(it, keep_images=True)
        tracer.allreduce()

    # ---- FP64 peak (dependent DFMA chains), measured here
    sink = torch.zeros(1, dtype=torch.float64, device=dev)
    flops = C.c_double()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream_ptr = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(lib.xrt_fp64_burn(20000, sink.data_ptr(), C.byref(flops), stream_ptr))
    torch.cuda.synchronize()
    fp64_peaks = []
    for _ in range(3):
        e0.record()
        L.check(lib.xrt_fp64_burn(200000, sink.data_ptr(), C.byref(flops), stream_ptr))
        e1.record()
        torch.cuda.synchronize()
        fp64_peaks.append(flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    fp64_peak = max(fp64_peaks)

    # ---- device-timed steps
    for w in range(args.warmup):
        step(1000 + w)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kstops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (outside the events)
        starts[k].record()
        tracer.begin_iteration(k)           # plasma sources: new bundle table, built on the device (inside the events)
        tracer.trace(k, keep_images=True)
        kstops[k].record()
        tracer.allreduce()
        stops[k].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [starts[k].elapsed_time(stops[k]) for k in range(args.steps)]
    kern_ms = [starts[k].elapsed_time(kstops[k]) for k in range(args.steps)]
    t_dev = torch.tensor([sum(step_ms) * 1e-3, sum(kern_ms) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    t_steps, t_kernel = (float(v) for v in t_dev.cpu())
    launched = tracer.n_rays if tracer.is_plasma else total_rays     # plasma: Poisson total of the last step
    value = launched * args.steps / t_steps

    meta, _ = tracer.counts_and_images(True)
    f_reflect = meta['crystal'] / meta['source']
    n_detected = meta['detector']

    # fraction inside the crystal bounds: same scene with the Bragg test off (1e7 rays, untimed)
    f_bounds, F = None, None
    if headline:
        cfg_b = spectrometer(10_000_000, seed=1)
        cfg_b['optics']['crystal']['check_bragg'] = False
        tb = xrt.Tracer(xconfig.get_config(xconfig.to_numpy(cfg_b)), seed=1)
        tb.trace(0)
        mb, _ = tb.counts_and_images(False)
        f_bounds = mb['crystal'] / mb['source']
        tb.close()
        F = flops_per_ray(f_bounds, f_reflect)
    kernel_rays_per_s = (launched / world) * args.steps / t_kernel
    achieved = kernel_rays_per_s * F / 1e12 if F else None

    # ---- history pass (HBM bound): replay 2^24 ray ids with every element stored
    hist_line = None
    if rank == 0 and headline:
        n_h = 1 << 24
        ids = torch.arange(n_h, dtype=torch.int64, device=dev)
        bufs = tracer.history(0, ids)
        torch.cuda.synchronize()
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
        hist_line = {'bound': 'hbm', 'achieved': bytes_h / t_h / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                     'frac': bytes_h / t_h / 1e9 / hbm_peak, 'traffic': None, 'peak_source': src,
                     'rays': n_h, 'elements': tracer.n_elem, 'ms': t_h * 1e3,
                     'kernel': 'k_record<0, PHILOX>: full replay of every ray + SoA stores'}
        del bufs
    tracer.close()

    # ---- end to end through the public API (host dict in, host dict out)
    for w in range(min(args.warmup, 2)):
        c = workload_config(args.workload, total_rays, seed=50 + w, history=False)
        xicsrt_b200.raytrace(c)
    barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    for k in range(args.steps):
        c = workload_config(args.workload, total_rays, seed=k, history=False)
        res = xicsrt_b200.raytrace(c)
        e2e_rays += int(res['total']['meta']['source']['num_out'])
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = e2e_rays / float(t_e2e.cpu()[0])
    n_elem = 3
    d2h = 8 * (n_elem + 100 * 100 + 100 * 50)
    h2d = C.sizeof(L.XrtSceneDesc)

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and headline:
        cores = os.cpu_count() or 1
        procs = min(cores, 64)
        cpu_reference_step(200000, procs, procs, seed=7)          # warm the pool / imports
        n, dt = cpu_reference_step(1_000_000, 8 * procs, procs, seed=8)
        cpu = {'value': n / dt, 'unit': UNIT, 'cores': procs, 'kind': 'port',
               'sample': f'{8 * procs} runs x 1e6 rays over {procs} processes, {dt:.1f} s '
                         f'(oracle port of the NumPy path; reference scheme xicsrt_multiprocessing)'}

    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (per launch)
    traffic, executed = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'r01_traffic.json')))['k_trace']
        traffic = prof['dram_bytes_per_launch']
        executed = prof.get('executed')         # instruction counts of the same capture (what the kernel really executes)
    except (OSError, KeyError, ValueError):
        pass

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * t_steps / args.steps, 'higher_is_better': True,
            'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD if headline else args.workload, 'rays_per_step': launched, 'rays_per_gpu_per_step': total_rays // world,
                       'history': False, 'images': True, 'parallelism': f'ray-id ranges over {world} GPU(s)',
                       'l2': 'flushed between timed steps (256 MiB memset outside the events); the kernel has no global inputs',
                       'launch': info},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'api': 'xicsrt_b200.raytrace(config)'},
            'gpu_launches': args.steps,
            'roofline': {'bound': 'fp64', 'achieved': achieved, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                         'frac': (achieved / fp64_peak) if achieved else None, 'traffic': traffic,
                         'flop_equiv_per_ray': F, 'f_bounds': f_bounds, 'f_reflect': f_reflect,
                         'peak_source': 'DFMA-chain microbenchmark (xrt_fp64_burn) measured in this run',
                         'kernel': 'k_trace', 'kernel_ms_per_step': 1e3 * t_kernel / args.steps,
                         'executed_ncu': executed if headline else None},
            'roofline_history': hist_line,
            'cpu_baseline': cpu,
            'detected_per_step': n_detected,
            'wall_s_timed_region': t_wall,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--rays', type=float, default=1e9, help='rays per GPU per step')
    ap.add_argument('--scaling', choices=['weak', 'strong'], default='weak')
    ap.add_argument('--impl', choices=['b200', 'reference'], default='b200')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--workload', choices=['config2', 'config3', 'config4', 'config5'], default='config2',
                    help='BASELINE.json config; config2 (default) is the headline')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == '__main__':
    main()
