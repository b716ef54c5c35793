"""Device-timed rays/s of bench workloads for the library named by XRT_LIB_PATH (development measurements).
usage: python tests/scripts/quick_rate.py [config2 box focused doppler step config3 config4 config5 ...] [--rays N]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from xicsrt_b200 import _driver, config as xconfig


def make(name, n):
    if name in ('config2', 'config3', 'config4', 'config5'):
        return bench.workload_config(name, n)
    if name.startswith('scene:'):     # a parity scene of oracle/scenes.py at n rays, history off
        from oracle import scenes
        cfg = scenes.get(name[6:])
        cfg['sources']['source']['intensity'] = n
        cfg['general']['keep_history'] = False
        return cfg
    cfg = bench.spectrometer(n)
    if name == 'box':
        cfg['sources']['source'].update({'xsize': 1e-3, 'ysize': 1e-3, 'zsize': 1e-3})
    elif name == 'focused':
        cfg['sources']['source'].update({'class_name': 'XicsrtSourceFocused', 'target': [0.0, 0.0, 0.80374151],
                                         'xsize': 0.02, 'ysize': 0.02, 'zsize': 0.02})
    elif name == 'doppler':
        cfg['sources']['source']['velocity'] = [0.0, 0.0, 1e4]
    elif name == 'plasma_mesh':       # the config 5 plasma in front of the config 4 mesh crystal (Bragg test on)
        cfg = bench.workload_config('config5', n)
        cfg['optics']['crystal'] = bench.workload_config('config4', n)['optics']['crystal']
        cfg['optics']['crystal'].update({'check_bragg': True, 'rocking_fwhm': 2e-3})
    elif name == 'step':
        cfg['optics']['crystal']['rocking_type'] = 'step'
    elif name == 'nocull':
        os.environ['XRT_NO_CULL'] = '1'
    elif name == 'nobroad':
        os.environ['XRT_NO_BROAD32'] = '1'
    else:
        raise KeyError(name)
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('names', nargs='*', default=['config2'])
    ap.add_argument('--rays', type=float, default=1e9)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--flush', action='store_true', help='256 MiB memset before every timed step (as bench.py); its time is included')
    args = ap.parse_args()
    for name in args.names:
        n = int(args.rays if name not in ('config3', 'config4', 'plasma_mesh') and not name.startswith('scene:') else min(args.rays, 1e8))
        cfg = make(name, n)
        tr = _driver.Tracer(xconfig.get_config(xconfig.to_numpy(cfg)), 0)
        for it in range(2):
            tr.begin_iteration(it); tr.trace(it)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda') if args.flush else None
        for it in range(args.steps):
            if flush is not None:
                flush.zero_()
            tr.begin_iteration(it); tr.trace(it)
        e1.record(); torch.cuda.synchronize()
        meta, _ = tr.counts_and_images(True)
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({'lib': os.path.basename(os.environ.get('XRT_LIB_PATH', 'libxrt.so')), 'workload': name,
                          'rays_per_s': tr.n_rays / (ms * 1e-3), 'ms': ms, 'launch': tr.scene.launch_info(), 'meta': meta}), flush=True)
        for k in ('XRT_NO_CULL', 'XRT_NO_BROAD32'):
            os.environ.pop(k, None)
        tr.close()


if __name__ == '__main__':
    main()
