import sys, json
sys.path.insert(0, '.')
import torch, bench
from xicsrt_b200 import _driver, config as xconfig
def run(cfg, label, n=int(1e9)):
    tr = _driver.Tracer(xconfig.get_config(xconfig.to_numpy(cfg)), 0)
    for it in range(2): tr.trace(it)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(3): tr.trace(it)
    e1.record(); torch.cuda.synchronize()
    meta, _ = tr.counts_and_images(True)
    print(label, tr.scene.launch_info(), 'rays/s %.3e' % (3 * n / (e0.elapsed_time(e1) * 1e-3)), meta)
    tr.close()
n = int(1e9)
cfg = bench.spectrometer(n)
run(cfg, 'point source (spectrometer variant)')
cfg = bench.spectrometer(n); cfg['sources']['source'].update({'xsize': 1e-3, 'ysize': 1e-3, 'zsize': 1e-3})
run(cfg, 'box source 1 mm (generic lean)')
cfg = bench.spectrometer(n); cfg['optics']['crystal']['rocking_type'] = 'step'
run(cfg, 'step rocking curve (spectrometer variant)')
cfg = bench.spectrometer(n); cfg['sources']['source']['velocity'] = [0.0, 0.0, 1e4]
run(cfg, 'Doppler-shifted line (lean extended-source variant, deferred deviate)')
