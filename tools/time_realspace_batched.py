"""BASELINE configs[2] shape through the batched real-space driver (nemo_b200/realspace.py): 64 tiles of 1024 x 1024 at
90 + 150 GHz, one ArnaudModelRealSpaceMatchedFilter (kernelMaxArcmin 7 -> 29 x 29 kernel), tiles resident on the device.
    python tools/time_realspace_batched.py [--tiles 64] [--tiles-per-batch 16] [--steps 3]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nemo_b200 import pipelines, realspace, device, signals, _lib
from nemo_b200.wcs import CARWCS
ap = argparse.ArgumentParser()
ap.add_argument('--tiles', type=int, default=64)
ap.add_argument('--tiles-per-batch', type=int, default=16)
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--size', type=int, default=1024)
args = ap.parse_args()
H = W = args.size
tileMaps = {}
for t in range(args.tiles):
    rng = np.random.default_rng(100 + t)
    md = []
    for nu, sig, fwhm in [(97.8, 60.0, 2.2), (149.6, 40.0, 1.4)]:
        beam = signals.gaussianBeam(fwhm)
        md.append({'data': rng.standard_normal((H, W)) * sig, 'weights': np.full((H, W), 1.0 / sig ** 2), 'wcs': CARWCS((H, W), 0.5 / 60),
                   'obsFreqGHz': nu, 'units': 'uK', 'beam': beam, 'beamFileName': beam, 'label': 'f%03d' % int(nu), 'solidAngle_nsr': 0.0,
                   'surveyMask': np.ones((H, W), np.uint8), 'pointSourceMask': np.ones((H, W), np.uint8), 'flagMask': np.zeros((H, W), np.uint8)})
    tileMaps['t%02d' % t] = md
p = {'M500MSun': 2e14, 'z': 0.4, 'outputUnits': 'yc', 'bckSub': True, 'bckSubScaleArcmin': 14.0, 'edgeTrimArcmin': 0.,
     'noiseParams': {'method': 'dataMap', 'matchedFilterClass': 'ArnaudModelMatchedFilter', 'RADecSection': None,
                     'kernelMaxArcmin': 7., 'symmetrize': True, 'noiseGridArcmin': 20.}}
f = {'label': 'ArnaudRS', 'class': 'ArnaudModelRealSpaceMatchedFilter', 'params': p}
parDict = {'mapFilters': [f], 'photFilter': None, 'thresholdSigma': 4.0, 'minObjPix': 1, 'findCenterOfMass': True, 'useInterpolator': True, 'rejectBorder': 0}
config = pipelines.SimpleConfig(parDict, tileMaps)
ops = device.getOps('f64')
filtersList = pipelines._orderedFilters(config.parDict)
assert realspace.canBatch(filtersList, tileMaps)
batches = pipelines.stageTiles(config, args.tiles_per_batch, 'f64')
for b in batches: b.upload()
tpl = [pipelines.templatesFor(config, ops, filtersList, b) for b in batches]
def step():
    n = 0
    for b, t in zip(batches, tpl):
        job = realspace.startRealSpaceTileBatch(ops, b, filtersList, t, config.parDict)
        cd, _ = pipelines.finishTileBatch(job)
        n += sum(len(v['catalog']) for v in cd.values())
    return n
step(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(args.steps): n = step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / args.steps
mpix = args.tiles * H * W / 1e6
print("real-space batched, %d tiles of %dx%d, F=2, S=1, device-resident: %.1f ms/step, %.0f Mpix/s, %d objects" % (args.tiles, H, W, dt * 1e3, mpix / dt, n))
_lib.profEnable(True)
step()
rep = _lib.profReport()
tot = sum(ms for _, ms in rep.values())
print("device kernels %.2f ms" % tot)
for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:12]:
    print("   %-32s %3d launches %8.3f ms %5.1f%%" % (k, c, ms, 100 * ms / tot))
# fp64 multiply-add rate of the convolution: per batch one launch over the maps (kh kw F multiply-adds per pixel) and one over
# the central window of the templates
job = realspace.startRealSpaceTileBatch(ops, batches[0], filtersList, tpl[0], config.parDict)
kh, kw = job['kern2d'].shape[-2:]
pipelines.finishTileBatch(job)
n, ms = rep.get('conv2d_kernel', (0, 0.0))
if n:
    win = 2 * realspace.TEMPLATE_WINDOW_RADIUS + 1 + 2 * (kh // 2)
    fma = args.tiles * 2 * kh * kw * (H * W + win * win)
    print("conv2d (%d x %d kernels): %.2f TFMA/s fp64 of the 18.6 TFMA/s of 148 SMs x 64 DFMA/clk x 1.965 GHz" % (kh, kw, fma / (ms * 1e-3) / 1e12))
