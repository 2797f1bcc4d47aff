"""Device-resident timing of the batched path for other BASELINE configurations (not the bench line).
    python tools/time_config.py --shape 840 1440 --tiles 8 --freqs 2 --scales 4 --steps 5"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nemo_b200 import pipelines, device, signals, _lib
from nemo_b200.wcs import CARWCS

ap = argparse.ArgumentParser()
ap.add_argument('--shape', type=int, nargs=2, default=[1024, 1024])
ap.add_argument('--tiles', type=int, default=8)
ap.add_argument('--freqs', type=int, default=2)
ap.add_argument('--scales', type=int, default=4)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--precision', default='f64')
ap.add_argument('--kind', default='arnaud')
args = ap.parse_args()
H, W = args.shape
rng = np.random.default_rng(3)
freqs = [(97.8, 60.0, 2.2), (149.6, 40.0, 1.4), (220.0, 90.0, 1.0)][:args.freqs]
tileMaps = {}
for t in range(args.tiles):
    md = []
    for nu, sig, fwhm in freqs:
        beam = signals.gaussianBeam(fwhm)
        md.append({'data': rng.standard_normal((H, W)) * sig, 'weights': np.full((H, W), 1.0 / sig ** 2), 'wcs': CARWCS((H, W), 0.5 / 60),
                   'obsFreqGHz': nu, 'units': 'uK', 'beam': beam, 'beamFileName': beam, 'label': 'f%03d' % int(nu), 'solidAngle_nsr': 0.0,
                   'surveyMask': np.ones((H, W), np.uint8), 'pointSourceMask': np.ones((H, W), np.uint8), 'flagMask': np.zeros((H, W), np.uint8)})
    tileMaps['t%d' % t] = md
masses = [1e14, 2e14, 4e14, 8e14]
zs = [0.2, 0.4, 0.8, 1.2]
filters = []
for i in range(args.scales):
    M, z = masses[i % 4], zs[(i // 4) % 4]
    if args.kind == 'arnaud':
        filters.append({'label': 'Arnaud_M%de14_z%d' % (int(M / 1e14), i), 'class': 'ArnaudModelMatchedFilter',
                        'params': {'M500MSun': M, 'z': z, 'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': 40.}, 'outputUnits': 'yc', 'edgeTrimArcmin': 0.}})
    else:
        filters.append({'label': 'Beam%d' % i, 'class': 'BeamMatchedFilter',
                        'params': {'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': 40.}, 'outputUnits': 'uK', 'edgeTrimArcmin': 0.}})
parDict = {'mapFilters': filters, 'photFilter': filters[0]['label'] if args.scales > 1 else None, 'thresholdSigma': 4.0, 'minObjPix': 1,
           'findCenterOfMass': True, 'useInterpolator': True, 'rejectBorder': 0}
config = pipelines.SimpleConfig(parDict, tileMaps)
ops = device.getOps(args.precision)
filtersList = pipelines._orderedFilters(config.parDict)
batches = pipelines.stageTiles(config, args.tiles, args.precision)
t0 = time.perf_counter()
templates = pipelines.FilterSetTemplates(ops, filtersList, batches[0].mapDicts[0])
pipelines.templateSpectraOf(ops, templates)      # once per geometry, as the drivers do
print("templates (host, once per geometry): %.2f s" % (time.perf_counter() - t0))
for b in batches: b.upload()
def steps(n):
    pipe = pipelines.BatchPipeline(ops, filtersList, templates, config.parDict)
    for _ in range(n):
        for b in batches: pipe.submit(b, upload=False)
    pipe.flush()
steps(4)      # every pipeline slot (three with two batches in flight) captures its CUDA graph before the timed steps
torch.cuda.synchronize(); t0 = time.perf_counter()
steps(args.steps)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / args.steps
mpix = args.tiles * H * W * args.scales / 1e6
print("%dx%d, %d tiles, F=%d, S=%d, %s: %.2f ms/step, %.0f Mpix/s per scale (%.0f Mpix-scales/s)" % (H, W, args.tiles, args.freqs, args.scales, args.precision, dt * 1e3, mpix / dt, mpix / dt))
_lib.profEnable(True)
steps(1)
rep = _lib.profReport()
tot = sum(ms for _, ms in rep.values())
for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:8]:
    print("   %-30s %3d launches %8.3f ms %5.1f%%" % (k, n, ms, 100 * ms / tot))
