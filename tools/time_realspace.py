"""Times the real-space matched filter class path (BASELINE config 3 shape: 90 + 150 GHz, kernelMaxArcmin 7) per tile."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nemo_b200 import filters, photometry, signals, _lib
from nemo_b200.wcs import CARWCS
H = W = 1024
rng = np.random.default_rng(5)
def tile(seed):
    rng = np.random.default_rng(seed)
    md = []
    for nu, sig, fwhm in [(97.8, 60.0, 2.2), (149.6, 40.0, 1.4)]:
        beam = signals.gaussianBeam(fwhm)
        md.append({'data': rng.standard_normal((H, W)) * sig, 'weights': np.full((H, W), 1.0 / sig ** 2), 'wcs': CARWCS((H, W), 0.5 / 60),
                   'obsFreqGHz': nu, 'units': 'uK', 'beam': beam, 'beamFileName': beam, 'label': 'f%03d' % int(nu), 'solidAngle_nsr': 0.0,
                   'surveyMask': np.ones((H, W), np.uint8), 'pointSourceMask': np.ones((H, W), np.uint8), 'flagMask': np.zeros((H, W), np.uint8)})
    return md
p = {'M500MSun': 2e14, 'z': 0.4, 'outputUnits': 'yc', 'bckSub': True, 'edgeTrimArcmin': 0.,
     'noiseParams': {'method': 'dataMap', 'matchedFilterClass': 'ArnaudModelMatchedFilter', 'RADecSection': None,
                     'kernelMaxArcmin': 7., 'symmetrize': True, 'noiseGridArcmin': 20.}}
f = {'label': 'ArnaudRS', 'class': 'ArnaudModelRealSpaceMatchedFilter', 'params': p}
import copy
d = tempfile.mkdtemp()
for it in range(3):
    md = tile(it)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = filters.filterMaps(md, copy.deepcopy(f), 'T%d' % it, diagnosticsDir=d, selFnDir=d, verbose=False)
    cat = photometry.findObjects(out, threshold=4.0, minObjPix=1, findCenterOfMass=True, removeRings=True, ringThresholdSigma=3, useInterpolator=True, rejectBorder=0)
    torch.cuda.synchronize()
    print("tile %d: %.1f ms (kernel build + apply + detect), %d objects" % (it, (time.perf_counter() - t0) * 1e3, len(cat)))
_lib.profEnable(True)
md = tile(9)
out = filters.filterMaps(md, copy.deepcopy(f), 'T9', diagnosticsDir=d, selFnDir=d, verbose=False)
rep = _lib.profReport()
tot = sum(ms for _, ms in rep.values())
print("device kernels %.2f ms" % tot)
for k, (n, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1])[:8]:
    print("   %-30s %3d launches %8.3f ms %5.1f%%" % (k, n, ms, 100 * ms / tot))
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable()
md = tile(11)
out = filters.filterMaps(md, copy.deepcopy(f), 'T11', diagnosticsDir=d, selFnDir=d, verbose=False)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(22); print(s.getvalue()[:4500])
