"""Small end-to-end exercise of every kernel family, for compute-sanitizer (memcheck / racecheck) runs."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
from oracle import synth
from nemo_b200 import filters, photometry, pipelines

ge._smoke_run(verbose=True, shape=(128, 128))
# non-power-of-two tile (Bluestein passes), two frequencies, weight bins
md, _ = synth.makeTile(shape=(96, 100), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4), numObjects=2,
                       kind='clusters', seed=3, margin=30)
f = {'label': 'A', 'class': 'ArnaudModelMatchedFilter', 'params': {'M500MSun': 2e14, 'z': 0.4, 'outputUnits': 'yc', 'edgeTrimArcmin': 5.0,
     'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': 10., 'numNoiseBins': 3}}}
out = filters.filterMaps(md, copy.deepcopy(f), 'T', verbose=False)
cat = photometry.findObjects(out, threshold=3.0, minObjPix=1, removeRings=True, ringThresholdSigma=2, findCenterOfMass=True, useInterpolator=True)
print("bluestein tile ok", len(cat))
# batched pipeline, graphs, photFilter, then the injection path
tileMaps = {}
for t in range(3):
    tileMaps['t%d' % t], _ = synth.makeTile(shape=(128, 128), numObjects=3, kind='clusters', seed=50 + t, margin=36)
fl = [{'label': 'A1', 'class': 'ArnaudModelMatchedFilter', 'params': {'M500MSun': 1e14, 'z': 0.6, 'outputUnits': 'yc', 'edgeTrimArcmin': 0.,
       'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': 20.}}},
      {'label': 'A2', 'class': 'ArnaudModelMatchedFilter', 'params': {'M500MSun': 4e14, 'z': 0.3, 'outputUnits': 'yc', 'edgeTrimArcmin': 0.,
       'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': 20.}}}]
par = {'mapFilters': fl, 'photFilter': 'A1', 'thresholdSigma': 3.5, 'minObjPix': 1, 'findCenterOfMass': True, 'useInterpolator': True, 'rejectBorder': 0}
cfg = pipelines.SimpleConfig(par, tileMaps)
for _ in range(2):
    opt = pipelines.filterMapsAndMakeCatalogs(cfg, verbose=False, tilesPerBatch=2)
pipelines.filterMapsAndMakeCatalogs(cfg, verbose=False, tilesPerBatch=2, cacheFilters=True)
opt = pipelines.filterMapsAndMakeCatalogs(cfg, verbose=False, tilesPerBatch=2, useCachedFilters=True, useCachedRMSMap=True, restage=True)
print("batched + injection ok", len(opt))
# real-space filter
md, _ = synth.makeTile(shape=(128, 128), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4), numObjects=2, kind='clusters',
                       cmb=True, seed=8, margin=36)
p = {'M500MSun': 2e14, 'z': 0.4, 'outputUnits': 'yc', 'bckSub': True, 'edgeTrimArcmin': 0.,
     'noiseParams': {'method': 'dataMap', 'matchedFilterClass': 'ArnaudModelMatchedFilter', 'RADecSection': None, 'kernelMaxArcmin': 5.,
                     'symmetrize': True, 'noiseGridArcmin': 20.}}
out = filters.filterMaps(md, {'label': 'RS', 'class': 'ArnaudModelRealSpaceMatchedFilter', 'params': p}, 'T', verbose=False)
print("real-space ok")
torch.cuda.synchronize()
print("ALL OK")
