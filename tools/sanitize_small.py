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
# this round's additions: 'smart' noise bins, percentile estimator, noise level / area tables, float32-staged tiles, a 2-3-5-7-smooth
# side (mixed-radix passes) next to a side with a large prime factor (Bluestein passes)
md, _ = synth.makeTile(shape=(120, 106), numObjects=2, kind='sources', seed=4, margin=30)      # 120 = 5 * 3 * 8, 106 = 2 * 53
yy, xx = np.mgrid[0:120, 0:106]
md[0]['weights'] = md[0]['weights'] * (0.7 + 0.3 * np.sin(yy / 20.0) * np.cos(xx / 25.0))
for noise in [{'method': 'dataMap', 'noiseGridArcmin': 'smart', 'numNoiseBins': 7},
              {'method': 'dataMap', 'noiseGridArcmin': 10., 'numNoiseBins': 2, 'RMSEstimator': 'percentile'}]:
    f = {'label': 'B', 'class': 'BeamMatchedFilter', 'params': {'outputUnits': 'uK', 'edgeTrimArcmin': 0., 'noiseParams': noise}}
    out, obj = filters.filterMaps(md, copy.deepcopy(f), 'T', verbose=False, returnFilter=True)
    from nemo_b200 import completeness
    tab = completeness.getRMSTab(obj.RMSMap, out['surveyMask'], md[0]['wcs'], obj.noiseTable[0])
    print("noise option ok:", noise['noiseGridArcmin'], len(tab), "levels")
tiles32 = {k: [dict(m, data=np.asarray(m['data'], dtype=np.float32), weights=np.asarray(m['weights'], dtype=np.float32)) for m in v]
           for k, v in tileMaps.items()}
cfg32 = pipelines.SimpleConfig(par, tiles32)
opt = pipelines.filterMapsAndMakeCatalogs(cfg32, verbose=False, tilesPerBatch=[1, 2], makeRMSTables=True)
print("float32-staged batches + RMS tables ok", len(opt), len(cfg32.RMSTabs))
torch.cuda.synchronize()
print("ALL OK")
