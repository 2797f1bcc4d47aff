"""Config 5 shape: the re-filtering half of a source-injection loop (maps.py:2093, pipelines.py:216-232): filters, normalisations and
RMS maps cached on the device once, then every realisation uploads the edited maps and runs K1, K4, K7 (cached RMS), K8-K11."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench_workloads
from nemo_b200 import pipelines
nT = int(sys.argv[1]) if len(sys.argv) > 1 else 16
parDict, tileMaps, truth = bench_workloads.pointSourceConfig(numTiles=nT)
parDict['photFilter'] = parDict['mapFilters'][0]['label']
# a survey mask with a closed border, as real tiles have: with the cached RMS map the reference (and this path) takes S/N = data where
# the cached RMS is zero (pipelines.py:226-229), i.e. raw uK values inside the 20-pixel apodisation border of an unmasked tile
for md in tileMaps.values():
    m = md[0]['surveyMask']
    m[:32], m[-32:], m[:, :32], m[:, -32:] = 0, 0, 0, 0
config = pipelines.SimpleConfig(parDict, tileMaps)
t0 = time.perf_counter()
pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=4, cacheFilters=True)
torch.cuda.synchronize()
print("first pass with cacheFilters: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
rng = np.random.default_rng(0)
original = {k: md[0]['data'].copy() for k, md in tileMaps.items()}
from nemo_b200 import signals
beam = signals.gaussianBeam(1.4)
def inject():
    # stand-in for the painter (out of scope, host): 50 beam-shaped sources of 300-900 uK per tile on top of the original map
    for k, md in tileMaps.items():
        xs, ys = rng.uniform(100, 900, 50), rng.uniform(100, 900, 50)
        np.copyto(md[0]['data'], original[k] + bench_workloads._paintSources((1024, 1024), xs, ys, rng.uniform(300, 900, 50), beam))
for _ in range(2):
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=4, useCachedFilters=True, useCachedRMSMap=True, restage=True)
torch.cuda.synchronize()
n = 10
tInject = tRun = tStage = 0.0
batches = config._staged['batches']
for _ in range(n):
    t0 = time.perf_counter(); inject(); t1 = time.perf_counter()
    for b in batches:
        b.refresh(config.tileMaps)            # host copy of the edited maps into the pinned staging buffers
    ts = time.perf_counter()
    cat = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=4, useCachedFilters=True, useCachedRMSMap=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    tInject += t1 - t0; tStage += ts - t1; tRun += t2 - ts
mpix = nT * 1024 * 1024 / 1e6
print("host painter: per realisation, %d tiles: painting on the host %.2f ms; host copy into the pinned buffers %.2f ms; upload + re-filter + detect + "
      "catalogs %.2f ms; %.2f ms in all = %.0f Mpix/s; %d rows"
      % (nT, tInject / n * 1e3, tStage / n * 1e3, tRun / n * 1e3, (tInject + tStage + tRun) / n * 1e3, mpix / ((tInject + tStage + tRun) / n), len(cat)))
if os.environ.get('PROFILE'):
    import cProfile, pstats, io
    pr = cProfile.Profile(); pr.enable()
    for _ in range(3):
        cat = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=4, useCachedFilters=True, useCachedRMSMap=True)
    torch.cuda.synchronize(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:5000])

# ---- the same loop with the painter on the device: maps resident, only the catalogs travel ---------------------------------
for b in batches:
    b.dev, b.devOriginal = None, None
for k, md in tileMaps.items():
    np.copyto(md[0]['data'], original[k])
def mockCatalogs():
    cats = {}
    for k, md in tileMaps.items():
        ra, dec = md[0]['wcs'].pix2wcs(rng.uniform(100, 900, 50), rng.uniform(100, 900, 50))
        cats[k] = {'RADeg': np.asarray(ra), 'decDeg': np.asarray(dec), 'deltaT_c': rng.uniform(300, 900, 50)}
    return cats
for _ in range(3):
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=4, useCachedFilters=True, useCachedRMSMap=True, injectCatalogs=mockCatalogs())
torch.cuda.synchronize()
N = int(os.environ.get('REALISATIONS', 100))
from nemo_b200 import device
device.TRAFFIC['h2d'] = device.TRAFFIC['d2h'] = 0
t0 = time.perf_counter()
rows = 0
for _ in range(N):
    cat = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=4, useCachedFilters=True, useCachedRMSMap=True,
                                              injectCatalogs=mockCatalogs())
    rows += len(cat)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / N
print("device painter: %d realisations x %d tiles, %.2f ms per realisation = %.0f Mpix/s; %.1f kB host-to-device and %.1f kB device-to-host per "
      "realisation; %d rows per realisation" % (N, nT, dt * 1e3, mpix / dt, device.TRAFFIC['h2d'] / N / 1e3, device.TRAFFIC['d2h'] / N / 1e3, rows // N))
