"""Wall-clock breakdown of the host 'finish' phase by wrapping its building blocks with timers."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench_workloads
from nemo_b200 import pipelines, device, photometry, catalogs, ops as opsmod
nT = int(os.environ.get('TILES', 8))
parDict, tileMaps, truth = (bench_workloads.dr5Config(numTiles=nT) if os.environ.get('WORKLOAD', 'dr5') == 'dr5' else bench_workloads.pointSourceConfig(numTiles=nT))
config = pipelines.SimpleConfig(parDict, tileMaps)
parDict = config.parDict
ops = device.getOps('f64')
filtersList = pipelines._orderedFilters(parDict)
batches = pipelines.stageTiles(config, nT, 'f64')
templates = pipelines.FilterSetTemplates(ops, filtersList, batches[0].mapDicts[0])
for b in batches: b.upload()
b = batches[0]
acc = collections.OrderedDict()
def wrap(mod, name, label=None):
    f = getattr(mod, name)
    label = label or name
    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        acc[label] = acc.get(label, 0.0) + time.perf_counter() - t0
        return r
    setattr(mod, name, g)
wrap(photometry, '_collectSegments'); wrap(photometry, '_flatObjects'); wrap(photometry, '_gatherLabels'); wrap(photometry, 'splitObjects')
wrap(photometry, 'finishFindObjectsBatch'); wrap(device, 'toHost'); wrap(photometry, '_galacticLat')
wrap(opsmod.Ops, 'spline_eval'); wrap(opsmod.Ops, 'to_device'); wrap(pipelines, 'detectAndMeasure')
def step():
    job = pipelines.startTileBatch(ops, b, filtersList, templates, parDict)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pipelines.finishTileBatch(job)
    acc['TOTAL finish'] = acc.get('TOTAL finish', 0.0) + time.perf_counter() - t0
for _ in range(3): step()
acc.clear()
N = 20
for _ in range(N): step()
for k, v in acc.items(): print("%-28s %.3f ms" % (k, v / N * 1e3))
