"""Host-side timeline of one pipeline step: how long the launches take, how long the first sync waits, and the rest."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench_workloads
from nemo_b200 import pipelines, device, photometry
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
def step(timing=None):
    t0 = time.perf_counter()
    job = pipelines.startTileBatch(ops, b, filtersList, templates, parDict)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    res = pipelines.finishTileBatch(job)
    t3 = time.perf_counter()
    if timing is not None: timing.append((t1 - t0, t2 - t1, t3 - t2))
for _ in range(3): step()
tm = []
for _ in range(10): step(tm)
import numpy as np
tm = np.array(tm) * 1e3
print("launch phase %.3f ms | wait for GPU %.3f ms | finish phase (GPU idle except small kernels) %.3f ms" % tuple(tm.mean(axis=0)))
pr = cProfile.Profile()
torch.cuda.synchronize()
job = pipelines.startTileBatch(ops, b, filtersList, templates, parDict)
torch.cuda.synchronize()
pr.enable()
for _ in range(1): pipelines.finishTileBatch(job)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:5000])
pr = cProfile.Profile()
pr.enable()
job = pipelines.startTileBatch(ops, b, filtersList, templates, parDict)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(14); print(s.getvalue()[:3000])
