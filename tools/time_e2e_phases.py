"""Where the end-to-end step (pinned host buffers -> host catalog) spends its wall time: the batched device phase
(uploads + kernels + per-batch catalog assembly), the noise tables, the cross-scale merge."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench_workloads
from nemo_b200 import pipelines, catalogs, device
parDict, tileMaps, truth = bench_workloads.pointSourceConfig(numTiles=16)
config = pipelines.SimpleConfig(parDict, tileMaps)
tb = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else [2, 4]
tb = tb[0] if len(tb) == 1 else tb
acc = {}
def timed(mod, name):
    fn = getattr(mod, name)
    def wrap(*a, **k):
        t0 = time.perf_counter()
        out = fn(*a, **k)
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return out
    setattr(mod, name, wrap)
for _ in range(3):
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=tb)
timed(pipelines, 'processBatches'); timed(catalogs, 'makeOptimalCatalog'); timed(device, 'toHost'); timed(pipelines, 'finishTileBatch')
torch.cuda.synchronize(); n = 20; t0 = time.perf_counter()
for _ in range(n):
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=tb)
torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / n * 1e3
print("e2e %.3f ms/step with batches %s" % (tot, tb))
for k, v in acc.items():
    print("   %-22s %.3f ms/step" % (k, v / n * 1e3))
