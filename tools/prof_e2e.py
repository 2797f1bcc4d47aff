"""cProfile of the end-to-end step (pinned host buffers -> host catalog)."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench_workloads
from nemo_b200 import pipelines, device
parDict, tileMaps, truth = bench_workloads.pointSourceConfig(numTiles=16)
config = pipelines.SimpleConfig(parDict, tileMaps)
tb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for _ in range(3):
    cat = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=tb)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    cat = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=tb)
torch.cuda.synchronize()
print("e2e %.3f ms/step, %d rows" % ((time.perf_counter() - t0) / 5 * 1e3, len(cat)))
# pure upload time
batches = config._staged['batches']
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    for b in batches: b.upload()
torch.cuda.synchronize()
print("upload only %.3f ms/step for %.1f MB" % ((time.perf_counter() - t0) / 5 * 1e3, sum(b.h2dBytes for b in batches) / 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    cat = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=tb)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(30); print(s.getvalue()[:6000])
