import sys, cProfile, pstats, io
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch, bench_workloads
from nemo_b200 import pipelines, device
parDict, tileMaps, truth = bench_workloads.pointSourceConfig(numTiles=16)
config = pipelines.SimpleConfig(parDict, tileMaps)
parDict = config.parDict
ops = device.getOps('f64')
filtersList = pipelines._orderedFilters(parDict)
batches = pipelines.stageTiles(config, 16, 'f64')
templates = pipelines.FilterSetTemplates(ops, filtersList, batches[0].mapDicts[0])
for b in batches: b.upload()
def step():
    for b in batches:
        pipelines.filterTileBatch(ops, b, filtersList, templates, parDict, undoPixelWindow=True)
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:9000])
