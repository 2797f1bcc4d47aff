"""Times the pieces of the multi-rank end-to-end step (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench_workloads
from nemo_b200 import pipelines, device, distributed, catalogs
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
comm = distributed.Comm()
nT = 16
parDict, tileMaps, _ = bench_workloads.pointSourceConfig(numTiles=nT, firstIndex=rank * nT)
config = pipelines.SimpleConfig(parDict, tileMaps)
config.rank = rank
config.allTileNames = ['%d_%d' % (i // 4, i % 4) for i in range(world * nT)]
for _ in range(3):
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=8, comm=comm)
import collections
acc = collections.OrderedDict()
def wrap(mod, name):
    f = getattr(mod, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
    setattr(mod, name, g)
for n in ['allGatherCatalog', 'gatherNoiseTables', 'packTable', 'unpackTable']: wrap(distributed, n)
wrap(distributed.Comm, 'allGatherArrays'); wrap(distributed.Comm, 'maxOverRanks')
wrap(catalogs, 'removeDuplicates'); wrap(catalogs, 'makeOptimalCatalog'); wrap(pipelines, 'processBatches')
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
N = 5
for _ in range(N):
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=8, comm=comm)
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / N * 1e3
if rank == 0:
    print("total %.2f ms/step" % tot)
    for k, v in acc.items(): print("  %-22s %.3f ms" % (k, v / N * 1e3))
dist.destroy_process_group()
