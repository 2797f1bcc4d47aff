"""Where the end-to-end step (pinned host buffers -> host catalog) spends its wall time: the batched device phase
(uploads + kernels + per-batch catalog assembly), the noise tables, the cross-scale merge and, under torchrun, the two gathers.
    python tools/time_e2e_phases.py [batchSizes]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/time_e2e_phases.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench_workloads
from nemo_b200 import pipelines, catalogs, device, distributed
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
comm = None
if world > 1:
    print('rank', rank, 'cpus', distributed.bindProcessToDeviceCpus(local)) if os.environ.get('BIND', '1') == '1' else None
    torch.distributed.init_process_group(backend='nccl', device_id=torch.device('cuda', local))
    comm = distributed.Comm()
if os.environ.get('WORKLOAD', 'dr5') == 'dr5':
    nT = int(os.environ.get('TILES', 35))
    parDict, tileMaps, truth = bench_workloads.dr5Config(numTiles=nT, firstIndex=nT * rank)
    allNames = [bench_workloads.dr5TileName(i) for i in range(world * nT)]
    defaultTb = [2, 4, 8]
else:
    parDict, tileMaps, truth = bench_workloads.pointSourceConfig(numTiles=16, firstIndex=16 * rank)
    allNames = ['%d_%d' % (i // 4, i % 4) for i in range(world * 16)]
    defaultTb = [4]
config = pipelines.SimpleConfig(parDict, tileMaps)
config.rank = rank
config.allTileNames = allNames
tb = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else defaultTb
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
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=tb, comm=comm)
for mod, name in [(pipelines, 'processBatches'), (catalogs, 'makeOptimalCatalog'), (device, 'toHost'), (pipelines, 'finishTileBatch'),
                  (distributed, 'allGatherCatalog'), (distributed, 'gatherNoiseTables'), (distributed, 'packTable'),
                  (distributed, 'unpackTable'), (catalogs, 'removeDuplicates'), (distributed.Comm, 'allGatherArrays'),
                  (pipelines, 'templatesFor'), (pipelines, 'startTileBatch'), (pipelines, 'detectAndMeasure')]:
    timed(mod, name)
torch.cuda.synchronize(); n = int(os.environ.get('STEPS', 5))
if comm is not None: comm.barrier()
t0 = time.perf_counter()
for _ in range(n):
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=tb, comm=comm)
torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / n * 1e3
if rank == 0:
    print("e2e %.3f ms/step with batches %s, %d rank(s)" % (tot, tb, world))
    for k, v in acc.items():
        print("   %-22s %.3f ms/step" % (k, v / n * 1e3))
if comm is not None:
    torch.distributed.destroy_process_group()
