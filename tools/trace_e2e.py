"""GPU timeline of one end-to-end step (torch.profiler / CUPTI): busy time, idle gaps and what runs around them.
    python tools/trace_e2e.py [batchSizes]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench_workloads
from torch.profiler import profile, ProfilerActivity
from nemo_b200 import pipelines
nT = int(os.environ.get('TILES', 35))
parDict, tileMaps, truth = bench_workloads.dr5Config(numTiles=nT, firstIndex=0)
config = pipelines.SimpleConfig(parDict, tileMaps)
config.allTileNames = [bench_workloads.dr5TileName(i) for i in range(nT)]
tb = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else [2, 4, 8]
tb = tb[0] if len(tb) == 1 else tb
for _ in range(3):
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=tb)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=tb)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
path = '/tmp/trace_e2e.json'
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))['traceEvents'] if e.get('ph') == 'X' and e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
ev.sort(key=lambda e: e['ts'])
print("wall %.1f ms under the profiler; %d GPU activities" % (wall, len(ev)))
start, end = ev[0]['ts'], max(e['ts'] + e['dur'] for e in ev)
# union of kernel intervals (compute busy), and of memcpy intervals
def union(items):
    tot, cur_s, cur_e, gaps = 0.0, None, None, []
    for e in items:
        s, t = e['ts'], e['ts'] + e['dur']
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                tot += cur_e - cur_s
                gaps.append((cur_e, s))
            cur_s, cur_e = s, t
        else:
            cur_e = max(cur_e, t)
    tot += cur_e - cur_s
    return tot, gaps
kern = [e for e in ev if e['cat'] == 'kernel']
busy, gaps = union(kern)
cp, _ = union([e for e in ev if e['cat'] == 'gpu_memcpy'])
print("GPU span %.1f ms, kernels busy %.1f ms (sum of kernel durations %.1f ms), memcpy busy %.1f ms" % ((end - start) / 1e3, busy / 1e3, sum(e['dur'] for e in kern) / 1e3, cp / 1e3))
big = sorted(gaps, key=lambda g: g[0] - g[1])[:12]
print("largest compute-idle gaps (ms into the step, length ms, kernel after):")
for s, t in sorted(big):
    nxt = next((e['name'][:60] for e in kern if e['ts'] >= t), '')
    print("   %7.2f  %6.2f  %s" % ((s - start) / 1e3, (t - s) / 1e3, nxt))
# per-stream timeline: runs of activities on one stream separated by less than 0.2 ms, with their time span and what they are
byStream = {}
for e in ev:
    byStream.setdefault(e.get('args', {}).get('stream', e.get('tid')), []).append(e)
print("timeline (stream, start ms, end ms, busy ms, n, first .. last activity):")
runs = []
for sid, items in byStream.items():
    cur = None
    for e in items:
        s, t = e['ts'], e['ts'] + e['dur']
        kind = e['cat']
        if cur is not None and s - cur['end'] < 200 and kind == cur['kind']:
            cur['end'] = max(cur['end'], t); cur['busy'] += e['dur']; cur['n'] += 1; cur['last'] = e['name']
            if kind == 'gpu_memcpy':
                cur['bytes'] += e.get('args', {}).get('bytes', 0)
        else:
            cur = {'sid': sid, 'start': s, 'end': t, 'busy': e['dur'], 'n': 1, 'first': e['name'], 'last': e['name'], 'kind': kind,
                   'bytes': e.get('args', {}).get('bytes', 0) if kind == 'gpu_memcpy' else 0}
            runs.append(cur)
for r in sorted(runs, key=lambda r: r['start']):
    if r['end'] - r['start'] < 30 and r['n'] < 3:
        continue
    extra = (" %.1f MB" % (r['bytes'] / 1e6)) if r['bytes'] else ''
    print("  s%-4s %7.2f %7.2f  busy %6.2f  n %4d  %s .. %s%s" % (r['sid'], (r['start'] - start) / 1e3, (r['end'] - start) / 1e3, r['busy'] / 1e3,
                                                               r['n'], r['first'][:38], r['last'][:38], extra))
# host side: the python-level calls of the step
cpu = [e for e in json.load(open(path))['traceEvents'] if e.get('ph') == 'X' and e.get('cat') in ('cuda_runtime',) and e.get('name') in
       ('cudaStreamSynchronize', 'cudaGraphLaunch', 'cudaEventSynchronize', 'cudaDeviceSynchronize')]
cpu.sort(key=lambda e: e['ts'])
print("host: graph launches and synchronisations (ms into the step, duration ms):")
for e in cpu:
    if e['name'] == 'cudaGraphLaunch' or e['dur'] > 300:
        print("   %7.2f %6.2f %s" % ((e['ts'] - start) / 1e3, e['dur'] / 1e3, e['name']))
