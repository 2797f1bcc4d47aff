#!/usr/bin/env python
"""
bench.py -- throughput of the per-tile filter -> noise/S-N -> detect -> photometry hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): a 4096 x 4096 synthetic 148 GHz map in 16 tiles of 1024 x 1024 per GPU,
BeamMatchedFilter (noiseGridArcmin 40, numNoiseBins 8, uK), S/N > 4 detection, ring removal, spline photometry
(bench_workloads.pointSourceConfig).  One step = one pass of the whole path over the rank's 16 tiles.  Tiles are
independent, so with N ranks every rank filters its own 16 tiles (weak scaling) and the only collective is the
catalog / noise-table gather at the end of the end-to-end step.

Metric: filtered+detected Mpix/s per filter scale = tiles * H * W * S / step time (whole job, all ranks).
  value : inputs resident in HBM when the timed region starts (CUDA events, max over ranks)
  e2e   : pipelines.filterMapsAndMakeCatalogs from pinned HOST buffers to the host catalog, copies inside the timed region
  roofline     : the kernel with the largest share of the step, algorithmic bytes / its CUDA-event time vs measured HBM peak
  cpu_baseline : the numpy/scipy oracle (the reference's arithmetic) on a bounded sample of the same tiles, on the host cores

--impl reference times that CPU path alone (all host cores, one tile per worker process per step).
"""

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "filtered+detected Mpix/s per filter scale"
UNIT = "Mpix/s"
SHAPE = (1024, 1024)
TILES_PER_GPU = 16


# ------------------------------------------------------------------------------------------------------
# CPU baseline (the oracle): one tile per worker process
def _cpuTile(index):
    """Worker: regenerates tile `index` (same seed as the GPU arm) and runs the reference arithmetic on it."""
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    import bench_workloads
    from oracle import filters_ref, photometry_ref
    parDict, tileMaps, _ = bench_workloads.pointSourceConfig(numTiles=1, shape=SHAPE, firstIndex=index)
    name, mapDicts = next(iter(tileMaps.items()))
    f = parDict['mapFilters'][0]
    t0 = time.perf_counter()
    filtered = filters_ref.filterMaps(mapDicts, {'label': f['label'], 'class': f['class'], 'params': dict(f['params'])}, name)
    cat = photometry_ref.findObjects(filtered, threshold=parDict['thresholdSigma'], minObjPix=parDict['minObjPix'],
                                     rejectBorder=parDict['rejectBorder'], findCenterOfMass=parDict['findCenterOfMass'],
                                     removeRings=parDict['removeRings'], ringThresholdSigma=parDict['ringThresholdSigma'],
                                     objIdent=parDict['objIdent'], longNames=parDict['longNames'],
                                     useInterpolator=parDict['useInterpolator'])
    photometry_ref.measureFluxes(cat, filtered, None, useInterpolator=parDict['useInterpolator'])
    return time.perf_counter() - t0, len(cat)


def _cpuPool(cores):
    import multiprocessing as mp
    os.environ['OMP_NUM_THREADS'] = '1'
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    os.environ['MKL_NUM_THREADS'] = '1'
    return mp.get_context('spawn').Pool(cores)


def cpuBaseline(cores=None, tiles=None, steps=1, warmup=0):
    """Times the oracle over `tiles` tiles per step spread over `cores` worker processes (the reference's MPI-over-tiles
    split, startUp.py:390-404).  Tile generation happens inside the workers but outside their timers; the step time
    is the slowest worker's summed compute time (what an MPI job's wall clock would be)."""
    cores = cores or os.cpu_count() or 1
    tiles = tiles or cores
    pool = _cpuPool(min(cores, tiles))
    try:
        times = []
        nObj = 0
        for it in range(warmup + steps):
            res = pool.map(_cpuTile, [i % TILES_PER_GPU for i in range(tiles)], chunksize=1)
            # wall of the step: tiles are dealt round-robin to `cores` workers
            perWorker = [0.0] * min(cores, tiles)
            for i, (dt, n) in enumerate(res):
                perWorker[i % len(perWorker)] += dt
            if it >= warmup:
                times.append(max(perWorker))
                nObj = sum(n for _, n in res)
    finally:
        pool.close()
        pool.join()
    tStep = sum(times) / len(times)
    mpix = tiles * SHAPE[0] * SHAPE[1] / 1e6
    return {'value': mpix / tStep, 'unit': UNIT, 'cores': min(cores, tiles), 'kind': 'port',
            'sample': "%d of the workload's 1024x1024 tiles per step, one per worker process, %d step(s); numpy/scipy oracle "
                      "(np.fft for pixell, vectorised np.linalg.inv), %.2f s per step, %d objects" % (tiles, steps, tStep, nObj)}, tStep


# ------------------------------------------------------------------------------------------------------
def algorithmicBytes(nT, F, S, H, W, E, rings=True):
    """Algorithmic HBM bytes per step of every kernel of the batched Fourier path: own inputs read once + own outputs
    written once (SURVEY.md 8d / DESIGN.md), E = bytes per real (8 for f64, 4 for f32), complex = 2E."""
    N, Nc = H * W, H * (W // 2 + 1)
    P = F * (F + 1) // 2
    nOut = nT * S
    nPass = 2 if rings else 1
    b = {}
    imgsFwd = nT * F + S * F + nOut          # data, unit templates, pixel-window undo
    b['rfft_rows_kernel'] = imgsFwd * (N * E + Nc * 2 * E)
    b['fft_cols_kernel'] = (nT * F + S * F) * 2 * Nc * 2 * E
    b['xpower_smooth_x_kernel'] = nT * (F * Nc * 2 * E + P * N * 8)
    b['smooth_y_f64_kernel'] = nT * P * N * 8 * 2
    b['build_filter_kernel'] = nT * P * N * 8 + nOut * F * Nc * 2 * E + nOut * F * Nc * E
    b['fft_cols_apply_kernel'] = 2 * nOut * (F * Nc * 2 * E + F * Nc * E + Nc * 2 * E)     # templates + maps
    b['irfft_rows_kernel'] = nOut * (Nc * 2 * E + N * E) + nOut * (Nc * 2 * E + N * E + N) + nOut * (Nc * 2 * E + 2 * N * E)
    b['template_argmax_partial_kernel'] = nOut * N * E
    b['cell_rms_onchip_kernel'] = nOut * 2 * N * E
    b['cell_rms_hybrid_kernel'] = nOut * 2 * N * E
    b['cell_rms_stream_kernel'] = nOut * 2 * N * E
    b['paint_rms_kernel'] = nOut * 2 * N * E
    b['edge_rows_kernel'] = nOut * (N * E + 2 * N)
    b['edge_cols_final_kernel'] = nOut * (2 * N * E + 3 * N + 3 * N * E + N)
    b['edge_cols_final_paint_kernel'] = nOut * (2 * N * E + 3 * N + 3 * N * E + N)   # map, median weights, masks in; data, S/N, RMS, mask out
    b['fft_cols_window_kernel'] = nOut * 2 * Nc * 2 * E
    b['detect_scan_kernel'] = nPass * nOut * (N * E + 4 * N)      # S/N read, zeroed label image written; the rest is sparse
    return b


def endToEndBytesPerPixel(F, S, E, rings=True):
    """SURVEY.md 8d end-to-end figure B(F,S) = 8F/S + 18F + 85 (+16) at 4-byte reals, with reals of E bytes:
    every 4-byte term scales with E/4 except masks (1 B) and labels (4 B)."""
    r = E / 4.0
    k1 = 8 * F / S * r
    k23 = 8 * F * r
    k4 = 6 * F * r + 4 * r + 1
    k6 = (4 * F + 8) * r
    k7 = 20 * r + 3
    k8 = 16 * r + 1
    k9 = 4 * r + 4
    k10 = 4 * r + 4
    k11 = 16 * r
    total = k1 + k23 + k4 + k6 + k7 + k8 + k9 + k10 + k11
    if rings:
        total += k9 + k10
    return total


class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpuIndex):
        self.proc = None
        self.gpuIndex = gpuIndex

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpuIndex), '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                                          '-lms', '50'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        sm, smMax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.splitlines():
            p = [x.strip() for x in line.split(',')]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smMax.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(smMax) if smMax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------------
def runReference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    base, tStep = cpuBaseline(cores=cores, tiles=min(cores, TILES_PER_GPU), steps=args.steps, warmup=args.warmup)
    line = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': tStep * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'configs[1]: BeamMatchedFilter + S/N>4 detection, 1024x1024 tiles of a 4096x4096 148 GHz map',
                       'tile': list(SHAPE), 'freqs': 1, 'scales': 1, 'tiles_per_step': min(cores, TILES_PER_GPU)},
            'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default='f64', choices=['f64', 'f32'])
    ap.add_argument('--tiles', type=int, default=TILES_PER_GPU, help='tiles per GPU')
    ap.add_argument('--tiles-per-batch', type=int, default=16)
    ap.add_argument('--e2e-tiles-per-batch', default='4', help='batch sizes of the end-to-end arm (uploads overlap compute): one size, or a list that is mirrored at the end (2,4 -> 2,4,4,...,4,2)')
    ap.add_argument('--pipeline-depth', type=int, default=2, help='batches in flight on the device behind the one whose catalogs the host assembles (device-resident arm)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-f32-host', action='store_true', help='skip the extra end-to-end arm with float32 host tiles')
    ap.add_argument('--kernel-table', default=None, help='write the per-kernel roofline table to this file')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    localRank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        runReference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    e2eBatch = [int(v) for v in str(args.e2e_tiles_per_batch).split(',')]
    e2eBatch = e2eBatch[0] if len(e2eBatch) == 1 else e2eBatch

    import numpy as np
    import torch
    import bench_workloads
    from nemo_b200 import _lib, device, pipelines, distributed

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path.  Use --impl reference for the CPU baseline.")
    torch.cuda.set_device(localRank)
    comm = None
    if world > 1:
        distributed.bindProcessToDeviceCpus(localRank)     # pinned staging buffers next to this rank's GPU
        import torch.distributed as dist
        dist.init_process_group(backend='nccl', device_id=torch.device('cuda', localRank))
        comm = distributed.Comm()

    def barrier():
        if comm is not None:
            comm.barrier()
        torch.cuda.synchronize()

    def maxOverRanks(v):
        return comm.maxOverRanks(v) if comm is not None else v

    H, W = SHAPE
    nT = args.tiles
    parDict, tileMaps, truth = bench_workloads.pointSourceConfig(numTiles=nT, shape=SHAPE, firstIndex=rank * nT)
    config = pipelines.SimpleConfig(parDict, tileMaps, rank=0, size=1)     # every rank owns all of its own tiles
    config.rank = rank
    config.allTileNames = ['%d_%d' % (i // 4, i % 4) for i in range(world * nT)]    # the same list on every rank
    parDict = config.parDict
    ops = device.getOps(args.precision)
    F, S = 1, len(parDict['mapFilters'])
    E = 8 if args.precision == 'f64' else 4
    filtersList = pipelines._orderedFilters(parDict)

    # ---- device-resident arm ---------------------------------------------------------------------------
    batches = pipelines.stageTiles(config, args.tiles_per_batch, args.precision)
    templates = pipelines.FilterSetTemplates(ops, filtersList, batches[0].mapDicts[0])
    config._templates = templates
    for b in batches:
        b.upload()
    torch.cuda.synchronize()

    # one pipeline for the whole run: the host half of step i (catalog assembly) overlaps the device half of step i + 1,
    # as it does in a run over many tiles; flush() inside the timed region completes the last step
    def countObjects(res):
        return sum(len(v['catalog']) for v in res[0].values())

    def deviceSteps(n):
        pipe = pipelines.BatchPipeline(ops, filtersList, templates, parDict, undoPixelWindow=True, depth=args.pipeline_depth)
        nObj = 0
        for _ in range(n):
            for b in batches:
                out = pipe.submit(b, upload=False)
                if out is not None:
                    nObj += countObjects(out)
        nObj += sum(countObjects(out) for out in pipe.flush())
        return nObj // n

    nObj = deviceSteps(args.warmup)
    sampler = ClockSampler(localRank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = _lib.launchCount()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nObj = deviceSteps(args.steps)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = (_lib.launchCount() - launches0) // args.steps
    # per-kernel CUDA-event times: the same steps once more, one at a time (no overlap between consecutive steps, which
    # would stretch the individual kernel times), with the library's event timers on
    _lib.profEnable(True)
    for _ in range(args.steps):
        deviceSteps(1)
    prof = _lib.profReport()
    _lib.profEnable(False)
    msStep = maxOverRanks(e0.elapsed_time(e1) / args.steps)
    mpixJob = world * nT * H * W * S / 1e6
    value = mpixJob / (msStep * 1e-3)

    # ---- end-to-end arm: pinned host buffers -> host catalog through the public driver ------------------------
    for b in batches:
        b.dev = None
    for _ in range(2):
        cat = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=e2eBatch, precision=args.precision, comm=comm)
    barrier()
    device.TRAFFIC['h2d'] = device.TRAFFIC['d2h'] = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cat = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=e2eBatch, precision=args.precision, comm=comm)
    torch.cuda.synchronize()
    tE2E = maxOverRanks((time.perf_counter() - t0) / args.steps)
    barrier()
    e2e = {'value': mpixJob / tE2E, 'unit': UNIT, 'h2d_bytes_per_step': device.TRAFFIC['h2d'] // args.steps,
           'd2h_bytes_per_step': device.TRAFFIC['d2h'] // args.steps, 'ms_per_step': tE2E * 1e3, 'catalog_rows': len(cat)}

    # the same arm with the tiles held as float32 on the host, the format survey maps have on disk (FITS BITPIX -32): staged and
    # uploaded as float32, widened to float64 on the device, everything after that unchanged.  Reported beside `e2e`, which stays
    # the float64-host number.
    e2eF32 = None
    if args.precision == 'f64' and not args.no_f32_host:
        tileMaps32 = {name: [dict(m, data=np.asarray(m['data'], dtype=np.float32), weights=np.asarray(m['weights'], dtype=np.float32))
                             for m in md] for name, md in tileMaps.items()}
        config32 = pipelines.SimpleConfig(parDict, tileMaps32, rank=0, size=1)
        config32.rank = rank
        config32.allTileNames = config.allTileNames
        config32._templates = templates
        for _ in range(2):
            cat32 = pipelines.filterMapsAndMakeCatalogs(config32, verbose=False, tilesPerBatch=e2eBatch, precision=args.precision, comm=comm)
        barrier()
        device.TRAFFIC['h2d'] = device.TRAFFIC['d2h'] = 0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cat32 = pipelines.filterMapsAndMakeCatalogs(config32, verbose=False, tilesPerBatch=e2eBatch, precision=args.precision, comm=comm)
        torch.cuda.synchronize()
        t32 = maxOverRanks((time.perf_counter() - t0) / args.steps)
        barrier()
        e2eF32 = {'value': mpixJob / t32, 'unit': UNIT, 'h2d_bytes_per_step': device.TRAFFIC['h2d'] // args.steps,
                  'd2h_bytes_per_step': device.TRAFFIC['d2h'] // args.steps, 'ms_per_step': t32 * 1e3, 'catalog_rows': len(cat32),
                  'note': 'host tiles float32 (as on disk), widened to float64 on the device; compute unchanged'}

    if comm is not None:
        torch.distributed.destroy_process_group()
    if rank != 0:
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------
    peaks = {}
    peakPath = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peakPath):
        peaks = json.load(open(peakPath))
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peakSrc = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
    algo = algorithmicBytes(nT, F, S, H, W, E, rings=parDict['removeRings'])
    table = []
    totalKernelMs = sum(ms for _, ms in prof.values()) / args.steps
    for name, (cnt, ms) in prof.items():
        msK = ms / args.steps
        row = {'kernel': name, 'launches_per_step': cnt / args.steps, 'ms_per_step': msK, 'share': msK / totalKernelMs}
        if name in algo:
            row['algorithmic_bytes_per_step'] = algo[name]
            row['achieved_gbs'] = algo[name] / (msK * 1e-3) / 1e9
            row['frac'] = row['achieved_gbs'] / peak
        table.append(row)
    table.sort(key=lambda r: -r['ms_per_step'])
    top = next((r for r in table if 'achieved_gbs' in r), None)
    traffic, trafficSrc = None, None
    trafficPath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if top is not None and os.path.exists(trafficPath):
        tj = json.load(open(trafficPath))
        ent = tj.get(args.precision, {}).get(top['kernel'])
        if ent is not None:
            traffic, trafficSrc = ent['dram_bytes_per_launch'], ent['source']
    roofline = None
    if top is not None:
        lp = top['launches_per_step']
        roofline = {'bound': 'hbm', 'kernel': top['kernel'], 'achieved': top['achieved_gbs'], 'peak': peak, 'unit': 'GB/s',
                    'frac': top['frac'], 'traffic': traffic, 'traffic_source': trafficSrc, 'peak_source': peakSrc, 'timed_in': 'same run, %d further steps executed one at a time with per-kernel CUDA events' % args.steps, 'share_of_step_kernel_time': top['share'],
                    'bytes_per_launch': top['algorithmic_bytes_per_step'] / lp, 'ms_per_launch': top['ms_per_step'] / lp,
                    'launches_per_step': lp}
    bpp = endToEndBytesPerPixel(F, S, E, rings=parDict['removeRings'])
    stepRoofline = {'algorithmic_bytes_per_pixel_scale': bpp, 'achieved_gbs': bpp * nT * H * W * S / (msStep * 1e-3) / 1e9}
    stepRoofline['frac'] = stepRoofline['achieved_gbs'] / peak
    stepRoofline['kernel_ms_per_step'] = totalKernelMs
    if args.kernel_table:
        with open(args.kernel_table, 'w') as fh:
            json.dump({'precision': args.precision, 'tiles': nT, 'shape': SHAPE, 'F': F, 'S': S, 'ms_per_step': msStep,
                       'kernel_ms_per_step': totalKernelMs, 'peak_gbs': peak, 'kernels': table, 'step': stepRoofline}, fh, indent=1)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu, _ = cpuBaseline(cores=cores, tiles=min(cores, TILES_PER_GPU), steps=1, warmup=0)

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': msStep, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': args.precision, 'data': 'synthetic',
            'config': {'workload': 'configs[1]: BeamMatchedFilter + S/N>4 detection on a 4096x4096 148 GHz map in 16 tiles of 1024x1024 per GPU',
                       'tiles_per_gpu': nT, 'tile': list(SHAPE), 'freqs': F, 'scales': S, 'tiles_per_batch': args.tiles_per_batch, 'e2e_tiles_per_batch': e2eBatch,
                       'l2': 'inputs and intermediates of one step (> 1 GB) exceed the 126 MB L2; no explicit flush',
                       'objects_per_step': nObj},
            'clocks': clocks, 'e2e': e2e, 'e2e_f32_host': e2eF32, 'gpu_launches': int(launches), 'roofline': roofline, 'step_roofline': stepRoofline,
            'cpu_baseline': cpu}
    print(json.dumps(line))


if __name__ == '__main__':
    main()
