#!/usr/bin/env python
"""
bench.py -- throughput of the per-tile filter -> noise/S-N -> detect -> photometry hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload dr5|ptsrc]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (default, BASELINE.json configs[3]): an ACT-DR5-scale cluster search -- 280 tiles of 840 x 1440 pixels at 149.6 + 97.8
GHz, the 16 Arnaud_M*_z* matched-filter scales of the reference's examples/ACT-DR5-clusters/DR5ClusterSearch.yml with photFilter
Arnaud_M2e14_z0p4, S/N > 4 detection, ring removal, spline photometry (bench_workloads.dr5Config).  Tiles are sharded over the
GPUs: 35 tiles per GPU (280 at 8 GPUs), weak scaling.  One step = one pass of the whole path over the rank's 35 tiles x 16
scales.  Tiles are independent, so there is no collective on the data path; the end-to-end step ends with the catalog /
noise-table gathers.  `--workload ptsrc` runs configs[1] (16 tiles of 1024 x 1024, BeamMatchedFilter, last round's bench line).

Metric: filtered+detected Mpix/s per filter scale = tiles * H * W * S / step time: every (pixel, scale) pair that was filtered
and searched counts once (whole job, all ranks).
  value : inputs resident in HBM when the timed region starts (CUDA events, max over ranks)
  e2e   : pipelines.filterMapsAndMakeCatalogs from pinned HOST buffers to the host catalog, copies inside the timed region
  e2e_from_pageable : the same with the numpy -> pinned staging copy of every tile inside the timed region as well
  roofline     : the kernel with the largest share of the step, algorithmic bytes / its CUDA-event time vs measured HBM peak
  cpu_baseline : the numpy/scipy oracle (the reference's arithmetic) on a bounded sample of the same tiles, on the host cores

--impl reference times that CPU path alone on all host cores: one tile per worker process per step and, for the 16-scale
workload, the photFilter plus one further scale per step (the reference shares nothing between scales, so its cost per scale
is the same for every scale; 16 scales would take a minute per step).
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

WORKLOADS = {
    'dr5': {'shape': (840, 1440), 'tiles': 35, 'freqs': 2, 'scales': 16, 'cpu_scales': 2, 'tiles_per_batch': 8,
            'name': "configs[3]: ACT-DR5-scale cluster search, 840x1440 tiles at 149.6+97.8 GHz, 16 Arnaud_M*_z* filter scales "
                    "(ArnaudModelMatchedFilter, photFilter Arnaud_M2e14_z0p4), S/N>4 detection; 35 tiles per GPU (280 tiles at 8 GPUs)"},
    'ptsrc': {'shape': (1024, 1024), 'tiles': 16, 'freqs': 1, 'scales': 1, 'cpu_scales': 1, 'tiles_per_batch': 16,
              'name': "configs[1]: BeamMatchedFilter + S/N>4 detection on a 4096x4096 148 GHz map, 16 tiles of 1024x1024 per GPU"},
}


def makeWorkload(kind, numTiles, firstIndex):
    import bench_workloads
    if kind == 'dr5':
        return bench_workloads.dr5Config(numTiles=numTiles, firstIndex=firstIndex)
    return bench_workloads.pointSourceConfig(numTiles=numTiles, shape=WORKLOADS[kind]['shape'], firstIndex=firstIndex)


# ------------------------------------------------------------------------------------------------------
# CPU arm (the oracle): one tile per worker process
def _cpuTile(job):
    """Worker: regenerates tile `index` (same seed as the GPU arm) and runs the reference arithmetic on it for the filters
    `scaleIdx` of the ordered list, in the reference's loop order (pipelines.py:179-286: photFilter first)."""
    kind, index, scaleIdx = job
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from oracle import filters_ref, photometry_ref
    from nemo_b200 import pipelines
    parDict, tileMaps, _ = makeWorkload(kind, 1, index)
    parDict = dict(pipelines.SimpleConfig.DEFAULTS, **parDict)
    name, mapDicts = next(iter(tileMaps.items()))
    ordered = pipelines._orderedFilters(parDict)
    t0 = time.perf_counter()
    phot, nObj = None, 0
    for s in scaleIdx:
        f = ordered[s]
        filtered = filters_ref.filterMaps(mapDicts, {'label': f['label'], 'class': f['class'], 'params': dict(f['params'])}, name)
        if f['label'] == parDict.get('photFilter'):
            phot = {'SNMap': filtered['SNMap'], 'data': filtered['data']}
        cat = photometry_ref.findObjects(filtered, threshold=parDict['thresholdSigma'], minObjPix=parDict['minObjPix'],
                                         rejectBorder=parDict['rejectBorder'], findCenterOfMass=parDict['findCenterOfMass'],
                                         removeRings=parDict['removeRings'], ringThresholdSigma=parDict['ringThresholdSigma'],
                                         objIdent=parDict['objIdent'], longNames=parDict['longNames'],
                                         useInterpolator=parDict['useInterpolator'])
        photometry_ref.measureFluxes(cat, filtered, None, photFilteredMapDict=phot, useInterpolator=parDict['useInterpolator'])
        nObj += len(cat)
    return time.perf_counter() - t0, nObj


def _cpuPool(cores):
    import multiprocessing as mp
    os.environ['OMP_NUM_THREADS'] = '1'
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    os.environ['MKL_NUM_THREADS'] = '1'
    return mp.get_context('spawn').Pool(cores)


def cpuBaseline(kind, cores=None, steps=1, warmup=0):
    """Times the oracle over one tile per worker process per step on `cores` processes (the reference's MPI-over-tiles split,
    startUp.py:390-404).  Tile generation happens inside the workers but outside their timers; the step time is the slowest
    worker's compute time (what an MPI job's wall clock would be)."""
    wl = WORKLOADS[kind]
    cores = cores or os.cpu_count() or 1
    tiles = cores
    S, nS = wl['scales'], wl['cpu_scales']
    pool = _cpuPool(cores)
    try:
        times = []
        nObj = 0
        for it in range(warmup + steps):
            # the photFilter (index 0 of the ordered list) and the next of the other scales in rotation
            scaleIdx = [0] + [1 + (it * (nS - 1) + k) % (S - 1) for k in range(nS - 1)] if S > 1 else [0]
            res = pool.map(_cpuTile, [(kind, i % wl['tiles'], scaleIdx) for i in range(tiles)], chunksize=1)
            if it >= warmup:
                times.append(max(dt for dt, _ in res))
                nObj = sum(n for _, n in res)
    finally:
        pool.close()
        pool.join()
    tStep = sum(times) / len(times)
    H, W = wl['shape']
    mpix = tiles * H * W * nS / 1e6
    sample = ("%d of the workload's %dx%d tiles per step, one per worker process, %d of its %d filter scales per tile (photFilter + "
              "%d in rotation; the reference shares nothing between scales), %d step(s); numpy/scipy oracle (np.fft for pixell, "
              "vectorised np.linalg.inv), %.2f s per step, %d objects" % (tiles, H, W, nS, S, nS - 1, steps, tStep, nObj))
    return {'value': mpix / tStep, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample}, tStep


# ------------------------------------------------------------------------------------------------------
def algorithmicBytes(nT, F, S, H, W, E, rings=True, yc=True, templateSpectraCached=True, separableWindow=False, edgeFilterRuns=True):
    """Algorithmic HBM bytes per step of every kernel of the batched Fourier path: own inputs read once + own outputs
    written once (SURVEY.md 8d / DESIGN.md), E = bytes per real (8 for f64, 4 for f32), complex = 2E.  Keyed by the names the
    library's event timers report; kernels that serve several passes are charged the sum of those passes.
    templateSpectraCached: the transforms of the signal templates are made once per tile geometry (pipelines.templateSpectraOf),
    not per batch, so no FFT pass of a step is charged for them.
    separableWindow: the pixel-window undo runs as a column pass on pairs of real columns and one row kernel (real rows in, real
    rows out, the zero pattern of the old data read once) instead of forward rows / column window / inverse rows.
    edgeFilterRuns=False: no image had a zero for K7's edge filter to find (nb_apply_filter's zeroFlags), so its row pass did not run
    and its column pass reads no row flags."""
    N, Nc = H * W, H * (W // 2 + 1)
    P = F * (F + 1) // 2
    nOut = nT * S
    nPass = 2 if rings else 1
    nTpl = S * F
    nTplFFT = 0 if templateSpectraCached else nTpl          # template images transformed inside the step
    rowFwd = N * E + Nc * 2 * E               # per image: real rows in, half spectrum out
    rowInv = Nc * 2 * E + N * E
    col = 2 * Nc * 2 * E                      # half spectrum in and out
    fwdImgs = nT * F + nTplFFT + (2 * nTplFFT if yc else 0) + (0 if separableWindow else nOut)   # maps, unit templates, (windowed) norm templates, window undo
    plainColImgs = nT * F + nTplFFT + (nTplFFT if yc else 0)
    windowColImgs = nOut + (nTplFFT if yc else 0)
    applyBytes = 2 * nOut * (F * Nc * E + Nc * 2 * E) + (nT * F + nTpl) * Nc * 2 * E     # filters in, product out; spectra once
    invBytes = 2 * nOut * rowInv + nOut * N + (0 if separableWindow else nOut * rowInv + nOut * N * E) + (nTplFFT * rowInv if yc else 0)   # + mask, + zero reference
    b = {}
    for rows in ('rfft_rows_kernel', 'bs_rows_fwd_kernel', 'sfft_rows_fwd_kernel'):
        b[rows] = fwdImgs * rowFwd
    for rows in ('irfft_rows_kernel', 'bs_rows_inv_kernel', 'sfft_rows_inv_kernel'):
        b[rows] = invBytes
    if separableWindow:
        windowCol = 2 * H * (W // 2) * 2 * E          # pairs of real columns in and out
        b['sfft_rows_window_kernel'] = nOut * 3 * N * E      # filtered rows in, old data (zero pattern) in, data out
    else:
        windowCol = col
    b['fft_cols_kernel'] = plainColImgs * col
    b['fft_cols_window_kernel'] = windowColImgs * windowCol
    b['fft_cols_apply_kernel'] = applyBytes
    b['bs_cols_kernel'] = b['sfft_cols_kernel'] = plainColImgs * col + windowColImgs * windowCol + applyBytes
    b['xpower_smooth_x_kernel'] = b['xpower_smooth_x_rb_kernel'] = nT * (F * Nc * 2 * E + P * N * 8)
    b['smooth_y_f64_kernel'] = b['smooth_y_rb_kernel'] = nT * P * N * 8 * 2
    b['build_filter_kernel'] = nT * P * N * 8 + nTpl * Nc * 2 * E + nOut * F * Nc * E
    b['template_argmax_partial_kernel'] = nOut * N * E
    b['plane_values_partial_kernel'] = (nOut * Nc * E + S * Nc * 2 * E) * max(F - 1, 0)    # the last frequency's plane is the remainder of the total
    for k6 in ('cell_rms_onchip_kernel', 'cell_rms_hybrid_kernel', 'cell_rms_stream_kernel', 'cell_rms_kernel', 'paint_rms_kernel'):
        b[k6] = nOut * N * E + nT * N * E      # every image's map; the median weights of its tile
    b['zone_moments_kernel'] = nOut * N * E + nT * N * E     # one-bin K6: map once + the tile's median weights; then the map once more
    b['cell_clip_kernel'] = nOut * N * E
    if edgeFilterRuns:
        b['edge_rows_kernel'] = nOut * (N * E + 2 * N)
    nByteIn = 3 if edgeFilterRuns else 2      # point-source mask, survey mask, (row flags of the edge filter)
    b['edge_cols_final_kernel'] = nOut * (2 * N * E + nByteIn * N + 3 * N * E + N)
    b['edge_cols_final_paint_kernel'] = nOut * (2 * N * E + nByteIn * N + 2 * N * E + N)   # map, median weights, masks in; data, S/N, mask out (the RMS map is not written unless asked for)
    b['detect_scan_kernel'] = b['detect_scan_multi_kernel'] = nOut * (N * E)         # S/N read once for both thresholds; the rest is sparse
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


L2_NOTE = 'GPU arm: inputs and intermediates of one batch (> 1 GB) exceed the 126 MB L2, no explicit flush; CPU arm: not applicable'


def workloadConfig(kind, args=None):
    """The `config` object both arms print (same `workload` string, so the driver can tell they ran the same thing)."""
    wl = WORKLOADS[kind]
    return {'workload': wl['name'], 'tile': list(wl['shape']), 'freqs': wl['freqs'], 'scales': wl['scales'],
            'tiles_per_gpu': args.tiles if args is not None and args.tiles else wl['tiles']}


# ------------------------------------------------------------------------------------------------------
def runReference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    base, tStep = cpuBaseline(args.workload, cores=cores, steps=args.steps, warmup=args.warmup)
    cfg = workloadConfig(args.workload, args)
    cfg['l2'] = L2_NOTE
    line = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': tStep * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': cfg, 'cpu_baseline': base,
            'run': {'cpu_tiles_per_step': cores, 'cpu_scales_per_step': WORKLOADS[args.workload]['cpu_scales']},
            'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='dr5', choices=sorted(WORKLOADS))
    ap.add_argument('--precision', default='f64', choices=['f64', 'f32'])
    ap.add_argument('--tiles', type=int, default=None, help='tiles per GPU (default: the workload\'s)')
    ap.add_argument('--tiles-per-batch', type=int, default=None)
    ap.add_argument('--e2e-tiles-per-batch', default=None, help='batch sizes of the end-to-end arm (uploads overlap compute): one size, or a list that is mirrored at the end (2,4 -> 2,4,4,...,4,2)')
    ap.add_argument('--pipeline-depth', type=int, default=2, help='batches in flight on the device behind the one whose catalogs the host assembles (device-resident arm)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra-e2e', action='store_true', help='skip the extra end-to-end arms (pageable host tiles; float32 host tiles)')
    ap.add_argument('--kernel-table', default=None, help='write the per-kernel roofline table to this file')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    localRank = int(os.environ.get('LOCAL_RANK', '0'))
    wl = WORKLOADS[args.workload]
    if args.tiles is None:
        args.tiles = wl['tiles']
    if args.impl == 'reference':
        runReference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    tilesPerBatch = args.tiles_per_batch or wl['tiles_per_batch']
    if args.e2e_tiles_per_batch is None:
        e2eBatch = [2, 4, tilesPerBatch] if args.workload == 'dr5' else 4
    else:
        e2eBatch = [int(v) for v in str(args.e2e_tiles_per_batch).split(',')]
        e2eBatch = e2eBatch[0] if len(e2eBatch) == 1 else e2eBatch

    import numpy as np
    import torch
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

    H, W = wl['shape']
    nT = args.tiles
    parDict, tileMaps, truth = makeWorkload(args.workload, nT, rank * nT)
    config = pipelines.SimpleConfig(parDict, tileMaps, rank=0, size=1)     # every rank owns all of its own tiles
    config.rank = rank
    import bench_workloads
    if args.workload == 'dr5':
        config.allTileNames = [bench_workloads.dr5TileName(i) for i in range(world * nT)]    # the same list on every rank
    else:
        config.allTileNames = ['%d_%d' % (i // 4, i % 4) for i in range(world * nT)]
    parDict = config.parDict
    ops = device.getOps(args.precision)
    F, S = wl['freqs'], len(parDict['mapFilters'])
    E = 8 if args.precision == 'f64' else 4
    filtersList = pipelines._orderedFilters(parDict)
    outputUnits = filtersList[0]['params']['outputUnits']

    # ---- device-resident arm ---------------------------------------------------------------------------
    batches = pipelines.stageTiles(config, tilesPerBatch, args.precision)
    templatesList = [pipelines.templatesFor(config, ops, filtersList, b) for b in batches]
    for b in batches:
        b.upload()
    torch.cuda.synchronize()

    # one pipeline for the whole run: the host half of step i (catalog assembly) overlaps the device half of step i + 1,
    # as it does in a run over many tiles; flush() inside the timed region completes the last step
    def countObjects(res):
        return sum(len(v['catalog']) for v in res[0].values())

    def deviceSteps(n, serialize=False):
        pipe = pipelines.BatchPipeline(ops, filtersList, templatesList[0], parDict, undoPixelWindow=True, depth=args.pipeline_depth,
                                       serialize=serialize)
        nObj = 0
        for _ in range(n):
            for b, tpl in zip(batches, templatesList):
                out = pipe.submit(b, upload=False, templates=tpl)
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
    # per-kernel CUDA-event times: the same steps once more with the library's event timers on, every batch waited for before
    # the next is queued (kernels of batches running side by side on the pipeline's streams would stretch each other's times)
    profSteps = min(args.steps, 5)
    _lib.profEnable(True)
    for _ in range(profSteps):
        deviceSteps(1, serialize=True)
    prof = _lib.profReport()
    _lib.profEnable(False)
    msStep = maxOverRanks(e0.elapsed_time(e1) / args.steps)
    mpixJob = world * nT * H * W * S / 1e6
    value = mpixJob / (msStep * 1e-3)

    # ---- end-to-end arm: pinned host buffers -> host catalog through the public driver ------------------------
    for b in batches:
        b.dev = None
    # the device-resident arm's CUDA graphs (and the memory pools their outputs live in) are not needed any more
    del batches
    ops.__dict__.pop('_batchGraphs', None)
    import gc
    gc.collect()
    torch.cuda.empty_cache()

    def e2eArm(cfg, restage=False, steps=args.steps):
        for _ in range(2):
            cat = pipelines.filterMapsAndMakeCatalogs(cfg, verbose=False, tilesPerBatch=e2eBatch, precision=args.precision, comm=comm,
                                                      restage=restage, restageKeys=('data', 'weights'))
        barrier()
        device.TRAFFIC['h2d'] = device.TRAFFIC['d2h'] = 0
        t0 = time.perf_counter()
        for _ in range(steps):
            cat = pipelines.filterMapsAndMakeCatalogs(cfg, verbose=False, tilesPerBatch=e2eBatch, precision=args.precision, comm=comm,
                                                      restage=restage, restageKeys=('data', 'weights'))
        torch.cuda.synchronize()
        t = maxOverRanks((time.perf_counter() - t0) / steps)
        barrier()
        return {'value': mpixJob / t, 'unit': UNIT, 'h2d_bytes_per_step': device.TRAFFIC['h2d'] // steps,
                'd2h_bytes_per_step': device.TRAFFIC['d2h'] // steps, 'ms_per_step': t * 1e3, 'catalog_rows': len(cat)}

    e2e = e2eArm(config)
    e2ePageable = e2eF32 = None
    if not args.no_extra_e2e:
        # the numpy -> pinned staging copy of every tile inside the timed region too (restage: each batch's maps and weights are
        # copied into its pinned buffers right before it is submitted, under the device work of the batch before it)
        e2ePageable = e2eArm(config, restage=True, steps=max(2, args.steps // 2))
        e2ePageable['note'] = 'tiles start as pageable numpy arrays: staging copy into pinned buffers inside the timed region'
        if args.precision == 'f64':
            # tiles held as float32 on the host, the format survey maps have on disk (FITS BITPIX -32): staged and uploaded as
            # float32, widened to float64 on the device, everything after that unchanged
            tileMaps32 = {name: [dict(m, data=np.asarray(m['data'], dtype=np.float32), weights=np.asarray(m['weights'], dtype=np.float32))
                                 for m in md] for name, md in tileMaps.items()}
            config32 = pipelines.SimpleConfig(parDict, tileMaps32, rank=0, size=1)
            config32.rank = rank
            config32.allTileNames = config.allTileNames
            config32._templates = config._templates
            e2eF32 = e2eArm(config32, steps=max(2, args.steps // 2))
            e2eF32['note'] = 'host tiles float32 (as on disk), widened to float64 on the device; compute unchanged'

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
    algo = algorithmicBytes(nT, F, S, H, W, E, rings=parDict['removeRings'], yc=(outputUnits == 'yc'),
                            separableWindow='sfft_rows_window_kernel' in prof, edgeFilterRuns=prof.get('edge_rows_kernel', (0, 0.0))[1] / max(prof.get('edge_rows_kernel', (1, 0.0))[0], 1) > 0.1)      # its CTAs return at once when no image has a zero
    table = []
    totalKernelMs = sum(ms for _, ms in prof.values()) / profSteps
    for name, (cnt, ms) in prof.items():
        msK = ms / profSteps
        row = {'kernel': name, 'launches_per_step': cnt / profSteps, 'ms_per_step': msK, 'share': msK / totalKernelMs}
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
        ent = tj.get(args.workload, {}).get(args.precision, {}).get(top['kernel'])
        if ent is not None:
            # ncu captured one launch of some batch size: scale to the bytes one average launch of this run moves
            traffic = ent['dram_bytes_per_launch']
            if ent.get('algorithmic_bytes_per_launch'):
                traffic *= (top['algorithmic_bytes_per_step'] / top['launches_per_step']) / ent['algorithmic_bytes_per_launch']
            trafficSrc = ent['source']
    roofline = None
    if top is not None:
        lp = top['launches_per_step']
        roofline = {'bound': 'hbm', 'kernel': top['kernel'], 'achieved': top['achieved_gbs'], 'peak': peak, 'unit': 'GB/s',
                    'frac': top['frac'], 'traffic': traffic, 'traffic_source': trafficSrc, 'peak_source': peakSrc,
                    'timed_in': 'same run, %d further steps with per-kernel CUDA events, one batch at a time' % profSteps,
                    'share_of_step_kernel_time': top['share'],
                    'bytes_per_launch': top['algorithmic_bytes_per_step'] / lp, 'ms_per_launch': top['ms_per_step'] / lp,
                    'launches_per_step': lp}
    bpp = endToEndBytesPerPixel(F, S, E, rings=parDict['removeRings'])
    stepRoofline = {'algorithmic_bytes_per_pixel_scale': bpp, 'achieved_gbs': bpp * nT * H * W * S / (msStep * 1e-3) / 1e9}
    stepRoofline['frac'] = stepRoofline['achieved_gbs'] / peak
    stepRoofline['kernel_ms_per_step'] = totalKernelMs
    if args.kernel_table:
        with open(args.kernel_table, 'w') as fh:
            json.dump({'workload': args.workload, 'precision': args.precision, 'tiles': nT, 'shape': [H, W], 'F': F, 'S': S, 'ms_per_step': msStep,
                       'kernel_ms_per_step': totalKernelMs, 'peak_gbs': peak, 'kernels': table, 'step': stepRoofline}, fh, indent=1)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpuBaseline(args.workload, cores=os.cpu_count() or 1, steps=1, warmup=0)

    cfg = workloadConfig(args.workload, args)
    cfg['l2'] = L2_NOTE
    run = {'tiles_per_batch': tilesPerBatch, 'e2e_tiles_per_batch': e2eBatch, 'objects_per_step': nObj,
           'hbm_peak_gb': torch.cuda.max_memory_allocated() / 1e9}
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': msStep, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': args.precision, 'data': 'synthetic', 'config': cfg, 'run': run,
            'clocks': clocks, 'e2e': e2e, 'e2e_from_pageable': e2ePageable, 'e2e_f32_host': e2eF32, 'gpu_launches': int(launches),
            'roofline': roofline, 'step_roofline': stepRoofline, 'cpu_baseline': cpu}
    print(json.dumps(line))


if __name__ == '__main__':
    main()
