"""
Batched tile x filter-scale driver: the body of the reference's ``pipelines._filterMapsAndMakeCatalogs``
(nemo/pipelines.py:112-354) with its ``for tileName: for f in filtersList:`` loops turned into batched GPU
launches, followed by the same cross-scale merge (``catalogs.makeOptimalCatalog``).  Tiles are sharded over
ranks (one process per GPU); the only collectives are the catalog all-gather (pipelines.py:322) and the
noise-table gather (pipelines.py:405) -- see ``nemo_b200.distributed``.

Work items are independent (tile, scale) pairs; image index o = tile * S + scale.  All S scales of a tile stay on
the same GPU so that the input FFTs and the smoothed noise covariance are shared between them and the
``photFilter`` maps (pipelines.py:238-241) are local.
"""

import os
import numpy as np

from . import catalogs
from . import device
from . import engine
from . import maps
from . import photometry
from . import signals
from . import wcs as _wcs
from . import filters as _filters

FLUX_KEYS = {'yc': ['deltaT_c', 'err_deltaT_c', 'y_c', 'err_y_c'], 'uK': ['deltaT_c', 'err_deltaT_c', 'fluxJy', 'err_fluxJy']}


class SimpleConfig(object):
    """The slice of the reference's ``startUp.NemoConfig`` the hot path reads: ``parDict`` (thresholdSigma, minObjPix,
    findCenterOfMass, removeRings, ringThresholdSigma, rejectBorder, objIdent, longNames, useInterpolator,
    measureShapes, photFilter, mapFilters, catalogCuts), ``tileNames`` and the per-tile map dictionaries."""

    DEFAULTS = {'removeRings': True, 'ringThresholdSigma': 3, 'rejectBorder': 0, 'measureShapes': False, 'objIdent': 'ACT-CL',
                'longNames': False, 'useInterpolator': True, 'findCenterOfMass': True, 'photFilter': None, 'catalogCuts': [],
                'undoPixelWindow': True}

    def __init__(self, parDict, tileMaps, rank=0, size=1):
        self.parDict = dict(self.DEFAULTS)
        self.parDict.update(parDict)
        self.tileMaps = tileMaps                  # {tileName: [mapDict per frequency]}
        self.allTileNames = list(tileMaps.keys())
        self.rank, self.size = rank, size
        # tiles round-robin over ranks (the reference leaves rank 0 idle, startUp.py:390-404; here every rank works)
        self.tileNames = self.allTileNames[rank::size] if size > 1 else list(self.allTileNames)


# ------------------------------------------------------------------------------------------------------
def tileGeometry(mapDicts):
    """What the filters, templates and noise grid of a tile depend on besides its pixels: (H, W, degPerPixX, degPerPixY,
    pixel size, centre pixel x, y).  The reference rebuilds radiansMap, the pixel scales and the signal templates for every tile
    (filters.py:214-239, 617, 642-646); degPerPixX goes with cos(dec), so CAR tiles at different declinations have different
    filters.  Tiles with equal geometry share templates; a batch never mixes geometries."""
    wcs = mapDicts[0]['wcs']
    H, W = mapDicts[0]['data'].shape
    x0, y0 = int(W / 2), int(H / 2)
    ra0, dec0 = wcs.pix2wcs(x0, y0)
    ra1, dec1 = wcs.pix2wcs(x0 + 1, y0 + 1)
    cRADeg, cDecDeg = wcs.getCentreWCSCoords()
    cx, cy = wcs.wcs2pix(cRADeg, cDecDeg)
    # rounded so that tiles of one declination row, whose pixel scales and centres agree to the last few bits only (their RA
    # offsets round differently: the separations come out 1e-12 apart), land in one group: 1e-9 relative on the scales, 1e-6
    # pixel on the centre
    sig = lambda v: float('%.8e' % v)
    return (int(H), int(W), sig(_wcs.calcAngSepDeg(ra0, dec0, ra1, dec0)), sig(_wcs.calcAngSepDeg(ra0, dec0, ra0, dec1)),
            sig(wcs.getPixelSizeDeg()), round(float(cx), 6), round(float(cy), 6))


class TileBatch(object):
    """Maps of nT tiles resident on the device: data / weights [nT, F, H, W], masks u8 [nT, H, W], median weights."""

    def __init__(self, ops, tileNames, tileMaps, pinned=True):
        torch = ops.torch
        self.ops = ops
        self.tileNames = list(tileNames)
        self.mapDicts = [tileMaps[t] for t in self.tileNames]
        nT = len(self.tileNames)
        F = len(self.mapDicts[0])
        H, W = self.mapDicts[0][0]['data'].shape
        self.nT, self.F, self.H, self.W = nT, F, H, W
        self.geometry = tileGeometry(self.mapDicts[0])
        for md in self.mapDicts[1:]:
            if tuple(md[0]['data'].shape) == (H, W) and tileGeometry(md) != self.geometry:
                raise Exception("tiles of one batch need the same geometry (shape, pixel scales at the tile centre): "
                                "stageTiles groups tiles by geometry")
        # maps that arrive as float32 (the on-disk format: FITS BITPIX -32) are staged and uploaded as float32 and widened on
        # the device -- exact, and half the host-to-device bytes; anything else is staged in the compute type
        def stageType(key):
            allF32 = all(np.asarray(m[key]).dtype == np.float32 for md in self.mapDicts for m in md)
            return torch.float32 if allF32 else ops.dtype
        data = torch.empty((nT, F, H, W), dtype=stageType('data'), pin_memory=pinned)
        weights = torch.empty((nT, F, H, W), dtype=stageType('weights'), pin_memory=pinned)
        survey = torch.empty((nT, H, W), dtype=torch.uint8, pin_memory=pinned)
        ps = torch.empty((nT, H, W), dtype=torch.uint8, pin_memory=pinned)
        for t, md in enumerate(self.mapDicts):
            if len(md) != F:
                raise Exception("all tiles of a batch need the same number of frequency maps")
            for f, m in enumerate(md):
                if tuple(m['data'].shape) != (H, W):
                    raise Exception("all tiles of a batch need the same shape")
                data[t, f] = torch.from_numpy(np.ascontiguousarray(m['data']))
                weights[t, f] = torch.from_numpy(np.ascontiguousarray(m['weights']))
            survey[t] = torch.from_numpy(_filters._maskToU8(md[0].get('surveyMask'), (H, W)))
            ps[t] = torch.from_numpy(_filters._maskToU8(md[0].get('pointSourceMask'), (H, W)))
        # pixel area (deg^2) of every row of every tile: input of the noise level / area tables (completeness.getRMSTab)
        from . import completeness
        rowArea = torch.empty((nT, H), dtype=torch.float64, pin_memory=pinned)
        for t, md in enumerate(self.mapDicts):
            rowArea[t] = torch.from_numpy(completeness.rowAreasDeg2((H, W), md[0]['wcs']))
        self.host = {'data': data, 'weights': weights, 'surveyMask': survey, 'psMask': ps, 'rowArea': rowArea}
        self.h2dBytes = sum(int(v.numel()) * v.element_size() for v in self.host.values())
        self.dev = None
        self.wcsList = [md[0]['wcs'] for md in self.mapDicts]
        self.flagMasks = []
        for md in self.mapDicts:
            fm = np.zeros((H, W), dtype=int)
            for i, m in enumerate(md):
                if m.get('flagMask') is not None:
                    fm = fm + np.asarray(m['flagMask']) * (i + 1)
            self.flagMasks.append(fm if fm.any() else None)      # an all-zero mask flags nothing: no per-object look-up

    def refresh(self, tileMaps, keys=('data',)):
        """Re-reads the given arrays of every tile into the pinned staging buffers (the maps changed, e.g. sources were
        injected); masks and everything else stay as staged."""
        import torch
        self.mapDicts = [tileMaps[t] for t in self.tileNames]
        for t, md in enumerate(self.mapDicts):
            for f, m in enumerate(md):
                for k in keys:
                    self.host[k][t, f] = torch.from_numpy(np.ascontiguousarray(m[k]))

    def inject(self, catalogs, override=None, GNFWParams='default', applyPixelWindow=True, minSNR=-99, validAreaSection=None):
        """The painter half of the source-injection loop on the device (maps.py:2093: mapDict['data'] = data + modelMap, with
        modelMap = maps.makeModelImage(...) per frequency map): the tiles' maps stay resident as uploaded, a model image of
        catalogs[tileName] is painted for every (tile, frequency) -- each map's own beam; cluster catalogs converted to delta T at
        the map's frequency -- and their sum goes to a second resident buffer that the filter step then reads.  Nothing but the
        catalogs crosses the host-device link.  catalogs: {tileName: catalog}; tiles without an entry get an unmodified copy."""
        from . import maps
        ops = self.ops
        if self.dev is None:
            self.upload()
        if getattr(self, 'devOriginal', None) is None:
            self.devOriginal = self.dev['data']
            self.devInjected = ops.empty(tuple(self.devOriginal.shape))
        nT, F, H, W = self.nT, self.F, self.H, self.W
        empty = {'RADeg': np.zeros(0), 'decDeg': np.zeros(0), 'deltaT_c': np.zeros(0)}
        wcsList, catList, beamList, freqList = [], [], [], []
        for t, name in enumerate(self.tileNames):
            for f in range(F):
                m = self.mapDicts[t][f]
                wcsList.append(self.wcsList[t])
                catList.append(catalogs.get(name, empty))
                beamList.append(m.get('beamFileName', m.get('beam')))
                freqList.append(m.get('obsFreqGHz'))
        _, counts = maps.paintModelImages(ops, (H, W), wcsList, catList, beamList, freqList, GNFWParams=GNFWParams,
                                          applyPixelWindow=applyPixelWindow, override=override, validAreaSection=validAreaSection,
                                          minSNR=minSNR, base=self.devOriginal.reshape(nT * F, H, W), out=self.devInjected)
        self.dev['data'] = self.devInjected
        return counts

    def upload(self):
        """Host (pinned) -> device copies of this batch's inputs; asynchronous on the current stream."""
        ops = self.ops
        self.dev = {k: device.countH2D(v).to(ops.device, non_blocking=True) for k, v in self.host.items()}
        widenStaged(ops, self.dev)
        self.dev['wmed'] = ops.median_weights(self.dev['weights'])
        return self.dev


def widenStaged(ops, dev, keys=('data', 'weights')):
    """float32-staged maps / weights of a device dictionary -> the compute type, in place in the dictionary."""
    for k in keys:
        if dev[k].dtype != ops.dtype:
            dev[k] = ops.widen(dev[k])
    return dev


# ------------------------------------------------------------------------------------------------------
class FilterSetTemplates(object):
    """Host-made signal templates of a list of Fourier filters for one tile geometry, uploaded once and reused for
    every tile with the same shape and pixel scale (template maps are inputs to the hot path)."""

    _serial = 0

    def __init__(self, ops, filtersList, mapDicts, tileName='PRIMARY'):
        FilterSetTemplates._serial += 1
        self.serial = FilterSetTemplates._serial      # never reused, unlike id(): CUDA-graph caches key on it
        self.filtersList = filtersList
        self.S = len(filtersList)
        unit, norm = [], []
        self.w = None
        self.outputUnits = None
        for f in filtersList:
            cls = f['class'] if not isinstance(f['class'], str) else getattr(_filters, f['class'])
            if not (issubclass(cls, _filters.MatchedFilter) or issubclass(cls, _filters.RealSpaceMatchedFilter)):
                raise Exception("FilterSetTemplates handles matched filters (Fourier-space or real-space) only")
            obj = cls.__new__(cls)
            obj.params = f['params']
            obj.unfilteredMapsDictList = mapDicts
            obj.wcs = mapDicts[0]['wcs']
            obj.shape = tuple(mapDicts[0]['data'].shape)
            obj.makeRadiansMap()
            w = engine.spectralWeights(mapDicts, f['params'])
            if self.w is not None and (not np.array_equal(w, self.w) or f['params']['outputUnits'] != self.outputUnits):
                raise Exception("filters of one batch need the same outputUnits and spectral weights")
            self.w, self.outputUnits = w, f['params']['outputUnits']
            unit.append([obj.makeSignalTemplateMap(m['beamFileName']) for m in mapDicts])
            amps = engine.normAmplitudes(mapDicts, f['params'])
            if amps is not None:
                # makeSignalTemplateMap(beam, amplitude=a) scales the 1-D profile by a before painting and convolving, all linear:
                # a times the unit template to rounding, at none of the painting cost
                norm.append([u * a for u, a in zip(unit[-1], amps)])
        self.geometry = tileGeometry(mapDicts)
        self.unit = ops.to_device(np.array(unit))
        self.norm = ops.to_device(np.array(norm)) if len(norm) > 0 else None
        self._spectra = None
        wcs = mapDicts[0]['wcs']
        cRADeg, cDecDeg = wcs.getCentreWCSCoords()
        self.cx, self.cy = wcs.wcs2pix(cRADeg, cDecDeg)


def templateSpectraOf(ops, templates):
    """(fsig, fsigN) of a FilterSetTemplates, made on first use and kept with it: the template transforms (K1 on 3 S F images and
    a pixel-window pass on S F of them) are the same for every tile of a geometry -- a sixth of the FFT work of a DR5 batch.
    Called outside CUDA-graph capture (the drivers call it when they fetch the templates)."""
    if templates._spectra is None:
        templates._spectra = engine.templateSpectra(ops, templates.unit, templates.norm, templates.outputUnits)
    return templates._spectra


def _sameNoiseSetup(filtersList):
    def key(f):
        p = f['params']
        return (repr(sorted(p['noiseParams'].items())), p.get('edgeTrimArcmin'), p.get('bckSub'), p.get('bckSubScaleArcmin'))
    return len(set(key(f) for f in filtersList)) == 1


def _filtersKey(filtersList):
    return repr([(f['label'], f['class'] if isinstance(f['class'], str) else f['class'].__name__,
                  repr(sorted(f['params'].items(), key=lambda kv: kv[0]))) for f in filtersList])


def _beamsKey(mapDicts):
    return tuple((m['obsFreqGHz'], m.get('units'), _filters._beamKey(m['beamFileName'])) for m in mapDicts)


def templatesFor(config, ops, filtersList, batch):
    """The FilterSetTemplates of a batch, built once per (tile geometry, beams, filter set, precision) and kept on `config`."""
    cache = getattr(config, '_templates', None)
    if not isinstance(cache, dict):
        cache = config._templates = {}
    key = (batch.geometry, _beamsKey(batch.mapDicts[0]), _filtersKey(filtersList), ops.precision)
    hit = cache.get(key)
    if hit is None:
        hit = cache[key] = FilterSetTemplates(ops, filtersList, batch.mapDicts[0], tileName=batch.tileNames[0])
    templateSpectraOf(ops, hit)
    return hit


def startTileBatch(ops, batch, filtersList, templates, parDict, undoPixelWindow=True, invertMap=False, cache=None,
                   useCachedRMSMap=False, rmsTables=False, wantRMS=True):
    """Launches, without waiting for any of it, the device work of one batch: filtering of every tile at every scale of
    `filtersList` (Fourier matched filters with a common noise set-up), noise / S-N maps, and the thresholding,
    labelling and moments of the detection step.  Returns the job to pass to finishTileBatch.

    cache: {'filtH', 'signalNorm', 'RMSMap'} kept from an earlier run of the same batch (the job of that run): the filters
    are then not rebuilt (the reference's useCachedFilters, filters.py:536, 691-696).  With useCachedRMSMap and a photFilter
    the S/N map is made from the photFilter's cached RMS map and the pixel window is undone afterwards, as the reference's
    source-injection path does (pipelines.py:216-232)."""
    if not _sameNoiseSetup(filtersList):
        raise Exception("filters of one batch need identical noiseParams / edgeTrimArcmin / bckSub settings")
    if getattr(batch, 'geometry', None) is not None and templates.geometry != batch.geometry:
        raise Exception("signal templates were made for another tile geometry %s than this batch's %s" % (templates.geometry, batch.geometry))
    dev = batch.dev if batch.dev is not None else batch.upload()
    nT, F, H, W, S = batch.nT, batch.F, batch.H, batch.W, len(filtersList)
    Wc = W // 2 + 1
    params0 = filtersList[0]['params']
    wcs0 = batch.wcsList[0]
    pixDeg = wcs0.getPixelSizeDeg()
    bck = None
    if 'bckSub' in params0.keys() and 'bckSubScaleArcmin' in params0.keys() and params0['bckSub'] == True:
        bck = maps.smoothScalesPix(wcs0, smoothScaleDeg=params0['bckSubScaleArcmin'] / 60.)
    method = params0['noiseParams']['method']
    if method != 'dataMap' or params0.get('noiseModelCatalog') is not None:
        raise Exception("nemo_b200 implements the 'dataMap' noise method without noiseModelCatalog")

    # K1: apodised input FFTs, shared by the S scales of a tile (filters.py:526-529) and by the noise estimate (:551)
    fmaps = ops.rfft2(dev['data'].reshape(nT * F, H, W), apodPix=20).reshape(nT, F, H, Wc)
    # K2, K3, K5: filters and their normalisation
    if cache is None:
        res = engine.buildFourierFilters(ops, fmaps, templates.unit, templates.norm, templates.w, templates.outputUnits,
                                         templates.cy, templates.cx, bckSubSigmas=bck, spectra=templates._spectra)
    else:
        res = {'filtH': cache['filtH'], 'signalNorm': cache['signalNorm'], 'fRel': cache.get('fRel')}
    # K4: apply, times the point-source mask
    # (zeroFlags: which images have a pixel for K7's edge filter to find -- a by-product of the last row pass; nearly always none)
    filtered, zeroFlags = engine.applyFourierFilters(ops, fmaps, res['filtH'], S, res['signalNorm'], dev['psMask'], bckSubSigmas=bck, W=batch.W,
                                                     zeroFlags=True)
    del fmaps
    trimPix = engine.trimSizePix(params0, pixDeg, realSpace=False)
    photScale = None
    if parDict.get('photFilter') is not None:
        for i, f in enumerate(filtersList):
            if f['label'] == parDict['photFilter']:
                photScale = i
    if cache is not None and useCachedRMSMap and photScale is not None:
        # K7 with the cached noise: masks and `data` as usual, S/N = data / RMS(photFilter) where RMS > 0, else data
        rmsPhot = cache['RMSMap'].reshape(nT, S, H, W)[:, photScale].contiguous()
        rms = rmsPhot.repeat_interleave(S, dim=0) if S > 1 else rmsPhot.clone()
        data, _, maskOut, rms = ops.snr_mask(filtered, rms, dev['psMask'], dev['surveyMask'], S, trimPix, 20, False, zeroFlags=zeroFlags)
        sn = ops.snr_from_rms(data, rmsPhot, S)
        table = edges = None
    else:
        # K6, K7
        grid = engine.noiseGrid((H, W), params0['noiseParams'], pixDeg)
        # (wantRMS=False: job['RMSMap'] is None -- nothing downstream reads the painted map unless noise tables, filter caches or
        # RMS files were asked for)
        data, sn, maskOut, rms, table, edges = engine.noiseSNAndMasks(ops, filtered, dev['wmed'], dev['psMask'], dev['surveyMask'],
                                                                      S, grid, trimPix, 20, False, wantRMS=wantRMS or rmsTables, zeroFlags=zeroFlags)
    # noise level versus area (completeness.getRMSTab) of every image, while its RMS map is on the device
    rmsTabJob = None
    if rmsTables and table is not None:
        from . import completeness
        rmsTabJob = completeness.startRMSTables(ops, rms, maskOut, dev['rowArea'], S, table)
    # K8
    if undoPixelWindow:
        engine.undoPixelWindow(ops, data)
    # K9, K10
    detectJob = photometry.startFindObjectsBatch(ops, sn, threshold=parDict['thresholdSigma'],
                                                 findCenterOfMass=parDict['findCenterOfMass'], removeRings=parDict['removeRings'],
                                                 ringThresholdSigma=parDict['ringThresholdSigma'], invertMap=invertMap,
                                                 measureShapes=parDict['measureShapes'])
    return {'ops': ops, 'batch': batch, 'filtersList': filtersList, 'parDict': parDict, 'mapUnits': templates.outputUnits,
            'invertMap': invertMap, 'detectJob': detectJob, 'data': data, 'SNMap': sn, 'surveyMask': maskOut, 'RMSMap': rms,
            'noiseTable': table, 'noiseEdges': edges, 'signalNorm': res['signalNorm'], 'fRel': res['fRel'], 'filtH': res['filtH'],
            'rmsTabJob': rmsTabJob}


def finishTileBatch(job, measureFluxes=True, keepMaps=False, collect=()):
    """Waits for a startTileBatch job and assembles its catalogs.  Returns (catalogDict, extras) where catalogDict maps
    '<label>#<tileName>' -> {'catalog': Table or []} like the reference's, and extras holds the noise tables (and
    the device maps when keepMaps).  collect: 'areaMask' and / or 'RMSMap' -- host copies, per tile, of the area mask (uint8) and
    the RMS map (float32) of the first filter of the list (the photFilter when there is one), for the selFn/ files the reference
    writes (pipelines.py:266-267, 336; filters.py:760-762)."""
    catalogDict = detectAndMeasure(job['ops'], job['batch'], job['filtersList'], job['data'], job['SNMap'], job['mapUnits'],
                                   job['parDict'], measureFluxes=measureFluxes, invertMap=job['invertMap'], detectJob=job['detectJob'])
    # small per-image results go to the host here, on the job's stream and before its pipeline slot is reused: with CUDA graphs the
    # device tensors are the graph's static outputs, which the batch submitted three submits later overwrites
    def host(v):
        if v is None:
            return None
        if isinstance(v, tuple):
            return tuple(host(x) for x in v)
        return device.toHost(v) if hasattr(v, 'device') and not isinstance(v, np.ndarray) else v
    extras = {'noiseTable': host(job['noiseTable']), 'noiseEdges': host(job['noiseEdges'])}
    if keepMaps:      # filter caches stay on the device (no graphs in that mode)
        extras.update({'signalNorm': job['signalNorm'], 'fRel': job['fRel']})
    else:
        extras.update({'signalNorm': host(job['signalNorm']), 'fRel': host(job['fRel'])})
    if job.get('rmsTabJob') is not None:
        from . import completeness
        extras['RMSTabs'] = completeness.finishRMSTables(job['rmsTabJob'])      # [(y0RMS, areaDeg2)] per image
    if keepMaps:
        extras.update({k: job[k] for k in ('data', 'SNMap', 'surveyMask', 'RMSMap', 'filtH')})
    S = len(job['filtersList'])
    names = job['batch'].tileNames
    if 'areaMask' in collect:
        first = device.toHost(job['surveyMask'][::S].contiguous())
        extras['areaMask'] = {name: first[i] for i, name in enumerate(names)}
    if 'RMSMap' in collect:
        first = device.toHost(job['RMSMap'][::S].to(job['ops'].torch.float32).contiguous())
        extras['RMSMapFirst'] = {name: first[i] for i, name in enumerate(names)}
    return catalogDict, extras


def filterTileBatch(ops, batch, filtersList, templates, parDict, undoPixelWindow=True, measureFluxes=True,
                    invertMap=False, keepMaps=False):
    """startTileBatch + finishTileBatch for one batch."""
    job = startTileBatch(ops, batch, filtersList, templates, parDict, undoPixelWindow=undoPixelWindow, invertMap=invertMap)
    return finishTileBatch(job, measureFluxes=measureFluxes, keepMaps=keepMaps)


class _BatchView(object):
    """The geometry of a TileBatch with other device buffers (the static inputs a captured CUDA graph reads)."""

    def __init__(self, batch, dev):
        self.dev = dev
        self.nT, self.F, self.H, self.W = batch.nT, batch.F, batch.H, batch.W
        self.wcsList, self.mapDicts, self.tileNames, self.flagMasks = batch.wcsList, batch.mapDicts, batch.tileNames, batch.flagMasks
        self.geometry = getattr(batch, 'geometry', None)


class BatchPipeline(object):
    """Two-stage software pipeline over tile batches: ``submit(batch)`` queues the device work of that batch (its upload on the
    slot's stream, its kernels -- one CUDA graph -- on the compute stream all batches share) and only then assembles the
    catalogs of the batch submitted before it (on a high-priority stream), so the host part of batch i runs while the GPU works
    on batch i + 1.  ``submit`` returns the (catalogDict, extras) of the previous
    batch (None for the first one); ``flush()`` returns the last.  This is how a run over many tiles keeps both sides busy;
    nothing is shared between batches except the read-only templates."""

    def __init__(self, ops, filtersList, templates, parDict, undoPixelWindow=True, measureFluxes=True, invertMap=False,
                 keepMaps=False, useCachedRMSMap=False, useGraphs=True, depth=2, rmsTables=False, serialize=False, collect=()):
        self.rmsTables = rmsTables
        self.collect = tuple(collect)
        self.serialize = serialize      # wait for every batch's device work before queueing the next (per-kernel timing runs)
        self.ops, self.filtersList, self.templates, self.parDict = ops, filtersList, templates, parDict
        # the ~40 launches of a batch are captured once per (slot, shape, input buffers) into a CUDA graph and replayed:
        # the host then spends microseconds, not a millisecond, queueing a batch.  Graphs live on the Ops object.
        self.useGraphs = useGraphs
        self.graphs = ops.__dict__.setdefault('_batchGraphs', {})
        self.undoPixelWindow, self.measureFluxes, self.invertMap, self.keepMaps = undoPixelWindow, measureFluxes, invertMap, keepMaps
        # the painted RMS maps are written only when something reads them: filter caches (keepMaps), noise tables, RMS files
        self.wantRMS = bool(keepMaps or rmsTables or 'RMSMap' in tuple(collect))
        self.useCachedRMSMap = useCachedRMSMap
        # `depth` batches may be queued on the device while the host assembles the catalogs of the oldest one: with two, the
        # upload of batch i + 2 is already in the copy queue when the host starts waiting for batch i, so the copy engine
        # never waits for the host.  Every slot has its own stream (and its own input buffers when it uploads).
        self.depth = max(1, int(depth))
        torch = ops.torch
        streams = getattr(ops, '_pipeStreams', None)
        if streams is None:
            streams = ops._pipeStreams = []
        while len(streams) < self.depth + 1:
            streams.append(torch.cuda.Stream(device=ops.device))
        self.streams = streams[:self.depth + 1]
        # the host part of a batch (finishTileBatch) launches a few small kernels and copies between its waits; on the batch's
        # own stream they queue behind the CTAs of the batches running on the other streams and each wait costs about a
        # millisecond.  They go to a high-priority stream instead: its CTAs are placed first whenever an SM has room.
        if getattr(ops, '_finishStream', None) is None:
            ops._finishStream = torch.cuda.Stream(device=ops.device, priority=-1)
        self.finishStream = ops._finishStream
        # the device work of ALL batches runs on one stream, one batch after the other: side by side on their slots' streams the
        # kernels of two or three batches interleave, evict each other's spectra and filters from L2 and take 64 ms per DR5 step
        # instead of the 55.5 ms they take alone.  The slots' streams only carry the uploads (and `prepare`).
        if getattr(ops, '_computeStream', None) is None:
            ops._computeStream = torch.cuda.Stream(device=ops.device)
        self.compute = ops._computeStream if os.environ.get('NEMO_B200_SERIAL_COMPUTE', '1') != '0' else None
        self.main = torch.cuda.current_stream(ops.device)
        for st in self.streams + ([self.compute] if self.compute is not None else []):
            st.wait_stream(self.main)      # templates and resident inputs were produced on the caller's stream
        self.pending = []
        self.count = 0

    def _finish(self):
        torch = self.ops.torch
        job, st, release, done = self.pending.pop(0)
        fin = self.finishStream
        fin.wait_event(done)      # this batch only: the batches queued behind it keep the device busy meanwhile
        with torch.cuda.stream(fin):
            out = finishTileBatch(job, measureFluxes=self.measureFluxes, keepMaps=self.keepMaps, collect=self.collect)
        # what the finish read must be free before the slot's next batch overwrites it (finishTileBatch normally ends with a copy to
        # the host that it waits for, so this returns at once)
        fin.synchronize()
        if release:
            job['batch'].dev = None
        return out

    def _graphEntry(self, slot, st, batch, upload, cache=None, templates=None):
        """(graph, job template, static inputs, launches per replay) for this slot and batch shape, captured on first use."""
        from . import _lib
        ops, torch = self.ops, self.ops.torch
        templates = self.templates if templates is None else templates
        if upload:
            ident = ('staged', str(batch.host['data'].dtype), str(batch.host['weights'].dtype))
        else:
            ident = tuple(int(batch.dev[k].data_ptr()) for k in ('data', 'weights', 'psMask', 'surveyMask', 'wmed'))
        # cached filters / RMS maps (the source-injection loop) are device buffers that persist between calls: part of the key
        cacheIdent = None
        if cache is not None:
            cacheIdent = (self.useCachedRMSMap,) + tuple(int(cache[k].data_ptr()) for k in ('filtH', 'signalNorm', 'RMSMap') if cache.get(k) is not None)
        key = (slot, ident, cacheIdent, batch.nT, batch.F, batch.H, batch.W, ops.precision, templates.serial, self.undoPixelWindow, self.invertMap, self.rmsTables, self.wantRMS,
               repr(sorted((k, repr(v)) for k, v in self.parDict.items() if k != 'mapFilters')),
               _filtersKey(self.filtersList))
        entry = self.graphs.get(key)
        if entry is not None:
            return entry
        with torch.cuda.stream(st):
            if upload:
                static = {k: torch.empty(v.shape, dtype=v.dtype, device=ops.device) for k, v in batch.host.items()}
                for k, v in batch.host.items():
                    static[k].copy_(v, non_blocking=True)
            else:
                static = dict(batch.dev)
            view = _BatchView(batch, dict(static))
            photInList = self.parDict.get('photFilter') is not None and any(f['label'] == self.parDict['photFilter'] for f in self.filtersList)
            needWeights = not (cache is not None and self.useCachedRMSMap and photInList and not self.rmsTables)
            if upload:
                widenStaged(ops, view.dev) if needWeights else widenStaged(ops, view.dev, keys=('data',))
                view.dev['wmed'] = ops.median_weights(view.dev['weights']) if needWeights else None
            # one eager run first: twiddle tables, kernel attributes and other lazy set-up must not happen under capture
            startTileBatch(ops, view, self.filtersList, templates, self.parDict, undoPixelWindow=self.undoPixelWindow,
                           invertMap=self.invertMap, rmsTables=self.rmsTables, cache=cache, useCachedRMSMap=self.useCachedRMSMap,
                           wantRMS=self.wantRMS)
            st.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launchCount()
            # graphs of one slot run one after the other on the slot's stream, so they capture into one memory pool: the
            # intermediates of a batch shape reuse those of the shapes captured before it (each graph keeps only its outputs)
            pools = ops.__dict__.setdefault('_graphPools', {})
            if slot not in pools:
                pools[slot] = torch.cuda.graph_pool_handle()
            with torch.cuda.graph(graph, pool=pools[slot], stream=st):
                if upload:
                    view.dev.update(static)          # the widening of float32-staged maps is part of the captured work
                    widenStaged(ops, view.dev) if needWeights else widenStaged(ops, view.dev, keys=('data',))
                    view.dev['wmed'] = ops.median_weights(view.dev['weights']) if needWeights else None
                job = startTileBatch(ops, view, self.filtersList, templates, self.parDict,
                                     undoPixelWindow=self.undoPixelWindow, invertMap=self.invertMap, rmsTables=self.rmsTables,
                                     cache=cache, useCachedRMSMap=self.useCachedRMSMap, wantRMS=self.wantRMS)
            nLaunch = _lib.launchCount() - n0
        entry = (graph, job, static, nLaunch, templates)      # the templates the graph reads stay alive with it
        self.graphs[key] = entry
        return entry

    def submit(self, batch, upload=True, cache=None, templates=None, prepare=None):
        """Queues the device work of `batch`; returns the (catalogDict, extras) of the oldest queued batch once more than
        `depth` - 1 others are in flight behind it, else None.  prepare(batch): device work to queue on the slot's stream
        before the batch itself (painting injected sources into its resident maps)."""
        from . import _lib
        torch = self.ops.torch
        slot = self.count % len(self.streams)
        st = self.streams[slot]
        self.count += 1
        if prepare is not None:
            with torch.cuda.stream(st):
                prepare(batch)
            upload = upload and batch.dev is None
        useGraph = self.useGraphs and not _lib.profEnabled()
        # with cached filters AND a cached RMS map nothing reads the weights (they only feed the noise estimate): a realisation of
        # the injection loop uploads its maps and masks, not the weight maps
        photInList = self.parDict.get('photFilter') is not None and any(f['label'] == self.parDict['photFilter'] for f in self.filtersList)
        skipKeys = ('weights', 'rowArea') if (cache is not None and self.useCachedRMSMap and photInList and not self.rmsTables) else ()
        run = st if self.compute is None else self.compute
        if useGraph:
            graph, template, static, nLaunch, _ = self._graphEntry(slot, st, batch, upload, cache, templates)
            with torch.cuda.stream(st):
                if upload:
                    for k, v in batch.host.items():
                        if k in skipKeys:
                            continue
                        static[k].copy_(device.countH2D(v), non_blocking=True)
            run.wait_stream(st)
            with torch.cuda.stream(run):
                graph.replay()
            _lib.load().nb_launch_count_add(nLaunch)
            job = dict(template)
            job['batch'] = batch
            upload = False          # nothing of the batch itself is resident
        else:
            with torch.cuda.stream(st):
                if upload:
                    batch.upload()
            run.wait_stream(st)
            with torch.cuda.stream(run):
                job = startTileBatch(self.ops, batch, self.filtersList, self.templates if templates is None else templates, self.parDict,
                                     undoPixelWindow=self.undoPixelWindow, invertMap=self.invertMap, cache=cache,
                                     useCachedRMSMap=self.useCachedRMSMap, rmsTables=self.rmsTables, wantRMS=self.wantRMS)
        done = torch.cuda.Event()
        done.record(run)
        st.wait_event(done)       # the slot's next upload overwrites this batch's inputs
        if self.serialize:
            run.synchronize()
        self.pending.append((job, st, upload, done))
        # a slot's buffers (graph outputs, static inputs) are reused by the batch submitted len(streams) calls later, so at
        # most `depth` jobs may stay unread
        return self._finish() if len(self.pending) > self.depth else None

    def flush(self):
        """The results of every batch still in flight, oldest first (a list; empty when nothing is pending)."""
        outs = []
        while self.pending:
            outs.append(self._finish())
        for st in self.streams + ([self.compute] if self.compute is not None else []):
            self.main.wait_stream(st)
        return outs


def processBatches(ops, batches, filtersList, templates, parDict, upload=True, undoPixelWindow=True, measureFluxes=True,
                   invertMap=False, keepMaps=False, caches=None, useCachedRMSMap=False, useGraphs=True, depth=2, rmsTables=False,
                   beforeSubmit=None, collect=(), prepare=None):
    """Runs several tile batches through a BatchPipeline; returns the list of (catalogDict, extras), one per batch, in order.
    caches: per-batch filter / RMS caches (see startTileBatch), or None.  templates: one FilterSetTemplates, or one per batch.
    beforeSubmit(batch): host work to do right before a batch is queued (re-staging its maps); it runs while the device
    works on the batches queued before it."""
    pipe = BatchPipeline(ops, filtersList, templates[0] if isinstance(templates, (list, tuple)) else templates, parDict,
                         undoPixelWindow=undoPixelWindow, measureFluxes=measureFluxes,
                         invertMap=invertMap, keepMaps=keepMaps, useCachedRMSMap=useCachedRMSMap,
                         useGraphs=useGraphs and not keepMaps, depth=depth, rmsTables=rmsTables, collect=collect)
    perBatch = isinstance(templates, (list, tuple))
    results = []
    for i, batch in enumerate(batches):
        if beforeSubmit is not None:
            beforeSubmit(batch)
        out = pipe.submit(batch, upload=upload and batch.dev is None, cache=caches[i] if caches is not None else None,
                          templates=templates[i] if perBatch else None, prepare=prepare)
        if out is not None:
            results.append(out)
    results.extend(pipe.flush())
    return results


def catalogKeys(havePhotFilter, mapUnits, measureFluxes=True, measureShapes=False):
    """Column list of the per-(tile, scale) catalogs, the same on every rank (what catalogListToTab is asked to write)."""
    keys = [k for k in catalogs.COLUMN_NAMES if measureShapes == True or not k.startswith('ellipse_')]
    prefixes = ['', 'fixed_'] if havePhotFilter else ['']
    if havePhotFilter:
        keys.append('fixed_SNR')
    if measureFluxes:
        for prefix in prefixes:
            keys += [prefix + k for k in FLUX_KEYS[mapUnits]]
    return keys


def detectAndMeasure(ops, batch, filtersList, data, sn, mapUnits, parDict, measureFluxes=True, invertMap=False, detectJob=None):
    """K9-K11 for all (tile, scale) images, then the rows of every catalog (pipelines.py:251-284)."""
    nT, S = batch.nT, len(filtersList)
    H, W = batch.H, batch.W
    wcsList = [batch.wcsList[o // S] for o in range(nT * S)]
    flagList = [batch.flagMasks[o // S] for o in range(nT * S)]
    labelList = [filtersList[o % S]['label'] for o in range(nT * S)]
    tileList = [batch.tileNames[o // S] for o in range(nT * S)]
    if detectJob is None:
        detectJob = photometry.startFindObjectsBatch(ops, sn, threshold=parDict['thresholdSigma'],
                                                     findCenterOfMass=parDict['findCenterOfMass'], removeRings=parDict['removeRings'],
                                                     ringThresholdSigma=parDict['ringThresholdSigma'], invertMap=invertMap,
                                                     measureShapes=parDict['measureShapes'])
    photFilter = parDict.get('photFilter')
    photScale = None
    if photFilter is not None:
        for s, f in enumerate(filtersList):
            if f['label'] == photFilter:
                photScale = s
    extraSamples = []
    if photScale is not None:
        photIndex = lambda o: (o // S) * S + photScale
        extraSamples.append(('fixed_SNR', sn, photIndex, -1.0 if (invertMap and not measureFluxes) else 1.0))
    if measureFluxes:
        extraSamples.append(('', data, lambda o: o, 1.0))
        if photScale is not None:
            extraSamples.append(('fixed_', data, photIndex, 1.0))
    flat = photometry.finishFindObjectsBatch(detectJob, wcsList, flagList, labelList, tileList, minObjPix=parDict['minObjPix'],
                                             objIdent=parDict['objIdent'], longNames=parDict['longNames'],
                                             useInterpolator=parDict['useInterpolator'], extraSamples=extraSamples)
    cols = {}
    if flat['cols'] is not None and extraSamples:
        ownArr = flat['owner']
        vals = flat['samples']
        snrAll = flat['cols']['SNR']
        if photScale is not None:
            # bounds test of getSNRValues (photometry.py:246-253): int() truncates towards zero
            xs, ys = flat['xy2']
            xi, yi = np.trunc(xs), np.trunc(ys)
            inside = (xi > 0) & (xi < W) & (yi > 0) & (yi < H)
            cols['fixed_SNR'] = np.where(inside, vals['fixed_SNR'], 0.0)
        if measureFluxes:
            for prefix in ['', 'fixed_'] if photScale is not None else ['']:
                v = vals[prefix]
                snr = snrAll if prefix == '' else cols['fixed_SNR']
                with np.errstate(divide='ignore', invalid='ignore'):
                    if mapUnits == 'yc':
                        cols[prefix + 'y_c'] = v / 1e-4
                        cols[prefix + 'err_y_c'] = cols[prefix + 'y_c'] / snr
                        deltaTc = maps.convertToDeltaT(v, obsFrequencyGHz=148.0)
                        cols[prefix + 'deltaT_c'] = deltaTc
                        cols[prefix + 'err_deltaT_c'] = np.abs(deltaTc / snr)
                    else:
                        cols[prefix + 'deltaT_c'] = v
                        cols[prefix + 'err_deltaT_c'] = v / snr
                        # Jy fluxes need the beam solid angle and frequency of the image's tile (photometry.py:336-339)
                        nsr = np.array([_filters.beamSolidAngle_nsr(batch.mapDicts[o // S][0]) for o in range(nT * S)], dtype=np.float64)[ownArr]
                        freq = np.array([batch.mapDicts[o // S][0]['obsFreqGHz'] for o in range(nT * S)], dtype=np.float64)[ownArr]
                        flux = photometry.deltaTToJyPerSr(v, freq) * nsr * 1.e-9
                        errFlux = photometry.deltaTToJyPerSr(cols[prefix + 'err_deltaT_c'], freq) * nsr * 1.e-9
                        cols[prefix + 'fluxJy'] = np.where(nsr > 0, flux, 0.0)
                        cols[prefix + 'err_fluxJy'] = np.where(nsr > 0, errFlux, 0.0)
    keys = catalogKeys(photScale is not None, mapUnits, measureFluxes, parDict.get('measureShapes', False))
    # the batch's rows as one table: what the cross-scale merge stacks (filterMapsAndMakeCatalogs hands these to
    # makeOptimalCatalog); the per-image catalogs are row ranges of it
    stacked = photometry.flatTable(flat, cols, keysToWrite=keys)
    tables = photometry.splitObjects(flat, cols, keysToWrite=keys, stacked=stacked)
    catalogDict = _BatchCatalogs()
    for o in range(nT * S):
        catalogDict[labelList[o] + "#" + tileList[o]] = {'catalog': tables[o]}
    catalogDict.flat = stacked
    return catalogDict


class _BatchCatalogs(dict):
    """catalogDict of one tile batch ('<label>#<tileName>' -> {'catalog': ...}, the reference's layout) that also carries the
    batch's rows as one table (.flat)."""
    flat = None


# ------------------------------------------------------------------------------------------------------
def _orderedFilters(parDict):
    """photFilter first, then the others in config order (pipelines.py:153-167)."""
    photFilter = parDict.get('photFilter')
    filtersList = []
    if photFilter is not None:
        for f in parDict['mapFilters']:
            if f['label'] == photFilter:
                filtersList.append(f)
    for f in parDict['mapFilters']:
        if photFilter is not None and f['label'] == photFilter:
            continue
        filtersList.append(f)
    return filtersList


def _canBatch(filtersList):
    """True when every filter is a Fourier-space matched filter and they share the noise set-up (what the batched kernels
    take); anything else -- real-space filters, mixed noise parameters -- goes through the reference's per-tile loop."""
    for f in filtersList:
        cls = f['class'] if not isinstance(f['class'], str) else getattr(_filters, f['class'], None)
        if cls is None or not issubclass(cls, _filters.MatchedFilter):
            return False
    return _sameNoiseSetup(filtersList)


def _perTileLoop(config, filtersList, measureFluxes=True, invertMap=False, verbose=True, useCachedFilters=False):
    """The reference's own loop (pipelines.py:179-286) over this rank's tiles and the filters, one filterMaps /
    findObjects / measureFluxes call at a time, for filter sets the batched path does not take."""
    parDict = config.parDict
    photFilter = parDict.get('photFilter')
    catalogDict = {}
    for tileName in config.tileNames:
        if verbose:
            print(">>> [rank = %d] Making filtered maps - tileName = %s ..." % (config.rank, tileName))
        photFilteredMapDict = None
        for f in filtersList:
            label = f['label'] + "#" + tileName
            fCopy = {'label': f['label'], 'class': f['class'], 'params': dict(f['params'])}
            filteredMapDict = _filters.filterMaps(config.tileMaps[tileName], fCopy, tileName,
                                                  diagnosticsDir=getattr(config, 'diagnosticsDir', None),
                                                  selFnDir=getattr(config, 'selFnDir', None), verbose=verbose,
                                                  undoPixelWindow=parDict.get('undoPixelWindow', True), useCachedFilter=useCachedFilters)
            if invertMap:
                raise Exception("invertMap is only handled by the batched path")
            if f['label'] == photFilter:
                photFilteredMapDict = {'SNMap': filteredMapDict['SNMap'], 'data': filteredMapDict['data']}
            catalog = photometry.findObjects(filteredMapDict, threshold=parDict['thresholdSigma'], minObjPix=parDict['minObjPix'],
                                             findCenterOfMass=parDict['findCenterOfMass'], removeRings=parDict['removeRings'],
                                             ringThresholdSigma=parDict['ringThresholdSigma'], rejectBorder=parDict['rejectBorder'],
                                             objIdent=parDict['objIdent'], longNames=parDict['longNames'],
                                             useInterpolator=parDict['useInterpolator'], measureShapes=parDict['measureShapes'],
                                             verbose=False)
            if measureFluxes:
                photometry.measureFluxes(catalog, filteredMapDict, getattr(config, 'diagnosticsDir', None),
                                         photFilteredMapDict=photFilteredMapDict, useInterpolator=parDict['useInterpolator'])
            elif photFilter is not None:
                photometry.getSNRValues(catalog, photFilteredMapDict['SNMap'], filteredMapDict['wcs'], prefix='fixed_',
                                        useInterpolator=parDict['useInterpolator'], invertMap=invertMap)
            catalogDict[label] = {'catalog': catalog}
    return catalogDict


def stageTiles(config, tilesPerBatch=8, precision=None):
    """Copies this rank's tiles into pinned host staging buffers, `tilesPerBatch` tiles per batch, and keeps them on
    `config` so that repeated runs (source-injection iterations, benchmarks) start from pinned memory."""
    ops = device.getOps(precision)
    # tiles of one geometry (shape and pixel scales at the tile centre: CAR tiles of one declination row) form the batches;
    # groups keep the order in which their first tile appears
    groups = {}
    for name in config.tileNames:
        groups.setdefault((tileGeometry(config.tileMaps[name]), len(config.tileMaps[name])), []).append(name)
    # (a list of batch sizes ramps up at the start of the run and down at its end, not at every change of geometry)
    batches = []
    for g, names in enumerate(groups.values()):
        bounds = _batchBounds(len(names), tilesPerBatch, rampUp=g == 0, rampDown=g == len(groups) - 1)
        batches += [TileBatch(ops, names[i:j], config.tileMaps) for i, j in bounds]
    config._staged = {'tilesPerBatch': _batchKey(tilesPerBatch), 'precision': ops.precision, 'batches': batches}
    return config._staged['batches']


def _writeSelFnMaps(config, batches, results, label, comm=None):
    """The area mask (multi-extension, one per tile) and the RMS maps of filter `label` under config.selFnDir."""
    import os
    from . import fitsio
    selFnDir = getattr(config, 'selFnDir', None)
    if selFnDir is None:
        raise Exception("writeAreaMask / writeRMSMaps need config.selFnDir")
    os.makedirs(selFnDir, exist_ok=True)
    wcsOf = {name: b.wcsList[i] for b in batches for i, name in enumerate(b.tileNames)}
    masks, count = [], 0
    for _, extras in results:
        for name, m in (extras.get('areaMask') or {}).items():
            masks.append((name, m, wcsOf[name]))
        for name, rms in (extras.get('RMSMapFirst') or {}).items():
            os.makedirs(selFnDir + os.path.sep + name, exist_ok=True)
            fitsio.saveFITS(selFnDir + os.path.sep + name + os.path.sep + "RMSMap_%s#%s.fits" % (label, name), rms, wcsOf[name],
                            compressionType='RICE_1')
            count += 1
    if masks:
        multi = comm is not None and comm.size > 1
        fitsio.writeMEF(selFnDir + os.path.sep + ("areaMask_rank%d.fits" % comm.rank if multi else "areaMask.fits"), masks,
                        compressionType='PLIO_1')


def _batchKey(tilesPerBatch):
    return int(tilesPerBatch) if isinstance(tilesPerBatch, (int, np.integer)) else tuple(int(v) for v in tilesPerBatch)


def _batchBounds(numTiles, tilesPerBatch, rampUp=True, rampDown=True):
    """[(first, last + 1)] of the batches.  tilesPerBatch: a batch size, or a sequence of sizes -- the last size repeats until
    the tiles run out, except that the list is mirrored at the end when it starts small: [2, 4] gives 2, 4, 4, ..., 4, 2.  A
    small first batch starts the device work sooner (nothing overlaps its upload), a small last one shortens the tail that
    no upload overlaps.  rampUp / rampDown = False: without the small batches at the start / the mirrored ones at the end (the
    tiles are one geometry group in the middle of a run)."""
    if isinstance(tilesPerBatch, (int, np.integer)):
        sizes = [int(tilesPerBatch)] * ((numTiles + int(tilesPerBatch) - 1) // int(tilesPerBatch))
    else:
        head = [int(v) for v in tilesPerBatch]
        if not head or min(head) < 1:
            raise Exception("tilesPerBatch needs positive batch sizes")
        tail = head[:-1][::-1] if rampDown else []
        if not rampUp:
            head = head[-1:]
        sizes = []
        left = numTiles - sum(tail)
        for v in head:
            if left <= 0:
                break
            sizes.append(min(v, left))
            left -= sizes[-1]
        while left > 0:
            sizes.append(min(head[-1], left))
            left -= sizes[-1]
        left = numTiles - sum(sizes)
        for v in tail:
            if left <= 0:
                break
            sizes.append(min(v, left))
            left -= sizes[-1]
    if not isinstance(tilesPerBatch, (int, np.integer)):
        # a remainder in front of the mirrored tail joins the tail's first batch
        i = len(tilesPerBatch) - 1 if rampUp else 0
        while i + 1 < len(sizes):
            if sizes[i] < sizes[i + 1] and sizes[i] + sizes[i + 1] <= int(np.max(tilesPerBatch)):
                sizes[i:i + 2] = [sizes[i] + sizes[i + 1]]
            else:
                i += 1
    bounds, first = [], 0
    for v in sizes:
        if first >= numTiles:
            break
        bounds.append((first, min(first + v, numTiles)))
        first += v
    return bounds


def filterMapsAndMakeCatalogs(config, useCachedFilters=False, useCachedRMSMap=False, measureFluxes=True, invertMap=False,
                              verbose=True, tilesPerBatch=8, precision=None, comm=None, cacheFilters=False, restage=False,
                              makeRMSTables=False, restageKeys=('data',), writeAreaMask=False, writeRMSMaps=False,
                              injectCatalogs=None, injectOptions=None):
    """Runs the map filtering and catalog construction steps for this rank's tiles and returns the optimal catalog
    (highest S/N detection across filter scales), gathered over ranks when `comm` (a
    ``nemo_b200.distributed.Comm``) is given -- pipelines.py:33-109, 112-354.

    cacheFilters=True keeps every batch's filters, normalisations and RMS maps on the device (``config._cache``); a later
    call with useCachedFilters=True (and useCachedRMSMap=True) then only applies them to whatever the map dictionaries hold
    now -- the reference's source-injection loop (maps.py:2093), where the caches are FITS files on disk.  restage=True
    re-reads the `restageKeys` arrays ('data'; 'weights') of the map dictionaries into the pinned staging buffers (they were
    edited in place or replaced since the tiles were staged), batch by batch, each right before it is queued so that the copy
    runs under the device work of the batches before it.

    injectCatalogs = {tileName: catalog} paints model sources or clusters into the tiles ON THE DEVICE before they are filtered
    (TileBatch.inject: the painter of the source-injection loop, maps.py:1734-1899, 2093) -- the maps stay resident from one
    call to the next and only the catalogs travel; injectOptions: override / GNFWParams / minSNR / validAreaSection /
    applyPixelWindow of makeModelImage.  An empty dict restores the original maps.

    writeAreaMask / writeRMSMaps (need ``config.selFnDir``): this rank's ``areaMask.fits`` (one extension per tile; with several
    ranks ``areaMask_rank<r>.fits``) and ``<tile>/RMSMap_<filter>#<tile>.fits`` of the first filter of the list (the photFilter when
    there is one), as the reference writes them (pipelines.py:266-267, 336; filters.py:760-762) -- see nemo_b200/fitsio.py.

    Batched here: Fourier-space matched filters with a common noise set-up (startTileBatch), and real-space matched filters
    whose kernel is built from the whole tile and symmetrised (nemo_b200/realspace.py).  Any other filter set goes through the
    reference's per-tile loop over the class API (_perTileLoop).
    """
    ops = device.getOps(precision)
    parDict = config.parDict
    filtersList = _orderedFilters(parDict)
    from . import realspace
    realSpaceBatched = (not _canBatch(filtersList)) and realspace.canBatch(filtersList, config.tileMaps) and not (
        cacheFilters or useCachedFilters or useCachedRMSMap or restage or injectCatalogs is not None or writeAreaMask or writeRMSMaps)
    if not _canBatch(filtersList) and not realSpaceBatched:
        if cacheFilters or useCachedRMSMap:
            raise Exception("filter caches on the device are kept by the batched (Fourier matched filter) path only")
        catalogDict = _perTileLoop(config, filtersList, measureFluxes=measureFluxes, invertMap=invertMap, verbose=verbose,
                                   useCachedFilters=useCachedFilters)
        optimalCatalog = catalogs.makeOptimalCatalog(catalogDict, constraintsList=parDict.get('catalogCuts', []))
        if comm is not None and comm.size > 1:
            from . import distributed
            keys = [k for k in optimalCatalog.keys()] if len(optimalCatalog) > 0 else catalogKeys(False, 'yc', measureFluxes)
            nKeys = int(comm.maxOverRanks(len(keys)))
            if len(keys) != nKeys:      # a rank without detections: take the longest list any rank has
                raise Exception("ranks disagree on the catalog columns; give every rank at least one tile with detections")
            optimalCatalog = distributed.allGatherCatalog(comm, optimalCatalog, keys, [f['label'] for f in filtersList], config.allTileNames,
                                                          objIdent=parDict['objIdent'], longNames=parDict['longNames'])
        config.catalogDict = catalogDict
        config.noiseTables = {}
        config.RMSTabs = {}
        return optimalCatalog
    catalogDict = {}
    noiseTables = {}
    staged = getattr(config, '_staged', None)
    if staged is None or staged['tilesPerBatch'] != _batchKey(tilesPerBatch) or staged['precision'] != ops.precision:
        batches = stageTiles(config, tilesPerBatch, precision)
    else:
        batches = staged['batches']
    beforeSubmit = (lambda b: b.refresh(config.tileMaps, keys=restageKeys)) if restage else None
    prepare = None
    if injectCatalogs is not None:
        if restage:
            raise Exception("injectCatalogs paints into the maps resident on the device; restage re-uploads them -- use one or the other")
        opts = dict(injectOptions or {})
        prepare = lambda b: b.inject(injectCatalogs, **opts)
    templatesList = [templatesFor(config, ops, filtersList, b) for b in batches]
    templates = templatesList[0]
    if verbose:
        for batch in batches:
            print(">>> [rank = %d] Making filtered maps - tiles %s" % (config.rank, batch.tileNames))
    caches = None
    if useCachedFilters:
        caches = getattr(config, '_cache', None)
        if caches is None or len(caches) != len(batches):
            raise Exception("no cached filters for these tile batches: run filterMapsAndMakeCatalogs(config, cacheFilters=True) first")
    if realSpaceBatched:
        # real-space matched filters (filters.py:862-1218): kernels of all tiles of a batch built and applied together
        # (nemo_b200/realspace.py); the next batch is uploaded while this one's catalogs are assembled
        results = []
        for i, batch in enumerate(batches):
            uploadedHere = batch.dev is None
            if uploadedHere:
                batch.upload()
            job = realspace.startRealSpaceTileBatch(ops, batch, filtersList, templatesList[i], parDict,
                                                    undoPixelWindow=parDict.get('undoPixelWindow', True), invertMap=invertMap,
                                                    rmsTables=makeRMSTables)
            if i + 1 < len(batches) and batches[i + 1].dev is None:
                batches[i + 1].upload()
                batches[i + 1]._uploadedByDriver = True
            results.append(finishTileBatch(job, measureFluxes=measureFluxes))
            if uploadedHere or getattr(batch, '_uploadedByDriver', False):
                batch.dev = None
                batch._uploadedByDriver = False
    else:
        # with a cached RMS map the pixel window is undone after the S/N map has been made (pipelines.py:171-173, 216-232); the
        # batched path always makes the S/N map before undoing the window, so the same flag serves both cases
        results = processBatches(ops, batches, filtersList, templatesList, parDict, upload=True,
                                 undoPixelWindow=parDict.get('undoPixelWindow', True), measureFluxes=measureFluxes, invertMap=invertMap,
                                 keepMaps=cacheFilters, caches=caches, useCachedRMSMap=useCachedRMSMap, rmsTables=makeRMSTables,
                                 beforeSubmit=beforeSubmit, prepare=prepare,
                                 collect=(('areaMask',) if writeAreaMask else ()) + (('RMSMap',) if writeRMSMaps else ()))
    if writeAreaMask or writeRMSMaps:
        _writeSelFnMaps(config, batches, results, filtersList[0]['label'], comm)
    if cacheFilters:
        config._cache = [{'filtH': ex['filtH'], 'signalNorm': ex['signalNorm'], 'RMSMap': ex['RMSMap'], 'fRel': ex['fRel']}
                         for _, ex in results]
    RMSTabs = {}
    flatTables = []
    from . import distributed
    allTileNames = list(getattr(config, 'allTileNames', None) or config.tileNames)
    tileIndexOf = {n: i for i, n in enumerate(allTileNames)}
    noiseBlocks, noiseShape = [], None
    for batch, (cd, extras) in zip(batches, results):
        names = batch.tileNames
        catalogDict.update(cd)
        flatTables.append(getattr(cd, 'flat', None) if isinstance(cd, _BatchCatalogs) else False)
        for o, tab in enumerate(extras.get('RMSTabs') or []):
            RMSTabs[filtersList[o % len(filtersList)]['label'] + "#" + names[o // len(filtersList)]] = tab
        if extras['noiseTable'] is None:
            continue
        # the batch's tables stay one array with (filter, tile) index columns in front: the '<label>#<tileName>' dictionary is a
        # view made when somebody reads it, here and after the gather (hundreds of entries per step otherwise)
        tab = np.asarray(extras['noiseTable'])
        o = np.arange(tab.shape[0])
        for n in names:      # (a configuration object without allTileNames, or one that does not list this rank's tiles)
            if n not in tileIndexOf:
                if comm is not None and comm.size > 1:
                    raise Exception("config.allTileNames must list the tiles of every rank (tile %s is missing)" % n)
                tileIndexOf[n] = len(allTileNames)
                allTileNames.append(n)
        tileIdx = np.array([tileIndexOf[n] for n in names], dtype=np.int64)[o // len(filtersList)]
        noiseBlocks.append(distributed.noiseTableBlock(tab, o % len(filtersList), tileIdx))
        noiseShape = tuple(tab.shape[1:])
    # (every batch brought its rows as one table: stack those few instead of the hundreds of per-image tables -- same rows, same order)
    stacked = [t for t in flatTables if t is not None] if all(t is not False for t in flatTables) else None
    optimalCatalog = catalogs.makeOptimalCatalog(catalogDict, constraintsList=parDict.get('catalogCuts', []), ops=ops, stacked=stacked)
    if comm is not None and comm.size > 1:
        from . import distributed
        havePhot = parDict.get('photFilter') is not None and any(f['label'] == parDict['photFilter'] for f in filtersList)
        keys = catalogKeys(havePhot, templates.outputUnits, measureFluxes, parDict.get('measureShapes', False))
        labels = [f['label'] for f in filtersList]
        optimalCatalog = distributed.allGatherCatalog(comm, optimalCatalog, keys, labels, config.allTileNames,
                                                      objIdent=parDict['objIdent'], longNames=parDict['longNames'], ops=ops)
    if makeRMSTables and comm is not None and comm.size > 1:
        from . import distributed
        RMSTabs = distributed.gatherRMSTabs(comm, RMSTabs, [f['label'] for f in filtersList], config.allTileNames)
    noiseTables = distributed.gatherNoiseBlocks(comm, noiseBlocks, [f['label'] for f in filtersList], allTileNames,
                                                noiseShape if noiseShape is not None and len(noiseShape) == 2 else None)
    config.catalogDict = catalogDict
    config.noiseTables = noiseTables
    config.RMSTabs = RMSTabs      # {'<label>#<tileName>': (y0RMS ascending, areaDeg2)} -- completeness.getRMSTab per tile and filter
    return optimalCatalog


# ------------------------------------------------------------------------------------------------------
def fitQ(config, precision=None, tilesPerBatch=4, verbose=False, y0=2e-04, scales=None):
    """signals.fitQ (nemo/signals.py:864-1125) for this rank's tiles: model clusters of ~55 angular sizes are painted at the tile
    centre (beam-convolved, pixel window applied), filtered with the tile's reference (photFilter) filter, and the peak of the
    filtered signal -- a bicubic spline at the centre -- relative to the input y0 is Q(theta500).  Everything per-pixel runs on
    the device: the painter (nb_paint_objects + one Fourier multiplication for beam and window), K1 on the models, K2/K3/K5 for
    the tile's filter, K4 for every (tile, size), and the spline of K5.  Model spectra are shared by the tiles of one geometry.

    Returns {tileName: Table with columns Q, theta500Arcmin, z, sorted by theta500Arcmin}; with ``config.selFnDir`` set the
    tables are also written to ``selFnDir/QFit.fits`` as one binary-table extension per tile with ZDEPQ / TILENAME keywords,
    the layout the reference writes (:1086-1117) and signals.QFit reads.  scales: (zRange, MRange) to use instead of
    signals.fitQScales -- the reference scale must come first."""
    import os
    ops = device.getOps(precision)
    torch = ops.torch
    parDict = config.parDict
    ref = None
    for f in parDict['mapFilters']:
        if f['label'] == parDict.get('photFilter'):
            ref = f
    if ref is None:
        raise Exception("couldn't find filter that matches photFilter")
    className = ref['class'] if isinstance(ref['class'], str) else ref['class'].__name__
    if className.find("Arnaud") == -1:
        raise Exception("Signal model for Q calculation should be 'Arnaud' ('Battaglia' needs pyccl)")
    if className.find("RealSpace") != -1:
        raise Exception("fitQ with a real-space reference filter goes through the class API (filters.RealSpaceMatchedFilter.applyFilter)")
    refEntry = {'label': ref['label'], 'class': ref['class'], 'params': dict(ref['params'])}
    zRange, MRange = scales if scales is not None else signals.fitQScales(ref['params']['M500MSun'], ref['params']['z'])
    nS = len(zRange)
    GNFWParams = ref['params'].get('GNFWParams', parDict.get('GNFWParams', 'default'))
    staged = stageTiles(config, tilesPerBatch, precision)
    QTabDict = {}
    spectraByGeometry = {}
    for batch in staged:
        nT, F, H, W = batch.nT, batch.F, batch.H, batch.W
        Wc = W // 2 + 1
        templates = templatesFor(config, ops, [refEntry], batch)
        dev = batch.dev if batch.dev is not None else batch.upload()
        md0 = batch.mapDicts[0]
        # model spectra of every size and frequency, once per geometry (signals.py:1029-1055: makeSignalModelMap at the tile
        # centre with amplitude y0 -> delta T, beam-convolved, enmap.apply_window, enmap.fft)
        key = (batch.geometry, _beamsKey(md0), repr(GNFWParams), nS, float(MRange[-1]))
        if key not in spectraByGeometry:
            wcs0 = batch.wcsList[0]
            ra0, dec0 = wcs0.getCentreWCSCoords()
            wcsList, cats, beams, freqs = [], [], [], []
            for z, M in zip(zRange, MRange):
                for m in md0:
                    wcsList.append(wcs0)
                    cats.append({'RADeg': np.array([ra0]), 'decDeg': np.array([dec0]), 'true_y_c': np.array([y0 / 1e-4]),
                                 'true_M500c': np.array([M / 1e14]), 'redshift': np.array([z])})
                    beams.append(m['beamFileName'])
                    freqs.append(None if m.get('units') == 'yc' else m['obsFreqGHz'])
            model, _ = maps.paintModelImages(ops, (H, W), wcsList, cats, beams, freqs, GNFWParams=GNFWParams, applyPixelWindow=True,
                                              maxSizeDeg=15.0)
            spectraByGeometry[key] = ops.rfft2(model, apodPix=0).reshape(nS, F, H, Wc)
            del model
        spectra = spectraByGeometry[key]
        # the tiles' reference filters (filters.py:524-696), then every size through every tile's filter
        params0 = refEntry['params']
        bck = None
        if 'bckSub' in params0.keys() and 'bckSubScaleArcmin' in params0.keys() and params0['bckSub'] == True:
            bck = maps.smoothScalesPix(batch.wcsList[0], smoothScaleDeg=params0['bckSubScaleArcmin'] / 60.)
        fmaps = ops.rfft2(dev['data'].reshape(nT * F, H, W), apodPix=20).reshape(nT, F, H, Wc)
        res = engine.buildFourierFilters(ops, fmaps, templates.unit, templates.norm, templates.w, templates.outputUnits,
                                         templates.cy, templates.cx, bckSubSigmas=bck, spectra=templates._spectra)
        del fmaps
        for t, tileName in enumerate(batch.tileNames):
            filt = res['filtH'][t].unsqueeze(0).expand(nS, F, H, Wc).contiguous()
            norm = res['signalNorm'][t].reshape(1).expand(nS).contiguous()
            filtered = engine.applyFourierFilters(ops, spectra, filt, 1, norm, None, bckSubSigmas=bck, W=W)
            _, _, _, centre = ops.template_norm(filtered, 1, templates.cy, templates.cx, 1.0)
            peak = device.toHost(centre)
            del filt, filtered
            Q, theta, zs = [], [], []
            for v, z, M in zip(peak, zRange, MRange):
                if v not in Q:                        # signals.py:1068
                    Q.append(float(v))
                    theta.append(signals.calcTheta500Arcmin(z, M))
                    zs.append(z)
            Q = np.array(Q)
            if abs(1 - Q[0] / y0) > 1e-6:
                raise Exception("Q[0]/y0 outside tolerance")
            Q = Q / y0
            order = np.argsort(np.array(theta), kind='stable')
            QTabDict[tileName] = catalogs.Table.fromColumns([('Q', Q[order]), ('theta500Arcmin', np.array(theta)[order]),
                                                             ('z', np.array(zs)[order])])
            if verbose:
                print("... fitted Q in tile %s" % (tileName))
    selFnDir = getattr(config, 'selFnDir', None)
    if selFnDir is not None:
        from . import fitsio
        os.makedirs(selFnDir, exist_ok=True)
        fitsio.writeTableMEF(selFnDir + os.path.sep + "QFit.fits",
                             [(name, {k: np.asarray(tab[k]) for k in ('Q', 'theta500Arcmin', 'z')}, {'NEMOVER': 'nemo_b200', 'ZDEPQ': 0, 'TILENAME': name})
                              for name, tab in QTabDict.items()])
    return QTabDict
