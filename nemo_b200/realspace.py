"""
Batched driver of the real-space matched filter (nemo/filters.py:862-1218, RealSpaceMatchedFilter.buildKernel /
buildAndApply / applyFilter) for tile batches: every tile of a batch at every real-space filter of the list in single
launches, image index o = tile * S + filter, the layout of the Fourier path (engine.py / pipelines.startTileBatch).

Per batch:
  1. K1-K3 on the whole tiles: apodised spectra, smoothed noise covariance, the nested Fourier matched filter of every
     (tile, filter) -- what ``matchedFilterClass(...).buildAndApply()`` is run for in buildKernel (:941-942); its own
     normalisation and filtered map are never looked at and are not computed.
  2. the filter's 1-d real-space profile: row 0 of ifft2(filt).real is the inverse transform along x of the ky-mean of the
     filter, so only that mean (``nb_mean_over_rows`` on the folded half-plane filter) leaves the device, [nT S F, Wc] numbers;
     profile, symmetrised 2-d kernel and its cut-out are small host work shared with the class API
     (filters.realSpaceProfileFromRow / realSpaceKernelFromProfile).
  3. signalNorm and fRelWeights: the signal templates, high-passed once per geometry, through every (tile, filter) kernel
     (``nb_conv2d_batch`` with the templates shared by the tiles); maximum and its position on the device, the per-frequency
     values at that pixel from the host copy of the high-passed templates.
  4. the maps: Gaussian high-pass (K12), kernel convolution with normalisation and point-source mask fused (K13), then the
     noise / S-N / mask block, pixel-window undo and detection exactly as for Fourier filters.

What this driver does not take (the caller falls back to the per-tile class loop): a noise region other than the whole tile
(RADecSection), kernels that are not symmetrised (they are cut out of the 2-d inverse transform of the filter), a background
scale derived from the profile (no 'bckSubScaleArcmin' in the filter parameters), separate raw maps ('rawData'), cached
kernels on disk.
"""

import numpy as np
from numpy import fft

from . import device, engine, maps, photometry
from . import filters as _filters


def _cls(f):
    return f['class'] if not isinstance(f['class'], str) else getattr(_filters, f['class'], None)


def canBatch(filtersList, tileMaps=None):
    """True when every filter is a real-space matched filter this driver handles and they share noise set-up and units."""
    keys = set()
    for f in filtersList:
        cls = _cls(f)
        if cls is None or not issubclass(cls, _filters.RealSpaceMatchedFilter):
            return False
        p = f['params']
        n = p['noiseParams']
        if n.get('RADecSection') is not None or n.get('symmetrize') is not True or n.get('method') != 'dataMap':
            return False
        if p.get('noiseModelCatalog') is not None:
            return False
        if p.get('bckSub') == True and 'bckSubScaleArcmin' not in p.keys():
            return False
        if p.get('saveFilter') or p.get('saveRMSMap'):
            return False
        keys.add((repr(sorted(n.items())), p.get('edgeTrimArcmin'), p.get('bckSub'), p.get('bckSubScaleArcmin'), p.get('outputUnits')))
    if len(keys) != 1:
        return False
    if tileMaps is not None:
        for md in tileMaps.values():
            for m in md:
                if hasattr(m, 'loadTile') or m.get('rawData') is not None:
                    return False
    return True


def _bckSigmas(batch, params0):
    """(sigmaY, sigmaX) in pixels of the background-subtraction Gaussian (maps.subtractBackground at the tile centre,
    filters.py:1190-1196), or None."""
    if params0.get('bckSub') == True and params0.get('bckSubScaleArcmin', 0) > 0:
        wcs0 = batch.wcsList[0]
        RAMin, RAMax, decMin, decMax = wcs0.getImageMinMaxWCSCoords()
        return maps.smoothScalesPix(wcs0, RADeg=(RAMax + RAMin) / 2., decDeg=(decMax + decMin) / 2.,
                                    smoothScaleDeg=params0['bckSubScaleArcmin'] / 60.)
    return None


def _templateSet(ops, templates, batch, params0, bck):
    """Per geometry, cached on the templates object: radians map and x pixel scale of the tile, the normalisation templates
    high-passed like the maps (device and host copies)."""
    hit = getattr(templates, '_realSpace', None)
    if hit is not None:
        return hit
    md0 = batch.mapDicts[0]
    probe = _filters.MapFilter.__new__(_filters.MapFilter)
    probe.wcs = md0[0]['wcs']
    probe.shape = (batch.H, batch.W)
    probe.makeRadiansMap()
    sig = templates.norm if templates.norm is not None else templates.unit           # [S, F, H, W]
    S, F, H, W = sig.shape
    hp = sig.reshape(S * F, H, W).contiguous()
    if bck is not None:
        hp = ops.gauss_filter(hp, bck[0], bck[1], subtract=True)
    hp = hp.reshape(S, F, H, W)
    hpHost = device.toHost(hp).astype(np.float64)
    hit = templates._realSpace = {'radiansMap': probe.radiansMap, 'degPerPixX': probe.degPerPixX, 'hp': hp, 'hpHost': hpHost,
                                  'crops': {}}
    return hit


TEMPLATE_WINDOW_RADIUS = 128


def _templateCrop(tset, ry, rx):
    """The central window of the high-passed templates that the normalisation pass convolves instead of the whole map, for
    kernels of half-sizes (ry, rx): (device crop [S, F, hc, wc], y0, x0, outsideMax [S, F]) or None when the tile is not larger
    than the window.  outsideMax[s, f] is the largest |template| among the pixels that outputs OUTSIDE the inner window
    (Chebyshev radius R around the centre) can read: if sum_f sum|kernel_f| outsideMax_f is below the maximum found inside, that
    maximum is the map's (a bound on every output outside the window)."""
    hit = tset['crops'].get((ry, rx))
    if hit is not None:
        return hit if hit != 'none' else None
    hp, hpHost = tset['hp'], tset['hpHost']
    S, F, H, W = hpHost.shape
    R = TEMPLATE_WINDOW_RADIUS
    cy, cx = H // 2, W // 2
    y0, y1, x0, x1 = cy - R - ry, cy + R + ry + 1, cx - R - rx, cx + R + rx + 1
    if y0 < 0 or x0 < 0 or y1 > H or x1 > W:
        tset['crops'][(ry, rx)] = 'none'
        return None
    yy, xx = np.ogrid[:H, :W]
    outside = (np.abs(yy - cy) > R - ry) | (np.abs(xx - cx) > R - rx)
    outsideMax = np.array([[np.abs(hpHost[s, f][outside]).max() for f in range(F)] for s in range(S)])
    hit = tset['crops'][(ry, rx)] = (hp[:, :, y0:y1, x0:x1].contiguous(), y0, x0, outsideMax)
    return hit


def buildKernels(ops, batch, filtersList, templates, fmaps):
    """Queues K2, K3 and the ky-mean of every (tile, filter) Fourier filter of a batch; returns the pending host copy of the
    means [nT S F, Wc]."""
    nT, F, H, W, S = batch.nT, batch.F, batch.H, batch.W, len(filtersList)
    Wc = W // 2 + 1
    # K2, K3: the nested Fourier matched filters (filters.py:568-630)
    spectra = getattr(templates, '_spectra', None)
    fsig = spectra[0] if spectra is not None else ops.rfft2(templates.unit.reshape(S * F, H, W), apodPix=0).reshape(S, F, H, Wc)
    cov = ops.xpower_smooth(fmaps, 3.0, W=W)
    filtH, _ = ops.build_filter(cov, fsig, templates.w, S, False, wantFull=False)
    del cov, fsig
    # ky-mean of every filter plane, then on the host the 1-d inverse transform along x (the filter is real and the kernel its
    # real part: the even extension of the half-plane mean)
    # (asynchronous copy: the caller queues the maps' high-pass behind it before it waits for the numbers)
    return device.HostCopy(ops.mean_over_rows(filtH.reshape(nT * S * F, H, Wc)))


def startRealSpaceTileBatch(ops, batch, filtersList, templates, parDict, undoPixelWindow=True, invertMap=False, rmsTables=False):
    """The real-space counterpart of pipelines.startTileBatch: returns the job pipelines.finishTileBatch takes."""
    dev = batch.dev if batch.dev is not None else batch.upload()
    nT, F, H, W, S = batch.nT, batch.F, batch.H, batch.W, len(filtersList)
    Wc = W // 2 + 1
    nOut = nT * S
    params0 = filtersList[0]['params']
    pixDeg = batch.wcsList[0].getPixelSizeDeg()
    bck = _bckSigmas(batch, params0)
    tset = _templateSet(ops, templates, batch, params0, bck)

    # 1, 2: kernels
    fmaps = ops.rfft2(dev['data'].reshape(nT * F, H, W), apodPix=20).reshape(nT, F, H, Wc)
    meanCopy = buildKernels(ops, batch, filtersList, templates, fmaps)
    del fmaps
    # the maps' Gaussian high-pass does not depend on the kernels: queued now, it runs while the host builds them
    cube = dev['data']
    if bck is not None:
        cube = ops.gauss_filter(cube.reshape(nT * F, H, W).contiguous(), bck[0], bck[1], subtract=True).reshape(nT, F, H, W)
    meanH = meanCopy.get().reshape(nOut, F, Wc)
    rowCut = fft.irfft(meanH, n=W, axis=-1)
    # (all images at once: the filters share kernelMaxArcmin and geometry, see canBatch)
    kernHost = np.ascontiguousarray(_filters.realSpaceKernelsFromRows(rowCut, W, tset['degPerPixX'], tset['radiansMap'], params0),
                                    dtype=np.float64)                             # [nOut, F, kh, kw]
    if kernHost.ndim != 4:
        raise Exception("real-space kernels of one batch came out with different footprints")
    kernT = ops.to_device(kernHost, ops.torch.float64)

    # 3: normalisation (filters.py:1029-1035, 1204-1212): filtered templates, signalNorm still 1.  The template sits at the tile
    # centre, so the kernels are run over a central window only; the maximum found there is accepted as the map's when a bound
    # on every output outside the window stays below it (_templateCrop), else the whole map is convolved
    numer = 2e-4 if templates.outputUnits == 'yc' else 1.0
    kh, kw = kernHost.shape[-2:]
    ry, rx = kh // 2, kw // 2
    idx = None
    crop = _templateCrop(tset, ry, rx)
    if crop is not None:
        hpCrop, cy0, cx0, outsideMax = crop
        wc = int(hpCrop.shape[-1])
        filteredSignal = ops.conv2d_batch(hpCrop, kernT, nOut, inMod=S)
        signalNorm, argmaxIdx, maxVal, _ = ops.template_norm(filteredSignal, 0, 0.0, 0.0, numer)
        del filteredSignal
        peak = device.toHost(maxVal)
        bound = (np.abs(kernHost).sum(axis=(2, 3)) * outsideMax[np.arange(nOut) % S]).sum(axis=1)
        if np.all(peak > bound):
            ic = device.toHost(argmaxIdx)
            idx = (ic // wc + cy0) * W + (ic % wc + cx0)
    if idx is None:
        filteredSignal = ops.conv2d_batch(tset['hp'], kernT, nOut, inMod=S)
        signalNorm, argmaxIdx, maxVal, _ = ops.template_norm(filteredSignal, 0, 0.0, 0.0, numer)
        del filteredSignal
        idx = device.toHost(argmaxIdx)
    fRel = np.empty((nOut, F), dtype=np.float64)
    yy = np.arange(kh)[:, None]
    xx = np.arange(kw)[None, :]
    for o in range(nOut):
        py, px = int(idx[o] // W), int(idx[o] % W)
        # ndimage.convolve with 'reflect': out[y, x] = sum kern[a, c] in[y + ry - a, x + rx - c]
        rows = _reflect(py + ry - yy, H)
        cols = _reflect(px + rx - xx, W)
        vals = (kernHost[o] * tset['hpHost'][o % S][:, rows, cols]).sum(axis=(1, 2))
        fRel[o] = vals / vals.sum()

    # 4: the maps
    filtered = ops.conv2d_batch(cube.contiguous(), kernT, nOut, norm=signalNorm, mask=dev['psMask'], inDiv=S, maskDiv=S)
    del cube
    grid = engine.noiseGrid((H, W), params0['noiseParams'], pixDeg)
    trimPix = engine.trimSizePix(params0, pixDeg, realSpace=True)
    data, sn, maskOut, rms, table, edges = engine.noiseSNAndMasks(ops, filtered, dev['wmed'], dev['psMask'], dev['surveyMask'],
                                                                  S, grid, trimPix, 20, True)
    rmsTabJob = None
    if rmsTables and table is not None:
        from . import completeness
        rmsTabJob = completeness.startRMSTables(ops, rms, maskOut, dev['rowArea'], S, table)
    if undoPixelWindow:
        engine.undoPixelWindow(ops, data)
    detectJob = photometry.startFindObjectsBatch(ops, sn, threshold=parDict['thresholdSigma'],
                                                 findCenterOfMass=parDict['findCenterOfMass'], removeRings=parDict['removeRings'],
                                                 ringThresholdSigma=parDict['ringThresholdSigma'], invertMap=invertMap,
                                                 measureShapes=parDict['measureShapes'])
    return {'ops': ops, 'batch': batch, 'filtersList': filtersList, 'parDict': parDict, 'mapUnits': templates.outputUnits,
            'invertMap': invertMap, 'detectJob': detectJob, 'data': data, 'SNMap': sn, 'surveyMask': maskOut, 'RMSMap': rms,
            'noiseTable': table, 'noiseEdges': edges, 'signalNorm': signalNorm, 'fRel': (fRel, None), 'filtH': None,
            'rmsTabJob': rmsTabJob, 'kern2d': kernHost}


def _reflect(i, n):
    """scipy.ndimage 'reflect' boundary (d c b a | a b c d | d c b a) for index arrays."""
    p = 2 * n
    i = np.mod(i, p)
    return np.where(i < n, i, p - 1 - i)
