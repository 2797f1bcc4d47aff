"""
Drop-in for the hot-path part of the reference's ``nemo.photometry`` (nemo/photometry.py:25-351 and :460-474):
``findObjects``, ``getObjectPositions``, ``getSNRValues``, ``measureFluxes``, ``deltaTToJyPerSr``.  Thresholding,
4-connected labelling (numbered exactly as scipy.ndimage.label), per-object moments and bicubic-spline
photometry run on the GPU; only per-object rows come back to the host.  There is no CPU path.

Maps may be numpy arrays, torch CUDA tensors or the :class:`~nemo_b200.device.DeviceArray` objects returned by
``nemo_b200.filters``; in the latter two cases nothing is uploaded.
"""

import numpy as np

from . import catalogs
from . import device
from . import maps
from . import signals
from . import wcs as _wcs
from .device import DeviceArray

MIN_RING_PIX = 30


def _ops(mapLike, ops=None):
    if ops is not None:
        return ops
    return device.getOps()


def _asImageBatch(mapData, ops):
    t = device.asTensor(mapData, ops)
    if t.dim() == 2:
        t = t.unsqueeze(0)
    return t.contiguous()


# -------------------------------------------------------------------------------------------------------------
def _segmentDense(ops, imgs, threshold, sign=1.0):
    """label + moments with the dense kernels, for images the sparse detector cannot hold."""
    labels, nLabels = ops.label4(imgs, threshold, sign=sign)
    n = device.toHost(nLabels)
    nmax = int(n.max()) if len(n) else 0
    table = device.toHost(ops.label_moments(imgs, labels, nmax, sign=sign))
    return labels, n, table


def _launchSegments(ops, imgs, thresholds, sign=1.0, wantArgmax=False):
    """Launches label + moments of a batch [B, H, W] at several thresholds; nothing is copied back yet."""
    return {'ops': ops, 'imgs': imgs, 'thresholds': list(thresholds), 'sign': sign,
            'launched': [ops.detect(imgs, thr, sign=sign, wantArgmax=wantArgmax) for thr in thresholds]}


def _collectSegments(job):
    """List of (labels tensor, nLabels host array, table host array), one per threshold of a _launchSegments job."""
    ops, imgs, launched = job['ops'], job['imgs'], job['launched']
    torch = ops.torch
    # one synchronising copy brings the label counts and overflow flags of every threshold
    info = device.toHost(torch.stack([torch.stack([nl, st]) for _, nl, st, _ in launched]))
    out = []
    for k, (labels, _, _, table) in enumerate(launched):
        n, status = info[k, 0], info[k, 1]
        nmax = int(n.max()) if len(n) else 0
        if status.any() or nmax > table.shape[1] - 1:
            out.append(_segmentDense(ops, imgs, job['thresholds'][k], sign=job['sign']))
            continue
        out.append((labels, n, device.toHost(table[:, :nmax + 1].contiguous())))
    return out


def _segmentMany(ops, imgs, thresholds, sign=1.0, wantArgmax=False):
    return _collectSegments(_launchSegments(ops, imgs, thresholds, sign=sign, wantArgmax=wantArgmax))


def _segment(ops, imgs, threshold, sign=1.0, wantArgmax=False):
    """label + moments for a batch [B, H, W].  Returns (labels tensor, nLabels host array, table host array)."""
    return _segmentMany(ops, imgs, [threshold], sign=sign, wantArgmax=wantArgmax)[0]


def _positionArrays(table, n, shape, findCenterOfMass=True):
    """(objIDs, ys, xs, objNumPix) arrays for one image's moment table, in the order of np.unique(segmentationMap)."""
    W = shape[1]
    hasBackground = table[0, 3] > 0
    ids = np.arange(0 if hasBackground else 1, n + 1)
    rows = table[ids]
    with np.errstate(divide='ignore', invalid='ignore'):
        if findCenterOfMass:
            ys = rows[:, 1] / rows[:, 0]
            xs = rows[:, 2] / rows[:, 0]
        else:
            ys = np.floor_divide(rows[:, 5], W)
            xs = np.mod(rows[:, 5], W)
    numPix = rows[:, 3].copy()
    if hasBackground:
        numPix[0] = 0.0         # ndimage.sum(sigPixMask, labels, 0): no significant pixel is in the background
    return ids, ys, xs, numPix


def _positionsFromTable(table, n, shape, findCenterOfMass=True):
    """(objIDs, objPositions, objNumPix) as getObjectPositions returns them, from one image's moment table."""
    ids, ys, xs, numPix = _positionArrays(table, n, shape, findCenterOfMass)
    if findCenterOfMass:
        positions = [(float(y), float(x)) for y, x in zip(ys, xs)]
    else:
        positions = [(int(y), int(x)) for y, x in zip(ys, xs)]
    return ids, positions, numPix


def getObjectPositions(mapData, threshold, findCenterOfMass=True, ops=None):
    """Creates a segmentation map and finds objects above the given threshold (photometry.py:193-222).

    Returns objIDs (including 0 when there is background, as np.unique does), objPositions as a list of (y, x),
    objNumPix, and the segmentation map (int32 numpy array).
    """
    if threshold < 0:
        raise Exception("Detection threshold (thresholdSigma in the config file) cannot be negative unless in forced photometry mode.")
    ops = _ops(mapData, ops)
    imgs = _asImageBatch(mapData, ops)
    labels, n, table = _segment(ops, imgs, threshold, wantArgmax=not findCenterOfMass)
    objIDs, objPositions, objNumPix = _positionsFromTable(table[0], int(n[0]), imgs.shape[1:], findCenterOfMass)
    return objIDs, objPositions, objNumPix, device.toHost(labels[0])


# -------------------------------------------------------------------------------------------------------------
def _gatherLabels(labels, bIdx, ys, xs):
    """Segmentation-map values at integer pixels of several images (one small gather, device -> host)."""
    if len(ys) == 0:
        return np.zeros(0, dtype=np.int32)
    import torch
    B, H, W = labels.shape
    ys = np.clip(np.nan_to_num(np.asarray(ys, dtype=np.float64)), 0, H - 1).astype(np.int64)
    xs = np.clip(np.nan_to_num(np.asarray(xs, dtype=np.float64)), 0, W - 1).astype(np.int64)
    flat = (np.asarray(bIdx, dtype=np.int64) * H + ys) * W + xs
    idx = device.countH2D(torch.from_numpy(flat)).to(labels.device)
    return device.toHost(labels.reshape(-1)[idx])


def startFindObjectsBatch(ops, snMaps, threshold=3.0, findCenterOfMass=True, removeRings=True, ringThresholdSigma=0,
                          invertMap=False, measureShapes=False):
    """Launches the device work of findObjects for a batch of S/N maps [B, H, W] (thresholding, labelling and moments at
    the detection and ring thresholds) without waiting for it; pass the result to finishFindObjectsBatch."""
    if threshold < 0 or (removeRings == True and ringThresholdSigma < 0):
        raise Exception("Detection threshold (thresholdSigma in the config file) cannot be negative unless in forced photometry mode.")
    sign = -1.0 if invertMap == True else 1.0
    dev = ops.find_objects(snMaps, threshold, ringThresholdSigma if removeRings == True else None, sign=sign,
                           minRingPix=MIN_RING_PIX, findCenterOfMass=findCenterOfMass, wantLabels=measureShapes == True)
    shapes = None
    if measureShapes == True:
        # second-order moments of every label for the SExtractor-style ellipses (photometry.py:126-178), queued behind the labelling
        shapes = ops.label_shapes(snMaps, dev[4], dev[2].shape[1] - 1, sign=sign)
    return {'shapes': shapes, 'ops': ops, 'imgs': snMaps, 'sign': sign, 'thresholds': [threshold, ringThresholdSigma] if removeRings == True else [threshold],
            'threshold': threshold, 'findCenterOfMass': findCenterOfMass, 'removeRings': removeRings, 'dev': dev}


def _collectDetections(job):
    """(nLabels host, moment table host [B, nmax + 1, 8], ring flags host [B, nmax + 1] or None, shape moments host or None) of a
    startFindObjectsBatch job.
    One synchronising copy of the counts, then two small ones; nothing at the ring threshold leaves the device.  Batches
    that overflowed the sparse detector are redone with the dense kernels."""
    ops, imgs, sign = job['ops'], job['imgs'], job['sign']
    torch = ops.torch
    nLabels, status, table, flags, _ = job['dev']
    shapes = job.get('shapes')
    info = device.toHost(torch.stack([nLabels, status]))
    n = info[0]
    if info[1].any() or (len(n) and int(n.max()) > table.shape[1] - 1):
        dense = []
        for thr in job['thresholds']:
            lab, nl = ops.label4(imgs, thr, sign=sign)
            nmaxK = int(device.toHost(nl).max())
            dense.append((lab, nl, ops.label_moments(imgs, lab, nmaxK, sign=sign)))
        _, nlDev, table = dense[0]
        n = device.toHost(nlDev)
        flags = None
        if job['removeRings'] == True:
            flags = ops.ring_flags(table, nlDev, dense[1][0], dense[1][2], minRingPix=MIN_RING_PIX, findCenterOfMass=job['findCenterOfMass'])
        if shapes is not None:
            shapes = ops.label_shapes(imgs, dense[0][0], table.shape[1] - 1, sign=sign)
    nmax = int(n.max()) if len(n) else 0
    tableHost = device.toHost(table[:, :nmax + 1].contiguous())
    flagsHost = device.toHost(flags[:, :nmax + 1].contiguous()) if flags is not None else None
    shapesHost = device.toHost(shapes[:, :nmax + 1].contiguous()) if shapes is not None else None
    return n, tableHost, flagsHost, shapesHost


ELLIPSE_KEYS = ('ellipse_PA', 'ellipse_A', 'ellipse_B', 'ellipse_x0', 'ellipse_y0', 'ellipse_e')


def ellipsesFromMoments(rows, numSigPix):
    """The six ellipse columns of findObjects(measureShapes = True) from the label moments of nb_label_shapes (rows [K, 8]:
    sum d, sum xs d, sum ys d, sum xs^2 d, sum ys^2 d, sum xs ys d, xMin, yMin) -- the reference's operations, in its order
    (photometry.py:139-178): -99 for objects of at most 9 pixels and where its sign test on theta and xy fails."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 8)
    numSigPix = np.asarray(numSigPix, dtype=np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        s0 = rows[:, 0]
        cx2 = rows[:, 1] / s0
        cy2 = rows[:, 2] / s0
        x2 = (rows[:, 3] / s0) - cx2 ** 2
        y2 = (rows[:, 4] / s0) - cy2 ** 2
        xy = (rows[:, 5] / s0) - cx2 * cy2
        theta = np.degrees(np.arctan(2 * (xy / (x2 - y2))) / 2.0)
        up = (xy > 0) & (theta < 0)
        down = ~up & (xy < 0) & (theta > 0)
        theta = np.where(up, theta + 90, np.where(down, theta - 90, theta))
        ok = (numSigPix > 9) & (((theta > 0) & (xy > 0)) | ((theta < 0) & (xy < 0)))
        root = np.sqrt(((x2 - y2) / 2) ** 2 + xy ** 2)
        A = np.sqrt((x2 + y2) / 2.0 + root)
        B = np.sqrt((x2 + y2) / 2.0 - root)
        scaleFactor = np.sqrt(numSigPix / (A * B * np.pi))
        A = A * scaleFactor
        B = B * scaleFactor
        ecc = np.sqrt(1 - B ** 2 / A ** 2)
    vals = (theta, A, B, cx2 + rows[:, 6], cy2 + rows[:, 7], ecc)
    return {k: np.where(ok, v, -99.0) for k, v in zip(ELLIPSE_KEYS, vals)}


def _flatObjects(table, n, shape, findCenterOfMass=True, minPix=None):
    """All labels of all images of a batch as flat arrays in np.unique order, image-major: (image index, label id, y, x,
    numPix).  Label 0 is listed for images that have background pixels, with numPix = 0, as ndimage.sum gives.  With
    minPix only labels of more than minPix pixels are listed."""
    W = shape[1]
    B, K1 = table.shape[0], table.shape[1]
    valid = np.arange(K1)[None, :] <= np.asarray(n)[:, None]
    valid[:, 0] = table[:, 0, 3] > 0
    if minPix is not None:
        numAll = table[:, :, 3].copy()
        numAll[:, 0] = 0.0
        valid &= numAll > minPix
    bIdx, ids = np.nonzero(valid)
    rows = table[bIdx, ids]
    with np.errstate(divide='ignore', invalid='ignore'):
        if findCenterOfMass:
            ys = rows[:, 1] / rows[:, 0]
            xs = rows[:, 2] / rows[:, 0]
        else:
            ys = np.floor_divide(rows[:, 5], W)
            xs = np.mod(rows[:, 5], W)
    numPix = np.where(ids == 0, 0.0, rows[:, 3])
    return bIdx, ids, ys, xs, numPix


def finishFindObjectsBatch(job, wcsList, flagMaskList, labelList, tileNameList, minObjPix=3, objIdent='ACT-CL', longNames=False,
                           useInterpolator=True, extraSamples=None):
    """Host part of findObjects for a job made by startFindObjectsBatch: positions from the moment tables, ring removal,
    S/N at the positions, sky coordinates and flags.  Returns one flat, image-major set of columns for the whole batch
    (see splitObjects).  extraSamples: list of (key, images [B', H, W], fn(imageIndex array) -> index into images, sign);
    each is sampled at the positions recomputed from (RADeg, decDeg) in the same launch round as the S/N and returned in
    flat['samples'][key] (what measureFluxes / getSNRValues read, photometry.py:243-253, 295-333)."""
    ops, snMaps, sign = job['ops'], job['imgs'], job['sign']
    threshold, findCenterOfMass, removeRings = job['threshold'], job['findCenterOfMass'], job['removeRings']
    B, H, W = snMaps.shape
    n, table, ringFlags, shapeTable = _collectDetections(job)
    cb, cid, cy, cx, cn = _flatObjects(table, n, (H, W), findCenterOfMass, minPix=minObjPix)
    keep = cn > minObjPix
    if ringFlags is not None:
        # rings: > 30 pixels at ringThresholdSigma whose centroid pixel is not one of their own (photometry.py:60-70)
        keep &= ringFlags[cb, cid] == 0
    ellipses = ellipsesFromMoments(shapeTable[cb[keep], cid[keep]], cn[keep]) if shapeTable is not None else None
    cb, cy, cx, cn = cb[keep].astype(np.int32), cy[keep], cx[keep], cn[keep]
    K = len(cb)
    flat = {'counts': np.zeros(B, dtype=np.int64), 'labels': list(labelList), 'tiles': list(tileNameList), 'objIdent': objIdent,
            'longNames': longNames, 'cols': None, 'samples': {}}
    if K == 0:
        return flat
    # sky coordinates and flags of every candidate; positions recomputed from them for the flux measurements, as the
    # reference's measureFluxes does (photometry.py:328)
    perImage = np.bincount(cb, minlength=B)
    ra, dec = np.empty(K), np.empty(K)
    xs2, ys2 = np.empty(K), np.empty(K)
    flags = np.zeros(K, dtype=int)
    lo = 0
    wcsParams = _wcs.batchParams(wcsList)
    if wcsParams is not None:      # every image on a CARWCS: all positions in one expression
        ra, dec = _wcs.batchPix2wcs(wcsParams, cb, cx, cy)
        ra = np.where(ra < 0, 360.0 + ra, ra)
        if extraSamples:
            xs2, ys2 = _wcs.batchWcs2pix(wcsParams, cb, ra, dec)
    for b in np.nonzero(perImage)[0]:
        sl = slice(lo, lo + perImage[b])
        lo += perImage[b]
        if wcsParams is None:
            r, d = wcsList[b].pix2wcs(cx[sl], cy[sl])
            r = np.where(np.asarray(r) < 0, 360.0 + np.asarray(r), r)
            ra[sl], dec[sl] = r, d
            if extraSamples:
                xs2[sl], ys2[sl] = wcsList[b].wcs2pix(ra[sl], dec[sl])
        if flagMaskList[b] is not None:
            flags[sl] = np.asarray(flagMaskList[b])[np.round(cy[sl]).astype(int), np.round(cx[sl]).astype(int)]
    # S/N at the object positions and the extra samples, all launched before the one copy back
    samples = {}
    if useInterpolator == True:
        torch = ops.torch
        pos = np.stack([np.stack([cy, cx], axis=1), np.stack([ys2, xs2], axis=1)]) if extraSamples else np.stack([cy, cx], axis=1)[None]
        posT = ops.to_device(pos, torch.float64)
        idxHost = np.stack([cb] + [np.asarray(fn(cb), dtype=np.int32) for _, _, fn, _ in extraSamples]) if extraSamples else cb[None]
        idxT = ops.to_device(np.ascontiguousarray(idxHost, dtype=np.int32), torch.int32)
        outs = [ops.spline_eval(snMaps, idxT[0], posT[0], sign=sign)]
        for k, (key, imgs, fn, sgn) in enumerate(extraSamples or []):
            outs.append(ops.spline_eval(imgs, idxT[k + 1], posT[1], sign=sgn))
        vals = device.toHost(torch.stack(outs))
        snr = vals[0]
        for k, (key, imgs, fn, sgn) in enumerate(extraSamples or []):
            samples[key] = vals[k + 1]
    else:
        snr = sign * _gatherValues(snMaps, cb, np.round(cy).astype(np.int64), np.round(cx).astype(np.int64))
        for key, imgs, fn, sgn in (extraSamples or []):
            samples[key] = sgn * _gatherValues(imgs, np.asarray(fn(cb), dtype=np.int64), np.round(ys2).astype(np.int64), np.round(xs2).astype(np.int64))
    # id counts every object that passed the size and ring cuts, kept or not (photometry.py:182)
    ids = np.arange(K) - (np.cumsum(perImage) - perImage)[cb] + 1
    sel = snr > threshold
    cb = cb[sel]
    flat['counts'] = np.bincount(cb, minlength=B)
    flat['owner'] = cb
    ra, dec = ra[sel], dec[sel]
    flat['cols'] = {'id': ids[sel], 'x': cx[sel], 'y': cy[sel], 'RADeg': ra, 'decDeg': dec, 'galacticLatDeg': _galacticLat(ra, dec),
                    'numSigPix': cn[sel], 'SNR': snr[sel].astype(np.float64), 'flags': flags[sel]}
    if ellipses is not None:
        flat['cols'].update({k: v[sel] for k, v in ellipses.items()})
    flat['samples'] = {k: v[sel] for k, v in samples.items()}
    flat['xy2'] = (xs2[sel], ys2[sel]) if extraSamples else None
    return flat


def columnsToTable(cols, extra=None, keysToWrite=None):
    """Table from one image's columns: dict with x, y, RADeg, decDeg, SNR, ... arrays and the scalars template, tileName,
    objIdent, longNames (None -> [])."""
    if cols is None:
        return []
    flat = {'counts': np.array([len(cols['x'])]), 'labels': [cols['template']], 'tiles': [cols['tileName']],
            'objIdent': cols['objIdent'], 'longNames': cols['longNames'],
            'cols': {k: np.asarray(v) for k, v in cols.items() if k not in ('template', 'tileName', 'objIdent', 'longNames')}}
    return splitObjects(flat, extra, keysToWrite)[0]


def splitObjects(flat, extra=None, keysToWrite=None, stacked=None):
    """Per-image tables (``[]`` where nothing was found) from the flat columns of finishFindObjectsBatch, plus optional
    extra flat columns (fluxes).  The tables are row ranges of the batch's one stacked table (`stacked`: flatTable of the same
    arguments, made here when not given) that are cut out when first read (catalogs.RowsOf); object names are formatted the
    first time a 'name' column is read."""
    counts = flat['counts']
    if stacked is None:
        stacked = flatTable(flat, extra, keysToWrite)
    out = []
    lo = 0
    for cnt in counts:
        if cnt == 0:
            out.append([])
            continue
        out.append(catalogs.RowsOf(stacked, lo, lo + cnt))
        lo += cnt
    return out


def flatTable(flat, extra=None, keysToWrite=None):
    """All rows of a batch as ONE table, in the order splitObjects' tables stack (image by image): what
    catalogs.makeOptimalCatalog stacks anyway, without the hundreds of small tables in between.  Columns are the flat arrays
    themselves (no copies); None when nothing was found."""
    counts = np.asarray(flat['counts'])
    if flat['cols'] is None or int(counts.sum()) == 0:
        return None
    keys = keysToWrite if keysToWrite is not None else catalogs.COLUMN_NAMES
    cols = dict(flat['cols'])
    if extra:
        cols.update(extra)
    ident, longNames = flat['objIdent'], flat['longNames']
    ordered = []
    for key in keys:
        if key == 'name':
            ra, dec = cols['RADeg'], cols['decDeg']
            ordered.append((key, catalogs.deferred(lambda ra=ra, dec=dec: catalogs.makeNames(ra, dec, prefix=ident, longNames=longNames))))
        elif key == 'template':
            ordered.append((key, np.repeat(np.array(flat['labels'], dtype=object), counts)))
        elif key == 'tileName':
            ordered.append((key, np.repeat(np.array(flat['tiles'], dtype=object), counts)))
        elif key in cols:
            ordered.append((key, cols[key]))
    return catalogs.Table.fromColumns(ordered)


def findObjectsBatch(ops, snMaps, wcsList, flagMaskList, labelList, tileNameList, threshold=3.0, minObjPix=3,
                     rejectBorder=10, findCenterOfMass=True, removeRings=True, ringThresholdSigma=0, invertMap=False,
                     objIdent='ACT-CL', longNames=False, useInterpolator=True, measureShapes=False):
    """findObjects for a batch of S/N maps [B, H, W] on the device; returns a list of B tables ([] where nothing was found)."""
    job = startFindObjectsBatch(ops, snMaps, threshold=threshold, findCenterOfMass=findCenterOfMass, removeRings=removeRings,
                                ringThresholdSigma=ringThresholdSigma, invertMap=invertMap, measureShapes=measureShapes)
    return splitObjects(finishFindObjectsBatch(job, wcsList, flagMaskList, labelList, tileNameList, minObjPix=minObjPix,
                                               objIdent=objIdent, longNames=longNames, useInterpolator=useInterpolator))


def _gatherValues(imgs, bIdx, ys, xs):
    import torch
    B, H, W = imgs.shape
    flat = (np.asarray(bIdx, dtype=np.int64) * H + ys) * W + xs
    return device.toHost(imgs.reshape(-1)[device.countH2D(torch.from_numpy(flat)).to(imgs.device)]).astype(np.float64)


def _galacticLat(ra, dec):
    a, d = np.radians(ra), np.radians(dec)
    v = np.stack([np.cos(d) * np.cos(a), np.cos(d) * np.sin(a), np.sin(d)])
    g = _wcs._EQ2GAL @ v
    return np.degrees(np.arcsin(np.clip(g[2], -1, 1)))


def findObjects(filteredMapDict, threshold=3.0, minObjPix=3, rejectBorder=10, findCenterOfMass=True,
                removeRings=True, ringThresholdSigma=0, invertMap=False, objIdent='ACT-CL', longNames=False,
                verbose=True, useInterpolator=True, measureShapes=False, DS9RegionsPath=None):
    """Finds objects in the S/N map of a filtered map dictionary (photometry.py:25-190).  Returns a table
    (``catalogs.Table``, or ``[]`` when nothing is found, as the reference does)."""
    ops = _ops(None)
    sn = _asImageBatch(filteredMapDict['SNMap'], ops)
    cats = findObjectsBatch(ops, sn, [filteredMapDict['wcs']], [filteredMapDict.get('flagMask')],
                            [filteredMapDict['label']], [filteredMapDict['tileName']], threshold=threshold,
                            minObjPix=minObjPix, rejectBorder=rejectBorder, findCenterOfMass=findCenterOfMass,
                            removeRings=removeRings, ringThresholdSigma=ringThresholdSigma, invertMap=invertMap,
                            objIdent=objIdent, longNames=longNames, useInterpolator=useInterpolator,
                            measureShapes=measureShapes)
    return cats[0]


# -------------------------------------------------------------------------------------------------------------
def _valuesAt(ops, mapLike, xs, ys, useInterpolator, sign=1.0):
    imgs = _asImageBatch(mapLike, ops)
    if len(xs) == 0:
        return np.zeros(0)
    if useInterpolator == True:
        yx = np.stack([np.asarray(ys, dtype=np.float64), np.asarray(xs, dtype=np.float64)], axis=1)
        return device.toHost(ops.spline_eval(imgs, None, ops.to_device(yx, ops.torch.float64), sign=sign))
    torch = ops.torch
    yi = torch.as_tensor(np.round(np.asarray(ys)).astype(np.int64), device=imgs.device)
    xi = torch.as_tensor(np.round(np.asarray(xs)).astype(np.int64), device=imgs.device)
    return sign * device.toHost(imgs[0][yi, xi]).astype(np.float64)


def getSNRValues(catalog, SNMap, wcs, useInterpolator=True, invertMap=False, prefix=''):
    """Adds <prefix>SNR measured in SNMap at the catalog positions (photometry.py:225-255)."""
    ops = _ops(None)
    if len(catalog) == 0:
        return
    shape = SNMap.shape
    catalog.add_column(catalogs.Column(np.zeros(len(catalog)), prefix + 'SNR'))
    xs, ys = wcs.wcs2pix(np.asarray(catalog['RADeg'], dtype=float), np.asarray(catalog['decDeg'], dtype=float))
    xs, ys = np.atleast_1d(xs), np.atleast_1d(ys)
    inside = np.array([int(x) > 0 and int(x) < shape[1] and int(y) > 0 and int(y) < shape[0] for x, y in zip(xs, ys)], dtype=bool)
    vals = _valuesAt(ops, SNMap, xs[inside], ys[inside], useInterpolator, sign=-1.0 if invertMap == True else 1.0)
    col = catalog[prefix + 'SNR']
    col[np.where(inside)[0]] = vals


def measureFluxes(catalog, filteredMapDict, diagnosticsDir, photFilteredMapDict=None, useInterpolator=True,
                  ycObsFreqGHz=148.0):
    """Adds flux measurements at the catalog positions (photometry.py:258-351)."""
    ops = _ops(None)
    mapData = filteredMapDict['data']
    wcs = filteredMapDict['wcs']
    mapUnits = wcs.header['BUNIT']
    if photFilteredMapDict is not None:
        getSNRValues(catalog, photFilteredMapDict['SNMap'], wcs, prefix='fixed_', useInterpolator=useInterpolator)
    if 'BEAMNSR' in wcs.header.keys():
        beamSolidAngle_nsr = wcs.header['BEAMNSR']
        obsFreqGHz = wcs.header['FREQGHZ']
        reportJyFluxes = True
    else:
        reportJyFluxes = False
    mapDataList = [mapData]
    prefixList = ['']
    if photFilteredMapDict is not None:
        mapDataList.append(photFilteredMapDict['data'])
        prefixList.append('fixed_')
    keysToAdd = ['deltaT_c', 'err_deltaT_c']
    if mapUnits == 'yc':
        keysToAdd = keysToAdd + ['y_c', 'err_y_c']
    elif mapUnits == 'uK':
        keysToAdd = keysToAdd + ['fluxJy', 'err_fluxJy']
    if len(catalog) == 0:
        return
    for prefix in prefixList:
        for k in keysToAdd:
            catalog.add_column(catalogs.Column(np.zeros(len(catalog)), prefix + k))
    xs, ys = wcs.wcs2pix(np.asarray(catalog['RADeg'], dtype=float), np.asarray(catalog['decDeg'], dtype=float))
    xs, ys = np.atleast_1d(xs), np.atleast_1d(ys)
    for data, prefix in zip(mapDataList, prefixList):
        mapValue = _valuesAt(ops, data, xs, ys, useInterpolator)
        snr = np.asarray(catalog[prefix + 'SNR'], dtype=float)
        with np.errstate(divide='ignore', invalid='ignore'):
            if mapUnits == 'yc':
                yc = mapValue
                catalog[prefix + 'y_c'] = yc / 1e-4
                catalog[prefix + 'err_y_c'] = catalog[prefix + 'y_c'] / snr
                deltaTc = maps.convertToDeltaT(yc, obsFrequencyGHz=ycObsFreqGHz)
                catalog[prefix + 'deltaT_c'] = deltaTc
                catalog[prefix + 'err_deltaT_c'] = abs(deltaTc / snr)
            elif mapUnits == 'uK':
                deltaTc = mapValue
                catalog[prefix + 'deltaT_c'] = deltaTc
                catalog[prefix + 'err_deltaT_c'] = deltaTc / snr
                if reportJyFluxes == True:
                    catalog[prefix + "fluxJy"] = deltaTToJyPerSr(catalog[prefix + 'deltaT_c'], obsFreqGHz) * beamSolidAngle_nsr * 1.e-9
                    catalog[prefix + "err_fluxJy"] = deltaTToJyPerSr(catalog[prefix + 'err_deltaT_c'], obsFreqGHz) * beamSolidAngle_nsr * 1.e-9


def deltaTToJyPerSr(temp, obsFreqGHz):
    """Delta T (uK) to Jy/sr at the given frequency in GHz (photometry.py:460-474)."""
    kB = 1.380649e-16
    h = 6.62607015e-27
    c = 29979245800.0
    nu = obsFreqGHz * 1.e9
    T0 = signals.TCMB
    x = h * nu / (kB * T0)
    cNu = 2 * (kB * T0) ** 3 / (h ** 2 * c ** 2) * x ** 4 / (4 * (np.sinh(x / 2.)) ** 2)
    cNu *= 1e23
    return temp * cNu * 1e-6 / T0


def JyPerSrToDeltaT(JySr, obsFreqGHz):
    return JySr / deltaTToJyPerSr(1.0, obsFreqGHz)
