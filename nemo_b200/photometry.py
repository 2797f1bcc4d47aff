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
def _segment(ops, imgs, threshold, sign=1.0, wantArgmax=False):
    """label + moments for a batch [B, H, W].  Returns (labels tensor, nLabels host array, table host array)."""
    labels, nLabels = ops.label4(imgs, threshold, sign=sign)
    n = device.toHost(nLabels)
    nmax = int(n.max()) if len(n) else 0
    table = device.toHost(ops.label_moments(imgs, labels, nmax, sign=sign))
    return labels, n, table


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
    labels, n, table = _segment(ops, imgs, threshold)
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


def findObjectsBatch(ops, snMaps, wcsList, flagMaskList, labelList, tileNameList, threshold=3.0, minObjPix=3,
                     rejectBorder=10, findCenterOfMass=True, removeRings=True, ringThresholdSigma=0, invertMap=False,
                     objIdent='ACT-CL', longNames=False, useInterpolator=True, measureShapes=False):
    """findObjects for a batch of S/N maps [B, H, W] on the device; returns a list of B catalogs (lists of dicts).
    Host synchronisations per batch: label counts, moment tables, two label gathers (rings), spline values."""
    if measureShapes == True:
        raise Exception("measureShapes is not implemented in nemo_b200")
    if threshold < 0:
        raise Exception("Detection threshold (thresholdSigma in the config file) cannot be negative unless in forced photometry mode.")
    sign = -1.0 if invertMap == True else 1.0
    B, H, W = snMaps.shape
    labels, n, table = _segment(ops, snMaps, threshold, sign=sign)
    objs = [_positionArrays(table[b], int(n[b]), (H, W), findCenterOfMass) for b in range(B)]
    keepMasks = [o[3] > minObjPix for o in objs]
    if removeRings == True:
        if ringThresholdSigma < 0:
            raise Exception("Detection threshold (thresholdSigma in the config file) cannot be negative unless in forced photometry mode.")
        # rings: > 30 pixels at ringThresholdSigma whose centroid pixel is not one of their own (photometry.py:60-70)
        minRingPix = 30
        ringLabels, ringN, ringTable = _segment(ops, snMaps, ringThresholdSigma, sign=sign)
        rb, ry, rx, rid = [], [], [], []
        for b in range(B):
            rIDs, rys, rxs, rNumPix = _positionArrays(ringTable[b], int(ringN[b]), (H, W), True)
            big = rNumPix > minRingPix
            rb.append(np.full(int(big.sum()), b, dtype=np.int64))
            ry.append(np.trunc(rys[big]))
            rx.append(np.trunc(rxs[big]))
            rid.append(rIDs[big])
        rb, ry, rx, rid = np.concatenate(rb), np.concatenate(ry), np.concatenate(rx), np.concatenate(rid)
        at = _gatherLabels(ringLabels, rb, ry, rx)
        bad = at != rid
        ringBad = [set(rid[bad & (rb == b)].tolist()) for b in range(B)]
        if any(len(r) > 0 for r in ringBad):
            cb = np.concatenate([np.full(int(k.sum()), b, dtype=np.int64) for b, k in enumerate(keepMasks)])
            cy = np.concatenate([np.trunc(objs[b][1][keepMasks[b]]) for b in range(B)])
            cx = np.concatenate([np.trunc(objs[b][2][keepMasks[b]]) for b in range(B)])
            at = _gatherLabels(ringLabels, cb, cy, cx)
            pos = 0
            for b in range(B):
                k = np.where(keepMasks[b])[0]
                if len(ringBad[b]) > 0:
                    inRing = np.array([int(v) in ringBad[b] for v in at[pos:pos + len(k)]], dtype=bool)
                    keepMasks[b][k[inRing]] = False
                pos += len(k)

    cb = np.concatenate([np.full(int(k.sum()), b, dtype=np.int32) for b, k in enumerate(keepMasks)]) if B > 0 else np.zeros(0, np.int32)
    cy = np.concatenate([objs[b][1][keepMasks[b]] for b in range(B)]) if B > 0 else np.zeros(0)
    cx = np.concatenate([objs[b][2][keepMasks[b]] for b in range(B)]) if B > 0 else np.zeros(0)
    cn = np.concatenate([objs[b][3][keepMasks[b]] for b in range(B)]) if B > 0 else np.zeros(0)
    K = len(cb)
    catalogList = [None for _ in range(B)]
    if K == 0:
        return catalogList
    # S/N at the object positions
    yx = np.stack([cy, cx], axis=1).astype(np.float64)
    if useInterpolator == True:
        snr = device.toHost(ops.spline_eval(snMaps, ops.to_device(cb, ops.torch.int32), ops.to_device(yx, ops.torch.float64), sign=sign))
    else:
        snr = sign * _gatherValues(snMaps, cb, np.round(cy).astype(np.int64), np.round(cx).astype(np.int64))
    pos = 0
    for b in range(B):
        nb = int(keepMasks[b].sum())
        if nb == 0:
            continue
        sl = slice(pos, pos + nb)
        pos += nb
        keep = snr[sl] > threshold
        if not keep.any():
            continue
        wcs = wcsList[b]
        ys, xs = cy[sl][keep], cx[sl][keep]
        ra, dec = wcs.pix2wcs(xs, ys)
        ra, dec = np.atleast_1d(np.array(ra, dtype=np.float64)), np.atleast_1d(np.array(dec, dtype=np.float64))
        ra = np.where(ra < 0, 360.0 + ra, ra)
        flagMask = flagMaskList[b]
        if flagMask is not None:
            flags = np.asarray(flagMask)[np.round(ys).astype(int), np.round(xs).astype(int)]
        else:
            flags = np.zeros(len(ys), dtype=int)
        # id counts every object that passed the size and ring cuts, kept or not (photometry.py:182)
        catalogList[b] = {'id': (np.arange(nb) + 1)[keep], 'x': xs, 'y': ys, 'RADeg': ra, 'decDeg': dec,
                          'galacticLatDeg': _galacticLat(ra, dec), 'numSigPix': cn[sl][keep], 'SNR': snr[sl][keep].astype(np.float64),
                          'flags': flags, 'template': labelList[b], 'tileName': tileNameList[b],
                          'objIdent': objIdent, 'longNames': longNames}
    return catalogList


def columnsToTable(cols, extra=None, keysToWrite=None):
    """Table from one image's column dict as made by findObjectsBatch (None -> []).  The object names are formatted
    the first time the 'name' column is read."""
    if cols is None:
        return []
    n = len(cols['x'])
    tab = catalogs.Table()
    ra, dec, ident, longNames = cols['RADeg'], cols['decDeg'], cols['objIdent'], cols['longNames']
    values = dict(cols)
    if extra is not None:
        values.update(extra)
    for key in (keysToWrite if keysToWrite is not None else catalogs.COLUMN_NAMES):
        if key == 'name':
            tab.add_deferred_column('name', lambda: catalogs.makeNames(ra, dec, prefix=ident, longNames=longNames))
        elif key in ('template', 'tileName'):
            tab.add_column(catalogs.Column(np.full(n, values[key], dtype=object), key))
        elif key in values and key not in ('objIdent', 'longNames'):
            tab.add_column(catalogs.Column(values[key], key))
    return tab


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
    return columnsToTable(cats[0])


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
