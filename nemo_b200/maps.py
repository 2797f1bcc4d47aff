"""
Map helpers that sit on the hot path (reference nemo/maps.py): unit conversions (:660-688) are host scalars;
``smoothMap`` / ``subtractBackground`` (:1355-1368, :1433-1458) run on the GPU through nb_gauss_filter.
"""

import numpy as np
from scipy import interpolate

from . import signals
from . import wcs as _wcs


def convertToY(mapData, obsFrequencyGHz=148):
    fx = signals.fSZ(obsFrequencyGHz)
    return (mapData / (signals.TCMB * 1e6)) / fx


def convertToDeltaT(mapData, obsFrequencyGHz=148, TCMBAlpha=0.0, z=None):
    fx = signals.fSZ(obsFrequencyGHz, TCMBAlpha=TCMBAlpha, z=z)
    return mapData * fx * (signals.TCMB * 1e6)


def smoothScalesPix(wcs, RADeg='centre', decDeg='centre', smoothScaleDeg=5.0 / 60.0):
    """(sigmaY, sigmaX) in pixels of the Gaussian maps.smoothMap applies."""
    xPixScale, yPixScale = _wcs.pixelScalesDeg(wcs, RADeg, decDeg)
    return smoothScaleDeg / yPixScale, smoothScaleDeg / xPixScale


def smoothMap(data, wcs, RADeg='centre', decDeg='centre', smoothScaleDeg=5.0 / 60.0, ops=None):
    """Gaussian smoothing on the GPU.  `data` may be a numpy array (returned as numpy) or a device tensor."""
    from . import device
    return device.gaussFilter(data, *smoothScalesPix(wcs, RADeg, decDeg, smoothScaleDeg), subtract=False, ops=ops)


def subtractBackground(data, wcs, RADeg='centre', decDeg='centre', smoothScaleDeg=30.0 / 60.0, ops=None):
    from . import device
    return device.gaussFilter(data, *smoothScalesPix(wcs, RADeg, decDeg, smoothScaleDeg), subtract=True, ops=ops)


# ------------------------------------------------------------------------------------------------------
# Model images (reference nemo/maps.py:1734-1899 makeModelImage, :2414-2471 makeDegreesDistanceMap): painted on the GPU
def objectBoxes(wcs, RADeg, decDeg, maxDistDegrees, shape):
    """What maps.makeDegreesDistanceMap derives for each object before it fills its box (maps.py:2443-2466), for arrays of
    objects: pixel position, local pixel scales, and the box [minX, maxX) x [minY, maxY) clipped to the map.  Returns an
    [n, 8] array: x0, y0, xPixScale, yPixScale, minX, maxX, minY, maxY."""
    RADeg, decDeg = np.atleast_1d(np.asarray(RADeg, dtype=np.float64)), np.atleast_1d(np.asarray(decDeg, dtype=np.float64))
    x0, y0 = wcs.wcs2pix(RADeg, decDeg)
    x0, y0 = np.asarray(x0, dtype=np.float64), np.asarray(y0, dtype=np.float64)
    ra1, dec1 = wcs.pix2wcs(x0 + 1, y0 + 1)
    xPixScale = _wcs.calcAngSepDeg(RADeg, decDeg, np.asarray(ra1), decDeg)
    yPixScale = _wcs.calcAngSepDeg(RADeg, decDeg, RADeg, np.asarray(dec1))
    maxDist = np.broadcast_to(np.asarray(maxDistDegrees, dtype=np.float64), x0.shape)
    xDistPix = np.rint(maxDist / xPixScale)
    yDistPix = np.rint(maxDist / yPixScale)
    Y, X = shape
    out = np.empty((len(x0), 8), dtype=np.float64)
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = x0, y0, xPixScale, yPixScale
    out[:, 4] = np.maximum(np.rint(x0) - xDistPix, 0)
    out[:, 5] = np.minimum(np.rint(x0) + xDistPix, X)
    out[:, 6] = np.maximum(np.rint(y0) - yDistPix, 0)
    out[:, 7] = np.minimum(np.rint(y0) + yDistPix, Y)
    return out


def _catalogColumns(catalog):
    keys = list(catalog.keys())
    return {k: np.asarray(catalog[k]) for k in keys}


def selectModelObjects(catalog, shape, wcs, validAreaSection=None, minSNR=-99):
    """The rows of `catalog` makeModelImage paints: inside the image (catalogs.getCatalogWithinImage), above the S/N cut when
    the catalog has an 'SNR' or 'fixed_SNR' column, inside validAreaSection = [xMin, xMax, yMin, yMax] when given
    (maps.py:1783-1810).  Returns a dict of column arrays."""
    from . import catalogs
    cols = _catalogColumns(catalog)
    sel = catalogs.withinImage(cols['RADeg'], cols['decDeg'], shape, wcs)
    SNRKey = 'SNR' if 'SNR' in cols else ('fixed_SNR' if 'fixed_SNR' in cols else None)
    if SNRKey is not None:
        sel &= cols[SNRKey].astype(np.float64) > minSNR
    cols = {k: v[sel] for k, v in cols.items()}
    if validAreaSection is not None and len(cols['RADeg']) > 0:
        x0, x1, y0, y1 = validAreaSection
        x, y = wcs.wcs2pix(cols['RADeg'].astype(np.float64), cols['decDeg'].astype(np.float64))
        x, y = np.atleast_1d(x), np.atleast_1d(y)
        cMask = (x >= x0) & (x < x1) & (y >= y0) & (y < y1)
        cols = {k: v[cMask] for k, v in cols.items()}
    return cols


_beamTransformCache = {}


def _beamTimesWindow(ops, shape, wcs, beam, applyPixelWindow):
    """Device [H, Wc] real multiplier: transform of the unit-sum beam painted at the tile centre on the tile's own grid (the
    legacy painter's beam convolution), times the pixel window when asked."""
    from . import filters as _filters
    H, W = shape
    x0, y0 = int(W / 2), int(H / 2)            # pixel scales at the central pixel, as MapFilter.makeRadiansMap (filters.py:214-239)
    ra0, dec0 = wcs.pix2wcs(x0, y0)
    ra1, dec1 = wcs.pix2wcs(x0 + 1, y0 + 1)
    degPerPixX, degPerPixY = _wcs.calcAngSepDeg(ra0, dec0, ra1, dec0), _wcs.calcAngSepDeg(ra0, dec0, ra0, dec1)
    key = (H, W, float(degPerPixX), float(degPerPixY), _filters._beamKey(beam), bool(applyPixelWindow), ops.precision, str(ops.device))
    if key not in _beamTransformCache:
        xDeg = np.arange(int(np.floor(-W / 2)), int(W / 2), dtype=np.float64) * degPerPixX
        yDeg = np.arange(int(np.floor(-H / 2)), int(H / 2), dtype=np.float64) * degPerPixY
        beamMap = signals.makeBeamModelSignalMap(np.sqrt(xDeg[None, :] ** 2 + yDeg[:, None] ** 2), wcs, beam)
        beamMap = beamMap / beamMap.sum()
        mult = np.fft.rfft2(np.fft.ifftshift(beamMap)).real
        if applyPixelWindow:
            wy = np.sinc(np.fft.fftfreq(H))
            wx = np.sinc(np.fft.fftfreq(W))[:W // 2 + 1].copy()
            if W % 2 == 0:
                wx[W // 2] = np.sinc(0.5)
            mult = mult * wy[:, None] * wx[None, :]
        _beamTransformCache[key] = ops.to_device(np.ascontiguousarray(mult))
    return _beamTransformCache[key]


def paintModelImages(ops, shape, wcsList, catalogList, beamList, obsFreqGHzList=None, GNFWParams='default', applyPixelWindow=True,
                     override=None, validAreaSection=None, minSNR=-99, TCMBAlpha=0, base=None, out=None, maxSizeDeg=None):
    """makeModelImage for a batch of images on the device: image i is painted on wcsList[i] from catalogList[i] with beam
    beamList[i] (and converted to delta T at obsFreqGHzList[i] for cluster catalogs).  Returns a device tensor [nImg, H, W]
    (base + model when `base` [nImg, H, W] is given; written to `out` when given) and the number of objects painted per image.
    maxSizeDeg: paint clusters out to this radius instead of makeModelImage's 5 theta500 (signals.makeArnaudModelSignalMap's own
    default is 15 degrees, which is what the filter templates and signals.fitQ use)."""
    H, W = shape
    nImg = len(catalogList)
    rows, starts, profR, profV, profIndex = [], [0], [], [], {}
    counts = []
    cluster = False
    for i in range(nImg):
        cols = selectModelObjects(catalogList[i], shape, wcsList[i], validAreaSection=validAreaSection, minSNR=minSNR)
        n = len(cols['RADeg'])
        counts.append(n)
        beam = signals._asBeam(beamList[i])
        if n > 0 and ('y_c' in cols or 'true_y_c' in cols):
            cluster = True
            # clusters: the projected GNFW profile of every (M500, z) in the catalog, tabulated out to 5 theta500 (maps.py:1843, 1872)
            if override is not None:
                zs, Ms = np.full(n, float(override['redshift'])), np.full(n, float(override['M500']))
                y0 = cols['y_c'].astype(np.float64) * 1e-4
            elif 'true_M500c' in cols:
                zs, Ms, y0 = cols['redshift'].astype(np.float64), cols['true_M500c'].astype(np.float64) * 1e14, cols['true_y_c'].astype(np.float64) * 1e-4
            else:
                if 'template' not in cols:
                    raise Exception("No M500, z, or template column found in catalog.")
                bits = [str(t).split("#")[0].split("_") for t in cols['template']]
                Ms = np.array([float(b[1][1:].replace("p", ".")) for b in bits])
                zs = np.array([float(b[2][1:].replace("p", ".")) for b in bits])
                y0 = cols['y_c'].astype(np.float64) * 1e-4
            amp = y0
            if obsFreqGHzList is not None and obsFreqGHzList[i] is not None:
                amp = np.array([convertToDeltaT(a, obsFreqGHzList[i], TCMBAlpha=TCMBAlpha, z=z) for a, z in zip(y0, zs)])
            pidx = np.empty(n)
            maxSize = np.empty(n)
            for j in range(n):
                key = ('A10', float(zs[j]), float(Ms[j]), repr(GNFWParams))
                if key not in profIndex:
                    prof = signals.makeArnaudModelProfile(zs[j], Ms[j], GNFWParams=GNFWParams)
                    size = 5 * (prof['theta500Arcmin'] / 60) if maxSizeDeg is None else float(maxSizeDeg)
                    rDeg = np.linspace(0.0, size, 5000)
                    profIndex[key] = (len(profR), size)
                    profR.append(rDeg)
                    profV.append(np.asarray(interpolate.splev(rDeg, prof['tckP'], ext=1), dtype=np.float64))
                pidx[j], maxSize[j] = profIndex[key]
            boxes = objectBoxes(wcsList[i], cols['RADeg'], cols['decDeg'], maxSize, shape)
        elif n > 0:
            # point sources: the beam profile out to 5 FWHM, amplitude deltaT_c (maps.py:1803-1806, 1878-1889)
            key = ('beam', _beamIdent(beam))
            if key not in profIndex:
                profIndex[key] = (len(profR), (beam.FWHMArcmin * 5.0) / 60)
                profR.append(np.asarray(beam.rDeg, dtype=np.float64))
                profV.append(1.0 * np.asarray(beam.profile1d, dtype=np.float64))
            p, maxSizeDeg = profIndex[key]
            pidx = np.full(n, p)
            amp = cols['deltaT_c'].astype(np.float64)
            boxes = objectBoxes(wcsList[i], cols['RADeg'], cols['decDeg'], maxSizeDeg, shape)
        if n > 0:
            rows.append(np.column_stack([boxes[:, :4], amp, boxes[:, 4:], pidx]))
        starts.append(starts[-1] + n)
    if not profR:
        profR, profV = [np.array([0.0, 1.0])], [np.array([0.0, 0.0])]
    # tables of different lengths share one array: shorter ones are continued with radii beyond every map and value 0
    nR = max(len(r) for r in profR)
    tabR = np.stack([np.concatenate([r, r[-1] + 1e3 * (1 + np.arange(nR - len(r)))]) for r in profR])
    tabV = np.stack([np.concatenate([v, np.zeros(nR - len(v))]) for v in profV])
    for k, r in enumerate(profR):      # numpy.interp returns 0 beyond the last radius of the original table: keep that
        if len(r) < nR:
            tabV[k, len(r) - 1:] = np.concatenate([[profV[k][-1]], np.zeros(nR - len(r))])
    objects = np.concatenate(rows, axis=0) if rows else np.zeros((0, 10))
    needPost = cluster or applyPixelWindow
    model = ops.paint_objects((nImg, H, W), objects, np.array(starts, dtype=np.int32), tabR, tabV, base=None if needPost else base,
                              out=None if (needPost or out is None) else out.reshape(nImg, H, W))
    if cluster:
        # beam convolution on the periodic tile grid and the pixel window in one Fourier multiplication
        mult = ops.torch.stack([_beamTimesWindow(ops, shape, wcsList[i], signals._asBeam(beamList[i]), applyPixelWindow) for i in range(nImg)])
        fm = ops.rfft2(model).reshape(nImg, 1, H, W // 2 + 1)
        # (rfft2 carries enmap.fft's 1 / sqrt(HW); apply_filter's inverse is the un-normalised one of filters.py:851-857)
        model = ops.apply_filter(fm, mult.reshape(nImg, 1, H, W // 2 + 1).contiguous(), 1, scale=1.0 / np.sqrt(H * W), W=W)
    elif applyPixelWindow:
        model = ops.apply_window(model, 1.0, restoreZeros=False)
    if needPost and base is not None:
        model = ops.add_maps(base.reshape(nImg, H, W), model, out=None if out is None else out.reshape(nImg, H, W))
    elif needPost and out is not None:
        out.reshape(nImg, H, W).copy_(model)
        model = out.reshape(nImg, H, W)
    return model, counts


def _beamIdent(beam):
    from . import filters as _filters
    return _filters._beamKey(beam)


def makeModelImage(shape, wcs, catalog, beamFileName, obsFreqGHz=None, GNFWParams='default', profile='A10', cosmoModel=None,
                   applyPixelWindow=True, override=None, validAreaSection=None, minSNR=-99, TCMBAlpha=0, ops=None):
    """Map of the given shape and WCS containing model clusters or point sources with the properties listed in the catalog
    (maps.py:1734-1899); numpy array, or None when no object of the catalog lies in the map.  Point sources ('deltaT_c'
    catalogs) are painted exactly as the reference does; clusters ('y_c' catalogs, profile 'A10') with this package's painter:
    projected GNFW profile on the object's distance map, beam convolution on the periodic tile grid."""
    from . import device
    if profile != 'A10':
        raise Exception("only profile 'A10' is implemented (B12 needs pyccl)")
    ops = ops if ops is not None else device.getOps('f64')
    model, counts = paintModelImages(ops, tuple(shape), [wcs], [catalog], [beamFileName], [obsFreqGHz], GNFWParams=GNFWParams,
                                     applyPixelWindow=applyPixelWindow, override=override, validAreaSection=validAreaSection,
                                     minSNR=minSNR, TCMBAlpha=TCMBAlpha)
    if counts[0] == 0:
        return None
    return device.toHost(model[0]).astype(np.float64)
