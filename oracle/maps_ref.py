"""
TEST INFRASTRUCTURE (see oracle/__init__.py): restatement of the reference's model-image painter, the first half of the
source-injection loop -- ``maps.makeModelImage`` (nemo/maps.py:1734-1899), ``maps.makeDegreesDistanceMap``
(:2414-2471) and ``catalogs.getCatalogWithinImage`` (nemo/catalogs.py:855-892).

PARITY STATUS: the point-source branch (``deltaT_c`` catalogs: distance map in a box around each source, linear interpolation of
the beam profile, sum in catalog order, pixel window) is PINNED against a run of the reference's own ``maps.makeModelImage``
(tests/golden/ref_paint.npz, tests/golden/make_reference_golden.py::runPaint); the cluster branch paints with the restated
legacy painter of oracle/signals_ref.py (the reference's default painter is pixell's ``sim_objects``, absent here) and is
unpinned.
"""

import numpy as np
from scipy import interpolate

from . import pixell_shim as enmap
from . import signals_ref as signals
from .wcs_stub import calcAngSepDeg


def getCatalogWithinImage(RADeg, decDeg, shape, wcs, mask=None):
    """Boolean selection of the rows inside the image (catalogs.py:874-892)."""
    x, y = wcs.wcs2pix(np.asarray(RADeg, dtype=np.float64), np.asarray(decDeg, dtype=np.float64))
    x, y = np.atleast_1d(x), np.atleast_1d(y)
    sel = np.zeros(len(x), dtype=bool)
    for i in range(len(x)):
        if np.isnan(x[i]) or np.isnan(y[i]):
            continue
        if x[i] >= 0 and x[i] < shape[1] - 1 and y[i] >= 0 and y[i] < shape[0] - 1:
            sel[i] = True if mask is None else (mask[int(round(y[i])), int(round(x[i]))] == 1)
    return sel


def makeDegreesDistanceMap(degreesMap, wcs, RADeg, decDeg, maxDistDegrees):
    """maps.py:2414-2471, line for line."""
    x0, y0 = wcs.wcs2pix(RADeg, decDeg)
    ra0, dec0 = RADeg, decDeg
    ra1, dec1 = wcs.pix2wcs(x0 + 1, y0 + 1)
    xPixScale = calcAngSepDeg(ra0, dec0, ra1, dec0)
    yPixScale = calcAngSepDeg(ra0, dec0, ra0, dec1)
    xDistPix = int(round((maxDistDegrees) / xPixScale))
    yDistPix = int(round((maxDistDegrees) / yPixScale))
    Y, X = degreesMap.shape
    minX = max(int(round(x0)) - xDistPix, 0)
    maxX = min(int(round(x0)) + xDistPix, X)
    minY = max(int(round(y0)) - yDistPix, 0)
    maxY = min(int(round(y0)) + yDistPix, Y)
    xDeg = (np.arange(degreesMap.shape[1]) - x0) * xPixScale
    yDeg = (np.arange(degreesMap.shape[0]) - y0) * yPixScale
    for i in range(minY, maxY):
        degreesMap[i][minX:maxX] = np.sqrt(yDeg[i] ** 2 + xDeg[minX:maxX] ** 2)
    return degreesMap, [minX, maxX], [minY, maxY]


def centreDegreesMap(shape, wcs):
    """Distance in degrees from the tile centre on the tile's own grid, as MapFilter.makeRadiansMap builds it (filters.py:214-239)."""
    x0, y0 = int(shape[1] / 2), int(shape[0] / 2)
    ra0, dec0 = wcs.pix2wcs(x0, y0)
    ra1, dec1 = wcs.pix2wcs(x0 + 1, y0 + 1)
    degPerPixX = calcAngSepDeg(ra0, dec0, ra1, dec0)
    degPerPixY = calcAngSepDeg(ra0, dec0, ra0, dec1)
    xDeg = np.arange(int(np.floor(-shape[1] / 2)), int(shape[1] / 2), dtype=np.float64) * degPerPixX
    yDeg = np.arange(int(np.floor(-shape[0] / 2)), int(shape[0] / 2), dtype=np.float64) * degPerPixY
    return np.sqrt(xDeg[None, :] ** 2 + yDeg[:, None] ** 2)


def convolveWithBeamOnGrid(m, wcs, beam):
    """Beam convolution on the periodic tile grid with the unit-sum beam painted at the tile centre (the legacy painter's
    stand-in for maps.convolveMapWithBeam, as oracle/filters_ref.py ArnaudModelFilter does it)."""
    beamMap = signals.makeBeamModelSignalMap(centreDegreesMap(m.shape, wcs), wcs, beam)
    beamMap = beamMap / beamMap.sum()
    return np.fft.ifft2(np.fft.fft2(m) * np.fft.fft2(np.fft.ifftshift(beamMap))).real


def makeModelImage(shape, wcs, catalog, beam, obsFreqGHz=None, GNFWParams='default', applyPixelWindow=True, override=None,
                   validAreaSection=None, minSNR=-99):
    """maps.makeModelImage for a catalog given as a dict of columns.  Sources: 'RADeg', 'decDeg', 'deltaT_c'.  Clusters:
    'RADeg', 'decDeg', 'y_c' (units of 1e-4) with `override` = {'M500', 'redshift'} (the source-injection case, :1834-1849) or
    per-row 'template' names like Arnaud_M2e14_z0p4 (:1866-1872)."""
    modelMap = np.zeros(shape, dtype=float)
    cols = {k: np.asarray(v) for k, v in catalog.items()}
    sel = getCatalogWithinImage(cols['RADeg'], cols['decDeg'], shape, wcs)
    SNRKey = 'SNR' if 'SNR' in cols else ('fixed_SNR' if 'fixed_SNR' in cols else None)
    if SNRKey is not None:
        sel &= cols[SNRKey] > minSNR
    cols = {k: v[sel] for k, v in cols.items()}
    if validAreaSection is not None and len(cols['RADeg']) > 0:
        x0, x1, y0, y1 = validAreaSection
        x, y = wcs.wcs2pix(cols['RADeg'], cols['decDeg'])
        cMask = (x >= x0) & (x < x1) & (y >= y0) & (y < y1)
        cols = {k: v[cMask] for k, v in cols.items()}
    if len(cols['RADeg']) == 0:
        return None
    beam = signals.BeamProfile(beamFileName=beam) if isinstance(beam, str) else beam
    maxSizeDeg = (beam.FWHMArcmin * 5.0) / 60
    if 'y_c' in cols or 'true_y_c' in cols:
        def clusterMap(z, M500, RADeg, decDeg, amplitude, maxSize):
            # legacy painter: projected GNFW profile on the distance map of the object, beam convolution on the periodic grid
            tckP = signals.makeArnaudModelProfile(z, M500, GNFWParams=GNFWParams)['tckP']
            rDeg = np.linspace(0.0, maxSize, 5000)
            profile1d = interpolate.splev(rDeg, tckP, ext=1)
            r2p = interpolate.interp1d(rDeg, profile1d, bounds_error=False, fill_value=0.0)
            out = np.zeros(shape)
            for ra, dec, a in zip(np.atleast_1d(RADeg), np.atleast_1d(decDeg), np.atleast_1d(amplitude)):
                degreesMap = np.ones(shape, dtype=float) * 1e6
                degreesMap, _, _ = makeDegreesDistanceMap(degreesMap, wcs, ra, dec, maxSize)
                out = out + r2p(degreesMap) * a
            return convolveWithBeamOnGrid(out, wcs, beam)
        if override is not None:
            z, M500 = override['redshift'], override['M500']
            theta500Arcmin = signals.calcTheta500Arcmin(z, M500)
            modelMap = clusterMap(z, M500, cols['RADeg'], cols['decDeg'], cols['y_c'] * 1e-4, 5 * (theta500Arcmin / 60))
            if obsFreqGHz is not None:
                modelMap = signals.convertToDeltaT(modelMap, obsFreqGHz)
        else:
            for i in range(len(cols['RADeg'])):
                bits = str(cols['template'][i]).split("#")[0].split("_")
                M500 = float(bits[1][1:].replace("p", "."))
                z = float(bits[2][1:].replace("p", "."))
                theta500Arcmin = signals.calcTheta500Arcmin(z, M500)
                signalMap = clusterMap(z, M500, cols['RADeg'][i], cols['decDeg'][i], cols['y_c'][i] * 1e-4, 5 * (theta500Arcmin / 60))
                if obsFreqGHz is not None:
                    signalMap = signals.convertToDeltaT(signalMap, obsFreqGHz)
                modelMap = modelMap + signalMap
    else:
        for i in range(len(cols['RADeg'])):
            degreesMap = np.ones(modelMap.shape, dtype=float) * 1e6
            degreesMap, xBounds, yBounds = makeDegreesDistanceMap(degreesMap, wcs, cols['RADeg'][i], cols['decDeg'][i], maxSizeDeg)
            signalMap = signals.makeBeamModelSignalMap(degreesMap, wcs, beam) * cols['deltaT_c'][i]
            modelMap = modelMap + signalMap
    if applyPixelWindow:
        modelMap = enmap.apply_window(modelMap, pow=1.0)
    return modelMap
