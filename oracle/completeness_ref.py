"""
Oracle restatement of the noise-level-versus-area table the reference compresses each tile's RMS map into
(``completeness.getRMSTab``, nemo/completeness.py:949-1005, the array part: the reference reads the RMS map and the area mask
from disk first) and of ``maps.getPixelAreaArcmin2Map`` (nemo/maps.py:1461-1482).  numpy only.
TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""

import numpy as np

from . import wcs_stub


def getPixelAreaArcmin2Map(shape, wcs):
    """Pixel area in arcmin^2, constant along a row (maps.py:1461-1482)."""
    pixAreasDeg2 = []
    RACentre, decCentre = wcs.getCentreWCSCoords()
    x0, y0 = wcs.wcs2pix(RACentre, decCentre)
    x1 = x0 + 1
    for y0 in range(shape[0]):
        y1 = y0 + 1
        ra0, dec0 = wcs.pix2wcs(x0, y0)
        ra1, dec1 = wcs.pix2wcs(x1, y1)
        xPixScale = wcs_stub.calcAngSepDeg(ra0, dec0, ra1, dec0)
        yPixScale = wcs_stub.calcAngSepDeg(ra0, dec0, ra0, dec1)
        pixAreasDeg2.append(xPixScale * yPixScale)
    pixAreasDeg2 = np.array(pixAreasDeg2)
    pixAreasArcmin2 = pixAreasDeg2 * (60 ** 2)
    return np.array([pixAreasArcmin2] * shape[1]).transpose()


def getRMSTab(RMSMap, areaMap, wcs):
    """(RMSValues ascending, tileArea in deg^2 per value) -- completeness.py:981-999; raises like the reference when the areas
    do not add up to the masked area of the tile."""
    areaMapSqDeg = (getPixelAreaArcmin2Map(areaMap.shape, wcs) * areaMap) / (60 ** 2)
    RMSValues = np.unique(RMSMap[np.nonzero(RMSMap)])
    tileArea = np.zeros(len(RMSValues))
    for i in range(len(RMSValues)):
        tileArea[i] = areaMapSqDeg[np.equal(RMSMap, RMSValues[i])].sum()
    tol = 0.003
    if abs(tileArea.sum() - areaMapSqDeg.sum()) > tol:
        raise Exception("Mismatch between area map and area in RMSTab")
    if np.less(tileArea, 0).sum() > 0:
        raise Exception("Negative area in tile - check your survey mask")
    return RMSValues, tileArea
