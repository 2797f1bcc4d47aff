"""
Host-side coordinate helpers.  The reference uses ``astLib.astWCS.WCS`` / ``astCoords`` (third party, not
importable in every environment); any object with the same methods (``pix2wcs``, ``wcs2pix``,
``getCentreWCSCoords``, ``getPixelSizeDeg``, ``getImageMinMaxWCSCoords``, ``header``, ``copy``,
``updateFromHeader``) can be passed to nemo_b200 -- a real astWCS.WCS works unchanged.  :class:`CARWCS` is a
small plate-carree implementation of that interface for synthetic tiles and tests (0-indexed pixels, as
astLib's NUMPY_MODE).
"""

import numpy as np


class CARWCS(object):

    def __init__(self, shape, pixDeg=0.5 / 60.0, crval=(0.0, 0.0), crpix=None, header=None):
        self.shape = (int(shape[0]), int(shape[1]))
        self.pixDeg = float(pixDeg)
        self.crval = (float(crval[0]), float(crval[1]))
        self.crpix = (self.shape[1] / 2.0, self.shape[0] / 2.0) if crpix is None else (float(crpix[0]), float(crpix[1]))
        self.header = {'NAXIS1': self.shape[1], 'NAXIS2': self.shape[0], 'CTYPE1': 'RA---CAR', 'CTYPE2': 'DEC--CAR',
                       'CDELT1': -self.pixDeg, 'CDELT2': self.pixDeg, 'CRVAL1': self.crval[0], 'CRVAL2': self.crval[1],
                       'CRPIX1': self.crpix[0] + 1, 'CRPIX2': self.crpix[1] + 1}
        if header:
            self.header.update(header)

    def copy(self):
        return CARWCS(self.shape, self.pixDeg, self.crval, self.crpix, dict(self.header))

    def updateFromHeader(self):
        return None

    def pix2wcs(self, x, y):
        ra = self.crval[0] - (np.asarray(x, dtype=np.float64) - self.crpix[0]) * self.pixDeg
        dec = self.crval[1] + (np.asarray(y, dtype=np.float64) - self.crpix[1]) * self.pixDeg
        if np.ndim(ra) == 0:
            return float(ra), float(dec)
        return ra, dec

    def wcs2pix(self, RADeg, decDeg):
        dra = (np.asarray(RADeg, dtype=np.float64) - self.crval[0] + 180.0) % 360.0 - 180.0
        x = self.crpix[0] - dra / self.pixDeg
        y = self.crpix[1] + (np.asarray(decDeg, dtype=np.float64) - self.crval[1]) / self.pixDeg
        if np.ndim(x) == 0:
            return float(x), float(y)
        return x, y

    def getCentreWCSCoords(self):
        return self.pix2wcs(self.shape[1] / 2.0, self.shape[0] / 2.0)

    def getPixelSizeDeg(self):
        return self.pixDeg

    def getImageMinMaxWCSCoords(self):
        ra0, dec0 = self.pix2wcs(0, 0)
        ra1, dec1 = self.pix2wcs(self.shape[1] - 1, self.shape[0] - 1)
        return min(ra0, ra1), max(ra0, ra1), min(dec0, dec1), max(dec0, dec1)


def batchParams(wcsList):
    """Per-image (crval1, crval2, crpix1, crpix2, pixDeg) arrays when every wcs of the list is a CARWCS, else None: the batched
    driver then converts the positions of all images of a batch with one expression (same arithmetic as the methods)."""
    if not all(type(w) is CARWCS for w in wcsList):
        return None
    return (np.array([w.crval[0] for w in wcsList]), np.array([w.crval[1] for w in wcsList]),
            np.array([w.crpix[0] for w in wcsList]), np.array([w.crpix[1] for w in wcsList]),
            np.array([w.pixDeg for w in wcsList]))


def batchPix2wcs(params, owner, x, y):
    v0, v1, p0, p1, pd = (a[owner] for a in params)
    return v0 - (np.asarray(x, dtype=np.float64) - p0) * pd, v1 + (np.asarray(y, dtype=np.float64) - p1) * pd


def batchWcs2pix(params, owner, RADeg, decDeg):
    v0, v1, p0, p1, pd = (a[owner] for a in params)
    dra = (np.asarray(RADeg, dtype=np.float64) - v0 + 180.0) % 360.0 - 180.0
    return p0 - dra / pd, p1 + (np.asarray(decDeg, dtype=np.float64) - v1) / pd


def calcAngSepDeg(RADeg1, decDeg1, RADeg2, decDeg2):
    """Tangent-plane angular separation (what astCoords.calcAngSepDeg returns), degrees."""
    a0, d0 = np.radians(RADeg1), np.radians(decDeg1)
    a1, d1 = np.radians(RADeg2), np.radians(decDeg2)
    cosC = np.sin(d1) * np.sin(d0) + np.cos(d1) * np.cos(d0) * np.cos(a1 - a0)
    x = np.cos(d0) * np.sin(a1 - a0) / cosC
    y = (np.cos(d1) * np.sin(d0) - np.sin(d1) * np.cos(d0) * np.cos(a1 - a0)) / cosC
    return np.degrees(np.sqrt(x * x + y * y))


_EQ2GAL = np.array([[-0.0548755604162154, -0.8734370902348850, -0.4838350155487132],
                    [+0.4941094278755837, -0.4448296299600112, +0.7469822444972189],
                    [-0.8676661490190047, -0.1980763734312015, +0.4559837761750669]])


def j2000ToGalactic(RADeg, decDeg):
    """(l, b) in degrees; stands in for astCoords.convertCoords("J2000", "GALACTIC", ...)."""
    ra, dec = np.radians(RADeg), np.radians(decDeg)
    g = _EQ2GAL @ np.array([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)])
    return np.degrees(np.arctan2(g[1], g[0])) % 360.0, np.degrees(np.arcsin(np.clip(g[2], -1, 1)))


def pixelScalesDeg(wcs, RADeg='centre', decDeg='centre'):
    """(xPixScale, yPixScale) in degrees at a sky position, computed as maps.smoothMap does (reference
    nemo/maps.py:1444-1455)."""
    ra0, dec0 = wcs.getCentreWCSCoords()
    if not isinstance(RADeg, str):
        ra0 = float(RADeg)
    if not isinstance(decDeg, str):
        dec0 = float(decDeg)
    x0, y0 = wcs.wcs2pix(ra0, dec0)
    ra1, dec1 = wcs.pix2wcs(x0 + 1, y0 + 1)
    return calcAngSepDeg(ra0, dec0, ra1, dec0), calcAngSepDeg(ra0, dec0, ra0, dec1)
