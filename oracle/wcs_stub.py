"""
Minimal stand-in for ``astLib.astWCS.WCS`` / ``astCoords`` restricted to the CAR (plate carree)
projection the synthetic tiles use.  astLib (>= 0.11.8, reference setup.py:39) is absent from this image;
only the methods the hot path calls are provided (filters.py:227-230, 417, 655-656, 729;
photometry.py:109-112, 250, 328; maps.py:1444-1455).  Pixel coordinates are 0-indexed, as in astLib's
NUMPY_MODE.  TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""

import numpy as np


class SimpleCARWCS(object):

    def __init__(self, shape, pixDeg=0.5 / 60.0, crval=(0.0, 0.0), crpix=None, header=None):
        self.shape = (int(shape[0]), int(shape[1]))
        self.pixDeg = float(pixDeg)
        self.crval = (float(crval[0]), float(crval[1]))
        if crpix is None:
            # 0-indexed reference pixel at the centre of the tile
            crpix = (self.shape[1] / 2.0, self.shape[0] / 2.0)
        self.crpix = (float(crpix[0]), float(crpix[1]))
        self.header = {'NAXIS1': self.shape[1], 'NAXIS2': self.shape[0], 'CTYPE1': 'RA---CAR',
                       'CTYPE2': 'DEC--CAR', 'CDELT1': -self.pixDeg, 'CDELT2': self.pixDeg,
                       'CRVAL1': self.crval[0], 'CRVAL2': self.crval[1],
                       'CRPIX1': self.crpix[0] + 1, 'CRPIX2': self.crpix[1] + 1}
        if header is not None:
            self.header.update(header)

    def copy(self):
        return SimpleCARWCS(self.shape, self.pixDeg, self.crval, self.crpix, dict(self.header))

    def updateFromHeader(self):
        return None

    def pix2wcs(self, x, y):
        ra = self.crval[0] - (np.asarray(x, dtype=np.float64) - self.crpix[0]) * self.pixDeg
        dec = self.crval[1] + (np.asarray(y, dtype=np.float64) - self.crpix[1]) * self.pixDeg
        if np.ndim(ra) == 0:
            return float(ra), float(dec)
        return ra, dec

    def wcs2pix(self, RADeg, decDeg):
        ra = np.asarray(RADeg, dtype=np.float64)
        # RA wraps: bring within +/-180 of the reference value
        dra = (ra - self.crval[0] + 180.0) % 360.0 - 180.0
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


def calcAngSepDeg(RADeg1, decDeg1, RADeg2, decDeg2):
    """astLib ``astCoords.calcAngSepDeg``: tangent-plane separation of position 2 about position 1."""
    cRA = np.radians(RADeg1)
    cDec = np.radians(decDeg1)
    gRA = np.radians(RADeg2)
    gDec = np.radians(decDeg2)
    cosC = (np.sin(gDec) * np.sin(cDec)) + (np.cos(gDec) * np.cos(cDec) * np.cos(gRA - cRA))
    x = (np.cos(cDec) * np.sin(gRA - cRA)) / cosC
    y = ((np.cos(gDec) * np.sin(cDec)) - (np.sin(gDec) * np.cos(cDec) * np.cos(gRA - cRA))) / cosC
    return np.degrees(np.sqrt(x * x + y * y))


# J2000 -> Galactic rotation (IAU 1958 pole, J2000 values), for astCoords.convertCoords at photometry.py:112
_J2000_TO_GAL = np.array([[-0.0548755604162154, -0.8734370902348850, -0.4838350155487132],
                          [+0.4941094278755837, -0.4448296299600112, +0.7469822444972189],
                          [-0.8676661490190047, -0.1980763734312015, +0.4559837761750669]])


def j2000ToGalactic(RADeg, decDeg):
    ra = np.radians(RADeg)
    dec = np.radians(decDeg)
    v = np.array([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)])
    g = _J2000_TO_GAL @ v
    lon = np.degrees(np.arctan2(g[1], g[0])) % 360.0
    lat = np.degrees(np.arcsin(np.clip(g[2], -1, 1)))
    return lon, lat
