"""
Restatement of the signal / unit helpers the hot path calls.  TEST INFRASTRUCTURE -- see
oracle/__init__.py.  Citations are to /root/reference/nemo.

  fSZ                     signals.py:350-375
  convertToDeltaT         maps.py:670-688
  deltaTToJyPerSr         photometry.py:460-474
  smoothMap               maps.py:1433-1458
  subtractBackground      maps.py:1355-1368
  makeBeamModelSignalMap  signals.py:587-619
  gnfw line-of-sight integral  gnfw.py:35-110 ; makeArnaudModelProfile signals.py:448-502

Template maps are *inputs* to the hot path: the pixell painter
(signals.py:622-672) and pyccl cosmology are absent, so the Arnaud template here is the reference's
'legacy' painter idea (1-D profile -> radius map -> beam convolution) with a flat LCDM theta500; the
same arrays are handed to the oracle and to the CUDA path, so parity does not depend on it.
"""

import numpy as np
from scipy import ndimage, interpolate, integrate
from scipy.optimize import fmin

from .wcs_stub import calcAngSepDeg

TCMB = 2.72548
# CODATA 2018 (what astropy.constants returns)
_h = 6.62607015e-34
_kB = 1.380649e-23
_c = 299792458.0


def fSZ(obsFrequencyGHz):
    x = (_h * obsFrequencyGHz * 1e9) / (_kB * TCMB)
    return x * ((np.exp(x) + 1) / (np.exp(x) - 1)) - 4.0


def convertToDeltaT(mapData, obsFrequencyGHz=148):
    return mapData * fSZ(obsFrequencyGHz) * (TCMB * 1e6)


def deltaTToJyPerSr(temp, obsFreqGHz):
    kB = _kB * 1e7
    h = _h * 1e7
    c = _c * 100
    nu = obsFreqGHz * 1.e9
    x = h * nu / (kB * TCMB)
    cNu = 2 * (kB * TCMB) ** 3 / (h ** 2 * c ** 2) * x ** 4 / (4 * (np.sinh(x / 2.)) ** 2)
    cNu *= 1e23
    return temp * cNu * 1e-6 / TCMB


def pixelScalesDeg(wcs, RADeg='centre', decDeg='centre'):
    """x/y pixel scale at a sky position, as smoothMap computes it (maps.py:1444-1455)."""
    ra0, dec0 = wcs.getCentreWCSCoords()
    if RADeg != 'centre':
        ra0 = float(RADeg)
    if decDeg != 'centre':
        dec0 = float(decDeg)
    x0, y0 = wcs.wcs2pix(ra0, dec0)
    ra1, dec1 = wcs.pix2wcs(x0 + 1, y0 + 1)
    xPixScale = calcAngSepDeg(ra0, dec0, ra1, dec0)
    yPixScale = calcAngSepDeg(ra0, dec0, ra0, dec1)
    return xPixScale, yPixScale


def smoothMap(data, wcs, RADeg='centre', decDeg='centre', smoothScaleDeg=5.0 / 60.0):
    xPixScale, yPixScale = pixelScalesDeg(wcs, RADeg, decDeg)
    return ndimage.gaussian_filter(data, (smoothScaleDeg / yPixScale, smoothScaleDeg / xPixScale))


def subtractBackground(data, wcs, RADeg='centre', decDeg='centre', smoothScaleDeg=30.0 / 60.0):
    return data - smoothMap(data, wcs, RADeg=RADeg, decDeg=decDeg, smoothScaleDeg=smoothScaleDeg)


class BeamProfile(object):
    """Tabulated radial beam (signals.py:82-137 with profile1d/rDeg given directly)."""

    def __init__(self, profile1d=None, rDeg=None, beamFileName=None):
        if beamFileName is not None:        # two-column text file: degrees, response (signals.py:107-122)
            beamData = np.loadtxt(beamFileName).transpose()
            profile1d, rDeg = beamData[1], beamData[0]
        self.profile1d = np.asarray(profile1d, dtype=np.float64)
        self.rDeg = np.asarray(rDeg, dtype=np.float64)
        self.FWHMArcmin = self.rDeg[np.argmin(abs(self.profile1d - 0.5))] * 60 * 2      # signals.py:136


def gaussianBeam(FWHMArcmin, maxDeg=0.5, n=1800):
    rDeg = np.linspace(0.0, maxDeg, n)
    sigmaDeg = FWHMArcmin / 60.0 / np.sqrt(8 * np.log(2))
    return BeamProfile(np.exp(-0.5 * (rDeg / sigmaDeg) ** 2), rDeg)


def makeBeamModelSignalMap(degreesMap, wcs, beam, amplitude=None):
    if amplitude is None:
        amplitude = 1.0
    profile1d = amplitude * beam.profile1d
    r2p = interpolate.interp1d(beam.rDeg, profile1d, bounds_error=False, fill_value=0.0)
    return r2p(degreesMap)


# ---------------------------------------------------------------------------------------------------
# GNFW (gnfw.py)
GNFW_DEFAULT = {'P0': 8.403, 'c500': 1.177, 'gamma': 0.3081, 'alpha': 1.0510, 'beta': 5.4905,
                'tol': 1e-7, 'npts': 100}


def _gnfw_func(x, params):
    G, A, B, c500, P0 = params['gamma'], params['alpha'], params['beta'], params['c500'], params['P0']
    prof = np.zeros(x.shape)
    mask = np.greater(x, 0)
    prof[mask] = P0 * ((x[mask] * c500) ** -G * (1 + (x[mask] * c500) ** A) ** ((G - B) / A))
    return prof


def _gnfw_xfunc(x, b, params):
    x = np.array(x)
    if x.ndim == 0:
        return _gnfw_xfunc([x], b, params)[0]
    r = (x ** 2 + b ** 2) ** .5
    y = x * _gnfw_func(r, params)
    y[x == 0] = 0
    return y


def gnfwIntegrated(b, params=GNFW_DEFAULT):
    G, A, B = params['gamma'], params['alpha'], params['beta']
    TH, N = params.get('tol', 1e-6), params.get('npts', 200)
    x_max = fmin(lambda x: -_gnfw_xfunc(x, b, params), b, disp=0)[0]
    y_max = _gnfw_xfunc(x_max, b, params)
    x_lo = (y_max * TH) ** (1 / (1 - G))
    x_hi = (y_max * TH) ** (1 / (1 - B))
    u_lo, u_hi = np.log(x_lo), np.log(x_hi)
    du = (u_hi - u_lo) / N
    x = np.exp(np.arange(u_lo, u_hi, du))
    I1 = np.sum(du * _gnfw_xfunc(x, b, params))
    x_hi = np.exp(u_hi)
    I2 = x_lo ** (1 - G) / (1 - G) + x_hi ** (1 - B) / (1 - B)
    return I1 + I2


def calcTheta500Arcmin(z, M500, Om0=0.3, H0=70.0):
    """theta500 for a flat LCDM cosmology (stands in for the pyccl call at signals.py:440-444)."""
    G = 4.30091e-9  # Mpc (km/s)^2 / MSun
    Ez = np.sqrt(Om0 * (1 + z) ** 3 + (1 - Om0))
    rhoCrit = 3 * (H0 * Ez) ** 2 / (8 * np.pi * G)  # MSun / Mpc^3
    R500Mpc = (3 * M500 / (4 * np.pi * 500 * rhoCrit)) ** (1 / 3.)
    DC = integrate.quad(lambda zz: 1.0 / np.sqrt(Om0 * (1 + zz) ** 3 + (1 - Om0)), 0, z)[0] * (_c / 1e3) / H0
    DA = DC / (1 + z)
    return np.degrees(np.arctan(R500Mpc / DA)) * 60.0


def makeArnaudModelProfile(z, M500, GNFWParams='default'):
    if GNFWParams == 'default':
        GNFWParams = GNFW_DEFAULT
    bRange = np.logspace(np.log10(1e-6), np.log10(100), 300)
    cylPProfile = []
    tol = 1e-6
    for i in range(len(bRange)):
        cylPProfile.append(gnfwIntegrated(bRange[i], params=GNFWParams))
        if i > 0 and abs(cylPProfile[i] - cylPProfile[i - 1]) < tol:
            break
    cylPProfile = np.array(cylPProfile)
    bRange = bRange[:i + 1]
    cylPProfile = cylPProfile / cylPProfile.max()
    theta500Arcmin = calcTheta500Arcmin(z, M500)
    thetaDegRange = bRange * (theta500Arcmin / 60.)
    tckP = interpolate.splrep(thetaDegRange, cylPProfile)
    return {'tckP': tckP, 'theta500Arcmin': theta500Arcmin, 'rDeg': thetaDegRange}


# ------------------------------------------------------------------------------------------------------
# signals.fitQ (nemo/signals.py:864-1125), Arnaud profile (zDepQ = 0)
def calcM500FromTheta500Arcmin(theta500Arcmin, z, Om0=0.3, H0=70.0):
    """signals.py:907-912 with the flat LCDM of calcTheta500Arcmin above instead of pyccl."""
    from scipy import integrate
    G = 4.30091e-9
    Ez = np.sqrt(Om0 * (1 + z) ** 3 + (1 - Om0))
    rhoCrit = 3 * (H0 * Ez) ** 2 / (8 * np.pi * G)
    DC = integrate.quad(lambda zz: 1.0 / np.sqrt(Om0 * (1 + zz) ** 3 + (1 - Om0)), 0, z)[0] * (299792458.0 / 1e3) / H0
    R500Mpc = np.tan(np.radians(theta500Arcmin / 60.0)) * (DC / (1 + z))
    return (4 / 3.0) * np.pi * R500Mpc ** 3 * 500 * rhoCrit


def fitQScales(refM500MSun, refz):
    """signals.py:898-915."""
    theta500Arcmin_wanted = np.power(10, np.arange(np.log10(0.1), np.log10(50), 0.05055349))
    zRange_wanted = [2.0] * 10 + [1.0] * 10 + [0.6] * 10 + [0.3] * 10 + [0.1] * 10 + [0.07] * 4
    MRange = [refM500MSun]
    zRange = [refz]
    for theta500Arcmin, z in zip(theta500Arcmin_wanted, zRange_wanted):
        MRange.append(calcM500FromTheta500Arcmin(theta500Arcmin, z))
        zRange.append(z)
    return zRange, MRange


def fitQTile(filterObj, mapDicts, zRange, MRange, y0=2e-04):
    """The per-tile body of signals.fitQ (signals.py:1029-1083) for a Fourier-space reference filter object of
    oracle/filters_ref.py (already built: filt and signalNorm set): (Q, theta500Arcmin, z) sorted by theta500Arcmin."""
    import copy
    from scipy import interpolate as _interp
    from . import pixell_shim as enmap
    wcs = filterObj.wcs
    RADeg, decDeg = wcs.getCentreWCSCoords()
    x, y = wcs.wcs2pix(RADeg, decDeg)
    Q, QTheta500Arcmin, Qz = [], [], []
    for z, M500MSun in zip(zRange, MRange):
        painter = copy.copy(filterObj)                 # same geometry, beam handling and painter; another cluster size
        painter.params = dict(filterObj.params, z=z, M500MSun=M500MSun)
        signalMaps = []
        for mapDict in mapDicts:
            amplitude = convertToDeltaT(y0, mapDict['obsFreqGHz']) if mapDict.get('units') != 'yc' else y0
            signalMap = painter.makeSignalTemplateMap(mapDict['beamFileName'], amplitude=amplitude)
            signalMap = enmap.apply_window(signalMap, pow=1.0)
            signalMaps.append(enmap.fft(signalMap))
        filteredSignal = filterObj.applyFilter(np.array(signalMaps))
        mapInterpolator = _interp.RectBivariateSpline(np.arange(filteredSignal.shape[0]), np.arange(filteredSignal.shape[1]),
                                                      filteredSignal, kx=3, ky=3)
        peakFilteredSignal = mapInterpolator(y, x)[0][0]
        if peakFilteredSignal not in Q:
            Q.append(peakFilteredSignal)
            QTheta500Arcmin.append(calcTheta500Arcmin(z, M500MSun))
            Qz.append(z)
    Q = np.array(Q)
    if abs(1 - Q[0] / y0) > 1e-6:
        raise Exception("Q[0]/y0 outside tolerance")
    Q = Q / y0
    order = np.argsort(np.array(QTheta500Arcmin), kind='stable')
    return Q[order], np.array(QTheta500Arcmin)[order], np.array(Qz)[order]
