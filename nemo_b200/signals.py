"""
Host-side signal templates and unit conversions used by the filter classes.  Template maps are *inputs* to the
GPU hot path: they are made once per (filter scale, frequency) on the host and uploaded.

Mirrors the names in the reference's nemo/signals.py (fSZ :350-375, BeamProfile :82-137,
makeBeamModelSignalMap :587-619, makeArnaudModelProfile :448-502, makeArnaudModelSignalMap :675-743) and
nemo/gnfw.py.  pixell's object painter and pyccl are not required: the Arnaud template follows the reference's
'legacy' painter (1-D line-of-sight GNFW profile -> radius map -> beam convolution) with a flat LCDM theta500.
"""

import numpy as np
from scipy import interpolate, integrate
from scipy.optimize import fmin

TCMB = 2.72548
_H_PLANCK = 6.62607015e-34
_K_BOLTZ = 1.380649e-23
_C_LIGHT = 299792458.0


def fSZ(obsFrequencyGHz, TCMBAlpha=0.0, z=None):
    """Non-relativistic SZ spectral function."""
    x = (_H_PLANCK * obsFrequencyGHz * 1e9) / (_K_BOLTZ * TCMB)
    if TCMBAlpha != 0 and z is not None:
        assert (z >= 0)
        x = x * np.power(1 + z, TCMBAlpha)
    return x * ((np.exp(x) + 1) / (np.exp(x) - 1)) - 4.0


class BeamProfile(object):
    """Radial beam profile.  Either a two-column text file (degrees, response) or arrays."""

    def __init__(self, beamFileName=None, profile1d=None, rDeg=None):
        if beamFileName is not None:
            beamData = np.loadtxt(beamFileName).transpose()
            if beamData[0][1] - beamData[0][0] >= 1:
                raise Exception("beam transform (B_ell) files need pixell.curvedsky.harm2profile; supply a radial profile instead")
            self.profile1d = beamData[1]
            self.rDeg = beamData[0]
        else:
            self.profile1d = np.asarray(profile1d, dtype=np.float64)
            self.rDeg = np.asarray(rDeg, dtype=np.float64)
        self.FWHMArcmin = self.rDeg[np.argmin(abs(self.profile1d - 0.5))] * 60 * 2


def gaussianBeam(FWHMArcmin, maxDeg=0.5, n=1800):
    rDeg = np.linspace(0.0, maxDeg, n)
    sigmaDeg = FWHMArcmin / 60.0 / np.sqrt(8 * np.log(2))
    return BeamProfile(profile1d=np.exp(-0.5 * (rDeg / sigmaDeg) ** 2), rDeg=rDeg)


def _asBeam(beam):
    if isinstance(beam, str):
        return BeamProfile(beamFileName=beam)
    return beam


def makeBeamModelSignalMap(degreesMap, wcs, beam, amplitude=None):
    """2-D map of the beam on the given map of angular distance (degrees) from the object."""
    if amplitude is None:
        amplitude = 1.0
    beam = _asBeam(beam)
    profile1d = amplitude * beam.profile1d
    r2p = interpolate.interp1d(beam.rDeg, profile1d, bounds_error=False, fill_value=0.0)
    return r2p(degreesMap)


# ------------------------------------------------------------------------------------------------------
# GNFW line-of-sight integral (Arnaud et al. 2010 universal pressure profile by default)
GNFW_DEFAULT_PARAMS = {'P0': 8.403, 'c500': 1.177, 'gamma': 0.3081, 'alpha': 1.0510, 'beta': 5.4905, 'tol': 1e-7, 'npts': 100}


def _gnfwProfile(x, params):
    G, A, B, c500, P0 = params['gamma'], params['alpha'], params['beta'], params['c500'], params['P0']
    prof = np.zeros(x.shape)
    mask = np.greater(x, 0)
    prof[mask] = P0 * ((x[mask] * c500) ** -G * (1 + (x[mask] * c500) ** A) ** ((G - B) / A))
    return prof


def _gnfwIntegrand(x, b, params):
    x = np.array(x)
    if x.ndim == 0:
        return _gnfwIntegrand([x], b, params)[0]
    r = (x ** 2 + b ** 2) ** .5
    y = x * _gnfwProfile(r, params)
    y[x == 0] = 0
    return y


def gnfwIntegrated(b, params=GNFW_DEFAULT_PARAMS):
    """Line-of-sight integral of the GNFW profile at impact parameter b (log-spaced Riemann sum + wings)."""
    G, B = params['gamma'], params['beta']
    TH, N = params.get('tol', 1e-6), params.get('npts', 200)
    x_max = fmin(lambda x: -_gnfwIntegrand(x, b, params), b, disp=0)[0]
    y_max = _gnfwIntegrand(x_max, b, params)
    x_lo = (y_max * TH) ** (1 / (1 - G))
    x_hi = (y_max * TH) ** (1 / (1 - B))
    u_lo, u_hi = np.log(x_lo), np.log(x_hi)
    du = (u_hi - u_lo) / N
    x = np.exp(np.arange(u_lo, u_hi, du))
    I1 = np.sum(du * _gnfwIntegrand(x, b, params))
    x_hi = np.exp(u_hi)
    I2 = x_lo ** (1 - G) / (1 - G) + x_hi ** (1 - B) / (1 - B)
    return I1 + I2


def calcTheta500Arcmin(z, M500, Om0=0.3, H0=70.0):
    """theta500 in arcmin for a flat LCDM cosmology (the reference asks pyccl for the same quantity)."""
    G = 4.30091e-9  # Mpc (km/s)^2 / MSun
    Ez = np.sqrt(Om0 * (1 + z) ** 3 + (1 - Om0))
    rhoCrit = 3 * (H0 * Ez) ** 2 / (8 * np.pi * G)
    R500Mpc = (3 * M500 / (4 * np.pi * 500 * rhoCrit)) ** (1 / 3.)
    DC = integrate.quad(lambda zz: 1.0 / np.sqrt(Om0 * (1 + zz) ** 3 + (1 - Om0)), 0, z)[0] * (_C_LIGHT / 1e3) / H0
    return np.degrees(np.arctan(R500Mpc / (DC / (1 + z)))) * 60.0


_profileCache = {}


def makeArnaudModelProfile(z, M500, GNFWParams='default'):
    """Spline knots of the projected, peak-normalised GNFW profile against angle in degrees."""
    if GNFWParams == 'default':
        GNFWParams = GNFW_DEFAULT_PARAMS
    key = (float(z), float(M500), tuple(sorted(GNFWParams.items())))
    if key in _profileCache:
        return _profileCache[key]
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
    out = {'tckP': interpolate.splrep(thetaDegRange, cylPProfile), 'theta500Arcmin': theta500Arcmin, 'rDeg': thetaDegRange}
    _profileCache[key] = out
    return out


def makeArnaudModelSignalMap(z, M500, degreesMap, wcs, beam=None, GNFWParams='default', amplitude=None,
                             maxSizeDeg=15.0, convolveWithBeam=True, beamDegreesMap=None):
    """Arnaud-model cluster template centred where degreesMap is zero, optionally convolved with the beam on the
    periodic tile grid.  beamDegreesMap: distance map whose zero is the integer pixel (H/2, W/2), for the beam kernel (so that the
    convolution does not move the cluster) when degreesMap is centred elsewhere; default degreesMap."""
    tckP = makeArnaudModelProfile(z, M500, GNFWParams=GNFWParams)['tckP']
    rDeg = np.linspace(0.0, maxSizeDeg, 5000)
    profile1d = interpolate.splev(rDeg, tckP, ext=1)
    if amplitude is not None:
        profile1d = profile1d * amplitude
    r2p = interpolate.interp1d(rDeg, profile1d, bounds_error=False, fill_value=0.0)
    signalMap = r2p(degreesMap)
    if convolveWithBeam:
        if beam is None:
            raise Exception("No beam supplied.")
        beamMap = makeBeamModelSignalMap(degreesMap if beamDegreesMap is None else beamDegreesMap, wcs, beam)
        beamMap = beamMap / beamMap.sum()
        signalMap = np.fft.ifft2(np.fft.fft2(signalMap) * np.fft.fft2(np.fft.ifftshift(beamMap))).real
    return signalMap


# ------------------------------------------------------------------------------------------------------
# Filter mismatch function Q (reference nemo/signals.py:864-1125 fitQ)
def calcM500FromTheta500Arcmin(theta500Arcmin, z, Om0=0.3, H0=70.0):
    """M500 (MSun) of a cluster of angular size theta500 at redshift z: the inverse of calcTheta500Arcmin, as signals.fitQ does it
    (signals.py:907-912: R500 = tan(theta500) D_A, M500 = 4/3 pi R500^3 500 rho_crit(z))."""
    G = 4.30091e-9  # Mpc (km/s)^2 / MSun
    Ez = np.sqrt(Om0 * (1 + z) ** 3 + (1 - Om0))
    rhoCrit = 3 * (H0 * Ez) ** 2 / (8 * np.pi * G)
    DC = integrate.quad(lambda zz: 1.0 / np.sqrt(Om0 * (1 + zz) ** 3 + (1 - Om0)), 0, z)[0] * (_C_LIGHT / 1e3) / H0
    R500Mpc = np.tan(np.radians(theta500Arcmin / 60.0)) * (DC / (1 + z))
    return (4 / 3.0) * np.pi * R500Mpc ** 3 * 500 * rhoCrit


def fitQScales(refM500MSun, refz):
    """(zRange, MRange) of the scales signals.fitQ filters for an Arnaud-profile reference filter (no redshift dependence of the
    profile shape, zDepQ = 0; signals.py:898-915): the reference scale first, then 54 sizes from 0.1' to 50' at log spacing
    0.05055349, each at a redshift that keeps the mass sensible."""
    theta500Arcmin_wanted = np.power(10, np.arange(np.log10(0.1), np.log10(50), 0.05055349))
    zRange_wanted = [2.0] * 10 + [1.0] * 10 + [0.6] * 10 + [0.3] * 10 + [0.1] * 10 + [0.07] * 4
    MRange = [refM500MSun] + [calcM500FromTheta500Arcmin(t, z) for t, z in zip(theta500Arcmin_wanted, zRange_wanted)]
    zRange = [refz] + zRange_wanted[:len(theta500Arcmin_wanted)]
    return zRange, MRange


def fitQ(config, **kwargs):
    """The filter mismatch function Q on a grid of cluster sizes for every tile (signals.py:864-1125); see
    nemo_b200.pipelines.fitQ, which runs it on the GPU."""
    from . import pipelines
    return pipelines.fitQ(config, **kwargs)
