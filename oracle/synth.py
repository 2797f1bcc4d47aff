"""
Seeded synthetic inputs for the BASELINE.json configurations: CAR tiles at 0.5'
pixels, white noise, an optional Gaussian-random CMB-like sky, injected beam-convolved point sources or
Arnaud-profile clusters.  These are INPUT generators shared by the tests, the bench and the oracle; the
arrays they return are fed unchanged to both the CUDA path and the oracle.  TEST INFRASTRUCTURE -- see
oracle/__init__.py.

The CMB spectrum is an analytic damped fit (the reference's nemo/data/planck_lensedCls.dat does not
travel to the GPU box and reference files are not copied into this repo).
"""

import numpy as np
from scipy import interpolate

from . import signals_ref as signals
from .wcs_stub import SimpleCARWCS

PIX_DEG = 0.5 / 60.0


def cmbDl(ell):
    """Rough lensed-TT band power D_l = l(l+1)C_l/2pi in uK^2: acoustic-peak-free damped plateau."""
    ell = np.asarray(ell, dtype=np.float64)
    return 5000.0 * np.exp(-(ell / 1000.0) ** 1.25) * (1.0 - np.exp(-(ell / 30.0) ** 2)) + 1e-8


def simCMB(shape, rng, pixDeg=PIX_DEG):
    """Flat-sky Gaussian realisation with C_l = 2 pi D_l / (l (l+1))."""
    H, W = shape
    pixRad = np.radians(pixDeg)
    ly = 2 * np.pi * np.fft.fftfreq(H, d=pixRad)
    lx = 2 * np.pi * np.fft.fftfreq(W, d=pixRad)
    lmap = np.sqrt(ly[:, None] ** 2 + lx[None, :] ** 2)
    lmap[0, 0] = 1.0
    Cl = 2 * np.pi * cmbDl(lmap) / (lmap * (lmap + 1))
    Cl[0, 0] = 0.0
    white = rng.standard_normal((H, W))
    fw = np.fft.fft2(white)
    # P(k) per pixel^2: C_l / pixel area
    return np.fft.ifft2(fw * np.sqrt(Cl / pixRad ** 2)).real


def beamSmooth(m, FWHMArcmin, pixDeg=PIX_DEG):
    H, W = m.shape
    pixRad = np.radians(pixDeg)
    ly = 2 * np.pi * np.fft.fftfreq(H, d=pixRad)
    lx = 2 * np.pi * np.fft.fftfreq(W, d=pixRad)
    l2 = ly[:, None] ** 2 + lx[None, :] ** 2
    sigma = np.radians(FWHMArcmin / 60.0) / np.sqrt(8 * np.log(2))
    return np.fft.ifft2(np.fft.fft2(m) * np.exp(-0.5 * l2 * sigma ** 2)).real


def paintProfiles(shape, xs, ys, amps, rDeg, profile1d, pixDeg=PIX_DEG, maxRadiusPix=120):
    """Adds amps[i] * profile(r) at sub-pixel positions (xs[i], ys[i]); linear interpolation of the 1-D
    radial profile on a stamp of +/- maxRadiusPix."""
    out = np.zeros(shape, dtype=np.float64)
    r2p = interpolate.interp1d(rDeg, profile1d, bounds_error=False, fill_value=0.0)
    H, W = shape
    for x, y, a in zip(xs, ys, amps):
        x0 = max(int(x) - maxRadiusPix, 0)
        x1 = min(int(x) + maxRadiusPix + 1, W)
        y0 = max(int(y) - maxRadiusPix, 0)
        y1 = min(int(y) + maxRadiusPix + 1, H)
        yy, xx = np.mgrid[y0:y1, x0:x1]
        r = np.sqrt((xx - x) ** 2 + (yy - y) ** 2) * pixDeg
        out[y0:y1, x0:x1] += a * r2p(r)
    return out


def beamConvolvedProfile(rDeg, profile1d, FWHMArcmin, pixDeg=PIX_DEG, n=512):
    """Radial profile after convolution with a Gaussian beam, by rendering on an n x n grid."""
    yy, xx = np.mgrid[-n // 2:n // 2, -n // 2:n // 2]
    r = np.sqrt(xx ** 2 + yy ** 2) * pixDeg
    img = interpolate.interp1d(rDeg, profile1d, bounds_error=False, fill_value=0.0)(r)
    img = beamSmooth(np.fft.ifftshift(img), FWHMArcmin, pixDeg)
    img = np.fft.fftshift(img)
    return np.arange(n // 2) * pixDeg, img[n // 2, n // 2:]


def makeMapDict(data, weights, wcs, obsFreqGHz, beam, label, units='uK', surveyMask=None, psMask=None,
                flagMask=None, solidAngle_nsr=0.0):
    shape = data.shape
    return {'data': data, 'weights': weights, 'wcs': wcs, 'obsFreqGHz': obsFreqGHz, 'units': units,
            'beam': beam, 'beamFileName': beam, 'label': label, 'solidAngle_nsr': solidAngle_nsr,
            'surveyMask': np.ones(shape, dtype=np.uint8) if surveyMask is None else surveyMask,
            'pointSourceMask': np.ones(shape, dtype=np.uint8) if psMask is None else psMask,
            'flagMask': np.zeros(shape, dtype=np.uint8) if flagMask is None else flagMask}


def makeTile(shape=(1024, 1024), freqsGHz=(148.0,), noiseUK=(50.0,), beamFWHMArcmin=(1.4,), seed=1234,
             kind='clusters', numObjects=30, cmb=False, borderPix=0, psHoles=0, theta500Arcmin=2.4,
             clusterProfile=None, dtype=np.float64, margin=60):
    """One synthetic tile: list of per-frequency map dicts + truth table.

    kind = 'clusters' : Arnaud-like decrements, y0 in [0.5, 5]e-4 converted per frequency (config 1, 3-5)
           'sources'  : positive point sources, deltaT ~ lognormal(ln 600, 1.1) uK (config 2)
    borderPix > 0 zeroes a border of that width in data/weights/surveyMask (a tile with a zero edge, which
    is what edgeTrimArcmin acts on, filters.py:725-727).  psHoles punches that many 8-px-radius holes in
    the point-source mask.  Objects are kept `margin` pixels away from the (zeroed) edge.
    """
    rng = np.random.default_rng(seed)
    H, W = shape
    wcs = SimpleCARWCS(shape, PIX_DEG)
    margin = margin + borderPix
    xs = rng.uniform(margin, W - margin, numObjects)
    ys = rng.uniform(margin, H - margin, numObjects)
    if kind == 'clusters':
        y0s = rng.uniform(0.5e-4, 5e-4, numObjects)
        if clusterProfile is None:
            # beta-model stand-in for the GNFW line-of-sight profile when none is supplied (cheap)
            rDeg = np.linspace(0, 0.5, 3000)
            thetaC = theta500Arcmin / 60.0 / 3.0
            clusterProfile = (rDeg, (1 + (rDeg / thetaC) ** 2) ** -1.0)
    else:
        amps = np.exp(rng.normal(np.log(600.0), 1.1, numObjects))
    cmbMap = simCMB(shape, rng) if cmb else None
    psMask = np.ones(shape, dtype=np.uint8)
    for _ in range(psHoles):
        hx, hy = rng.uniform(margin, W - margin), rng.uniform(margin, H - margin)
        yy, xx = np.ogrid[:H, :W]
        psMask[(xx - hx) ** 2 + (yy - hy) ** 2 < 64] = 0
    surveyMask = np.ones(shape, dtype=np.uint8)
    if borderPix > 0:
        surveyMask[:borderPix, :] = 0
        surveyMask[-borderPix:, :] = 0
        surveyMask[:, :borderPix] = 0
        surveyMask[:, -borderPix:] = 0
    mapDicts = []
    for i, (nu, sig, fwhm) in enumerate(zip(freqsGHz, noiseUK, beamFWHMArcmin)):
        beam = signals.gaussianBeam(fwhm)
        m = rng.standard_normal(shape) * sig
        if cmbMap is not None:
            m = m + beamSmooth(cmbMap, fwhm)
        if kind == 'clusters':
            r, p = beamConvolvedProfile(clusterProfile[0], clusterProfile[1], fwhm)
            m = m + paintProfiles(shape, xs, ys, signals.convertToDeltaT(y0s, nu), r, p)
        else:
            m = m + paintProfiles(shape, xs, ys, amps, beam.rDeg, beam.profile1d, maxRadiusPix=30)
        wts = np.ones(shape) / sig ** 2
        if borderPix > 0:
            m = m * surveyMask
            wts = wts * surveyMask
        sigmaRad = np.radians(fwhm / 60.0) / np.sqrt(8 * np.log(2))
        mapDicts.append(makeMapDict(m.astype(dtype), wts.astype(dtype), wcs, nu, beam, 'f%03d' % int(round(nu)),
                                    surveyMask=surveyMask, psMask=psMask,
                                    solidAngle_nsr=2 * np.pi * sigmaRad ** 2 * 1e9))
    truth = {'x': xs, 'y': ys}
    truth['amp'] = y0s if kind == 'clusters' else amps
    return mapDicts, truth
