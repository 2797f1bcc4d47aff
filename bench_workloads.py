"""
Seeded synthetic inputs for bench.py -- the BASELINE.json configurations as concretised in SURVEY.md section 8d.
Pure numpy; nothing here is on the measured path (map making is the reference's ``nemoModel`` / FITS reader, out
of scope).  The same host arrays are fed to the CUDA path and to the CPU baseline.

Config 2 ("configs[1]", the single-GPU bench workload): a 4096 x 4096, 148 GHz, 0.5' map cut into 16 tiles of
1024 x 1024, white noise of 100 uK per pixel at the median weight (tests/configs/sim_ptsrc_f090.yml style) with a
smooth cross-linking pattern in the weights (so that numNoiseBins = 8 has something to bin), 100 beam-convolved
point sources per tile with amplitudes ~ lognormal(ln 600, 1.1) uK (reference catalogs.py:646).
"""

import numpy as np

from nemo_b200 import signals
from nemo_b200 import maps as _maps
from nemo_b200.wcs import CARWCS

PIX_DEG = 0.5 / 60.0


def _paintSources(shape, xs, ys, amps, beam, pixDeg=PIX_DEG, radiusPix=24):
    out = np.zeros(shape, dtype=np.float64)
    H, W = shape
    for x, y, a in zip(xs, ys, amps):
        x0, x1 = max(int(x) - radiusPix, 0), min(int(x) + radiusPix + 1, W)
        y0, y1 = max(int(y) - radiusPix, 0), min(int(y) + radiusPix + 1, H)
        yy, xx = np.mgrid[y0:y1, x0:x1]
        r = np.sqrt((xx - x) ** 2 + (yy - y) ** 2) * pixDeg
        out[y0:y1, x0:x1] += a * np.interp(r, beam.rDeg, beam.profile1d, right=0.0)
    return out


def pointSourceTile(index, shape=(1024, 1024), freqGHz=148.0, noiseUK=100.0, FWHMArcmin=1.4, sourcesPerTile=100,
                    seed=1234, beam=None):
    """One tile of config 2: ([mapDict], truth)."""
    rng = np.random.default_rng([seed, index])
    H, W = shape
    if beam is None:
        beam = signals.gaussianBeam(FWHMArcmin)
    wcs = CARWCS(shape, PIX_DEG, crval=(10.0 - 0.1 * index, 0.0))
    yy, xx = np.mgrid[0:H, 0:W]
    ph = rng.uniform(0, 2 * np.pi, 2)
    depth = 0.75 + 0.25 * np.sin(2 * np.pi * 1.5 * yy / H + ph[0]) * np.cos(2 * np.pi * xx / W + ph[1])
    weights = depth / noiseUK ** 2
    data = rng.standard_normal(shape) / np.sqrt(weights)
    margin = 60
    xs = rng.uniform(margin, W - margin, sourcesPerTile)
    ys = rng.uniform(margin, H - margin, sourcesPerTile)
    amps = np.exp(rng.normal(np.log(600.0), 1.1, sourcesPerTile))
    data += _paintSources(shape, xs, ys, amps, beam)
    sigmaRad = np.radians(FWHMArcmin / 60.0) / np.sqrt(8 * np.log(2))
    mapDict = {'data': data, 'weights': weights, 'wcs': wcs, 'obsFreqGHz': freqGHz, 'units': 'uK', 'beam': beam,
               'beamFileName': beam, 'label': 'f%03d' % int(round(freqGHz)),
               'solidAngle_nsr': 2 * np.pi * sigmaRad ** 2 * 1e9,
               'surveyMask': np.ones(shape, dtype=np.uint8), 'pointSourceMask': np.ones(shape, dtype=np.uint8),
               'flagMask': np.zeros(shape, dtype=np.uint8)}
    return [mapDict], {'x': xs, 'y': ys, 'amp': amps}


def pointSourceConfig(numTiles=16, shape=(1024, 1024), firstIndex=0, seed=1234):
    """(parDict, tileMaps, truth) of config 2, parameters as tests/configs/sim_ptsrc_f090.yml of the reference."""
    beam = signals.gaussianBeam(1.4)
    tileMaps, truth = {}, {}
    for i in range(firstIndex, firstIndex + numTiles):
        name = '%d_%d' % (i // 4, i % 4)
        tileMaps[name], truth[name] = pointSourceTile(i, shape=shape, seed=seed, beam=beam)
    mapFilters = [{'label': 'Beam148', 'class': 'BeamMatchedFilter',
                   'params': {'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': 40.0, 'numNoiseBins': 8},
                              'outputUnits': 'uK', 'edgeTrimArcmin': 0.0}}]
    parDict = {'mapFilters': mapFilters, 'photFilter': None, 'thresholdSigma': 4.0, 'minObjPix': 1,
               'findCenterOfMass': True, 'useInterpolator': True, 'rejectBorder': 0, 'objIdent': 'MOCK-S',
               'longNames': False, 'removeRings': True, 'ringThresholdSigma': 3, 'undoPixelWindow': True,
               'catalogCuts': []}
    return parDict, tileMaps, truth


# ------------------------------------------------------------------------------------------------------
# Config 4 ("configs[3]", the bench workload): an ACT-DR5-scale cluster search -- 280 tiles of 840 x 1440 pixels (10 deg x 5 deg
# at 0.5' plus a 1 deg overlap on every side, the shape the DR5 tiling gives), 149.6 + 97.8 GHz, the 16 Arnaud_M*_z* filter scales
# with photFilter Arnaud_M2e14_z0p4 and the detection options of examples/ACT-DR5-clusters/DR5ClusterSearch.yml:8-16, 27-37,
# 79-126 of the reference.  Tiles sit on 10 declination rows x 28 right-ascension columns; CAR pixels shrink with cos(dec) along
# x, so every row has its own filter geometry (filters.py:214-239).
DR5_SHAPE = (840, 1440)
DR5_ROWS, DR5_COLS = 10, 28
DR5_FREQS = [(149.6, 40.0, 1.4), (97.8, 60.0, 2.2)]       # GHz, white noise uK per pixel, beam FWHM arcmin
DR5_MASSES = [1.0e14, 2.0e14, 4.0e14, 8.0e14]
DR5_REDSHIFTS = [(0.2, '0p2'), (0.4, '0p4'), (0.8, '0p8'), (1.2, '1p2')]


def cmbDl(ell):
    """Smooth stand-in for the lensed TT band powers D_l in uK^2 (the reference's planck_lensedCls.dat does not travel)."""
    ell = np.asarray(ell, dtype=np.float64)
    return 5000.0 * np.exp(-(ell / 1000.0) ** 1.25) * (1.0 - np.exp(-(ell / 30.0) ** 2)) + 1e-8


def _lmap2(shape, pixDegY, pixDegX):
    ly = 2 * np.pi * np.fft.fftfreq(shape[0], d=np.radians(pixDegY))
    lx = 2 * np.pi * np.fft.rfftfreq(shape[1], d=np.radians(pixDegX))
    return ly[:, None] ** 2 + lx[None, :] ** 2


_dr5Cache = {}


def _dr5Profiles(pixDegX):
    """Beam-convolved radial cluster profiles (beta models of four core radii) per frequency, tabulated against radius in pixels
    of a square 0.5' grid; x offsets are stretched by the tile's pixel scale when painting."""
    if 'profiles' in _dr5Cache:
        return _dr5Cache['profiles']
    n = 384
    yy, xx = np.mgrid[-n // 2:n // 2, -n // 2:n // 2]
    rDeg = np.hypot(yy, xx) * PIX_DEG
    l2 = _lmap2((n, n), PIX_DEG, PIX_DEG)
    out = {}
    for ic, thetaCArcmin in enumerate([0.4, 0.8, 1.4, 2.0]):
        img = (1.0 + (rDeg / (thetaCArcmin / 60.0)) ** 2) ** -1.0
        fimg = np.fft.rfft2(np.fft.ifftshift(img))
        for nu, _, fwhm in DR5_FREQS:
            sigma = np.radians(fwhm / 60.0) / np.sqrt(8 * np.log(2))
            sm = np.fft.fftshift(np.fft.irfft2(fimg * np.exp(-0.5 * l2 * sigma ** 2), s=(n, n)))
            out[(ic, nu)] = sm[n // 2, n // 2:] / sm[n // 2, n // 2]
    _dr5Cache['profiles'] = out
    return out


def dr5TileCentre(index):
    """(RA, dec) of the centre of tile `index` of the 10 x 28 layout."""
    row, col = (index // DR5_COLS) % DR5_ROWS, index % DR5_COLS
    return 5.0 + 10.0 * col, -47.5 + 5.0 * row


def dr5Tile(index, shape=DR5_SHAPE, seed=1234, clustersPerTile=20, dtype=np.float64):
    """One two-frequency tile of config 4: ([mapDict at 149.6 GHz, mapDict at 97.8 GHz], truth)."""
    rng = np.random.default_rng([seed, 4, index])
    H, W = shape
    raC, decC = dr5TileCentre(index)
    wcs = CARWCS(shape, PIX_DEG, crval=(raC, decC))
    pixDegX = PIX_DEG * np.cos(np.radians(decC))
    # CMB: flat-sky Gaussian realisation on the tile's own (anisotropic) pixel grid, one sky for both frequencies
    l2 = _lmap2(shape, PIX_DEG, pixDegX)
    lmap = np.sqrt(l2)
    lmap[0, 0] = 1.0
    Cl = 2 * np.pi * cmbDl(lmap) / (lmap * (lmap + 1))
    Cl[0, 0] = 0.0
    pixArea = np.radians(PIX_DEG) * np.radians(pixDegX)
    fsky = np.fft.rfft2(rng.standard_normal(shape)) * np.sqrt(Cl / pixArea)
    margin = 70
    xs = rng.uniform(margin, W - margin, clustersPerTile)
    ys = rng.uniform(margin, H - margin, clustersPerTile)
    y0s = rng.uniform(0.4e-4, 3e-4, clustersPerTile)
    cores = rng.integers(0, 4, clustersPerTile)
    profiles = _dr5Profiles(pixDegX)
    yy, xx = np.mgrid[0:H, 0:W]
    ph = rng.uniform(0, 2 * np.pi, 2)
    depth = 0.8 + 0.2 * np.sin(2 * np.pi * yy / H + ph[0]) * np.cos(2 * np.pi * 1.5 * xx / W + ph[1])
    stretch = np.cos(np.radians(decC))
    R = 100
    mapDicts = []
    for nu, noiseUK, fwhm in DR5_FREQS:
        sigma = np.radians(fwhm / 60.0) / np.sqrt(8 * np.log(2))
        m = np.fft.irfft2(fsky * np.exp(-0.5 * l2 * sigma ** 2), s=shape)
        weights = depth / noiseUK ** 2
        m += rng.standard_normal(shape) / np.sqrt(weights)
        amps = _maps.convertToDeltaT(y0s, nu)
        for x, y, a, ic in zip(xs, ys, amps, cores):
            x0, x1 = max(int(x) - R, 0), min(int(x) + R + 1, W)
            ya, yb = max(int(y) - R, 0), min(int(y) + R + 1, H)
            r = np.hypot((xx[ya:yb, x0:x1] - x) * stretch, yy[ya:yb, x0:x1] - y)
            prof = profiles[(int(ic), nu)]
            m[ya:yb, x0:x1] += a * np.interp(r, np.arange(len(prof)), prof, right=0.0)
        beam = _dr5Cache.setdefault(('beam', fwhm), signals.gaussianBeam(fwhm))
        mapDicts.append({'data': m.astype(dtype), 'weights': weights.astype(dtype), 'wcs': wcs, 'obsFreqGHz': nu, 'units': 'uK',
                         'beam': beam, 'beamFileName': beam, 'label': 'f%03d' % int(round(nu)),
                         'solidAngle_nsr': 2 * np.pi * sigma ** 2 * 1e9,
                         'surveyMask': np.ones(shape, dtype=np.uint8), 'pointSourceMask': np.ones(shape, dtype=np.uint8),
                         'flagMask': np.zeros(shape, dtype=np.uint8)})
    return mapDicts, {'x': xs, 'y': ys, 'y0': y0s}


def dr5TileName(index):
    return '%d_%d' % (index // DR5_COLS, index % DR5_COLS)


def dr5Filters(numScales=16):
    """The mapFilters list of DR5ClusterSearch.yml:93-126 (allFilters merged in), in file order."""
    mapFilters = []
    for z, zs in DR5_REDSHIFTS:
        for M in DR5_MASSES:
            mapFilters.append({'label': 'Arnaud_M%de14_z%s' % (int(round(M / 1e14)), zs), 'class': 'ArnaudModelMatchedFilter',
                               'params': {'M500MSun': M, 'z': z, 'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': 40.0},
                                          'outputUnits': 'yc', 'edgeTrimArcmin': 0.0}})
    if numScales < 16:      # keep the photFilter in a shortened list
        phot = [f for f in mapFilters if f['label'] == 'Arnaud_M2e14_z0p4']
        rest = [f for f in mapFilters if f['label'] != 'Arnaud_M2e14_z0p4']
        mapFilters = phot + rest[:numScales - 1]
    return mapFilters


def dr5ParDict(numScales=16):
    return {'mapFilters': dr5Filters(numScales), 'photFilter': 'Arnaud_M2e14_z0p4', 'thresholdSigma': 4.0, 'minObjPix': 1,
            'findCenterOfMass': True, 'useInterpolator': True, 'rejectBorder': 0, 'objIdent': 'ACT-CL', 'longNames': False,
            'removeRings': True, 'ringThresholdSigma': 3, 'undoPixelWindow': True, 'catalogCuts': []}


def dr5Config(numTiles=35, shape=DR5_SHAPE, firstIndex=0, seed=1234, numScales=16, dtype=np.float64):
    """(parDict, tileMaps, truth) of config 4 for tiles firstIndex ... firstIndex + numTiles - 1."""
    tileMaps, truth = {}, {}
    for i in range(firstIndex, firstIndex + numTiles):
        tileMaps[dr5TileName(i)], truth[dr5TileName(i)] = dr5Tile(i, shape=shape, seed=seed, dtype=dtype)
    return dr5ParDict(numScales), tileMaps, truth
