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
