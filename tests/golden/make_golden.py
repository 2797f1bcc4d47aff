"""
Generates the golden fixtures in this directory from the oracle (oracle/, a numpy/scipy restatement of the
reference's arithmetic -- the reference itself cannot be imported in this image, see oracle/__init__.py).

    python tests/golden/make_golden.py

Every fixture holds the seeded inputs' checksums, the filter parameters and the oracle's outputs for one small
tile, so that (a) the CPU suite can tell when the oracle's behaviour drifts and (b) the GPU suite can check the CUDA
path against committed numbers without running the oracle.  Inputs are regenerated from their seeds by
oracle.synth.makeTile; the fixture records sha256 of the input arrays to pin the generator too.
"""

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import synth, filters_ref, photometry_ref  # noqa: E402

DETECT = dict(threshold=4.0, minObjPix=1, findCenterOfMass=True, removeRings=True, ringThresholdSigma=3,
              useInterpolator=True, rejectBorder=0)

CASES = {
    'arnaud_f148_128': {
        'tile': dict(shape=(128, 128), numObjects=3, kind='clusters', seed=42, margin=36),
        'filter': {'label': 'Arnaud_M2e14_z0p4', 'class': 'ArnaudModelMatchedFilter',
                   'params': {'M500MSun': 2e14, 'z': 0.4, 'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': 20.},
                              'outputUnits': 'yc', 'edgeTrimArcmin': 0.}},
    },
    'beam_f148_128_bins': {
        'tile': dict(shape=(128, 128), freqsGHz=(148.0,), noiseUK=(100.0,), numObjects=6, kind='sources', seed=7,
                     borderPix=6, psHoles=1, margin=30),
        'filter': {'label': 'Beam', 'class': 'BeamMatchedFilter',
                   'params': {'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': 20., 'numNoiseBins': 3},
                              'outputUnits': 'uK', 'edgeTrimArcmin': 5.0}},
    },
    'arnaud_f098_f150_128': {
        'tile': dict(shape=(128, 128), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4),
                     numObjects=3, kind='clusters', cmb=True, seed=5, margin=36),
        'filter': {'label': 'Arnaud_M4e14_z0p4', 'class': 'ArnaudModelMatchedFilter',
                   'params': {'M500MSun': 4e14, 'z': 0.4, 'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': 20.},
                              'outputUnits': 'yc', 'edgeTrimArcmin': 0.}},
    },
}

CATALOG_KEYS = ['x', 'y', 'RADeg', 'decDeg', 'SNR', 'numSigPix', 'id', 'flags', 'galacticLatDeg']


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def copyFilter(f):
    return json.loads(json.dumps(f))


def runCase(name):
    case = CASES[name]
    mapDicts, truth = synth.makeTile(**case['tile'])
    if name == 'beam_f148_128_bins':
        rng = np.random.default_rng(99)
        mapDicts[0]['weights'] = mapDicts[0]['weights'] * (0.6 + 0.8 * rng.random(mapDicts[0]['weights'].shape))
    filtered, obj = filters_ref.filterMaps(mapDicts, copyFilter(case['filter']), 'T0', returnFilter=True)
    cat = photometry_ref.findObjects(filtered, **DETECT)
    photometry_ref.measureFluxes(cat, filtered, None)
    out = {'SNMap': filtered['SNMap'], 'data': filtered['data'], 'surveyMask': (np.asarray(filtered['surveyMask']) != 0).astype(np.uint8),
           'signalNorm': np.float64(obj.signalNorm), 'numObjects': np.int64(len(cat))}
    fluxKeys = ['y_c', 'err_y_c', 'deltaT_c', 'err_deltaT_c'] if filtered['mapUnits'] == 'yc' else ['deltaT_c', 'err_deltaT_c', 'fluxJy', 'err_fluxJy']
    for k in CATALOG_KEYS + fluxKeys:
        out['cat_' + k] = np.array([row[k] for row in cat], dtype=np.float64)
    out['cat_name'] = np.array([row['name'] for row in cat])
    seg = photometry_ref.getObjectPositions(filtered['SNMap'], DETECT['threshold'])[3]
    out['segmentation'] = seg.astype(np.int32)
    out['input_sha256'] = np.array([sha(m['data']) for m in mapDicts] + [sha(m['weights']) for m in mapDicts])
    if hasattr(obj, 'fRelWeights') and obj.fRelWeights:
        out['fRelWeights'] = np.array([[k, v] for k, v in obj.fRelWeights.items()], dtype=np.float64)
    return mapDicts, out


def main():
    for name in CASES:
        _, out = runCase(name)
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print("%s: %d objects, %.0f kB" % (name, int(out['numObjects']), os.path.getsize(path) / 1e3))


if __name__ == '__main__':
    main()
