"""
Full-size golden cases: BASELINE.json configs[0] (one 1024 x 1024 tile at 148 GHz, ArnaudModelMatchedFilter theta500 = 2.4') and
one tile of configs[3] (840 x 1440, 149.6 + 97.8 GHz, the photFilter Arnaud_M2e14_z0p4 and a second scale with the photFilter
flow of nemo/pipelines.py:238-284).  ``make_reference_golden.py`` runs the REFERENCE's own filters.py / photometry.py on these
tiles and stores a compact form of what it computed in ``ref_big_<case>.npz``; the oracle (tests/test_oracle.py) and the CUDA
path (tests/test_golden_gpu.py) are compared with those files.

Compact form of one filtered tile (full maps would be 10 MB each): every 8th pixel of the S/N map and of the filtered map in
both directions, their full-resolution row / column sums and sums of squares, the area mask bit-packed, the segmentation map as
(pixel index, label) pairs with the S/N at those pixels, the normalisation, and the catalog column by column.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DETECT = dict(threshold=4.0, minObjPix=1, findCenterOfMass=True, removeRings=True, ringThresholdSigma=3,
              useInterpolator=True, rejectBorder=0)


def _arnaud(label, M, z, grid=40.):
    return {'label': label, 'class': 'ArnaudModelMatchedFilter',
            'params': {'M500MSun': M, 'z': z, 'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': grid},
                       'outputUnits': 'yc', 'edgeTrimArcmin': 0.}}


def _tileConfig0():
    from oracle import synth
    return synth.makeTile(shape=(1024, 1024), freqsGHz=(148.0,), noiseUK=(50.0,), beamFWHMArcmin=(1.4,), seed=1234, kind='clusters',
                          numObjects=30)[0]


def _tileDR5():
    import bench_workloads
    from oracle.wcs_stub import SimpleCARWCS
    mapDicts, _ = bench_workloads.dr5Tile(29)           # second declination row: dec -42.5, x pixel scale 0.74 of y
    wcs = SimpleCARWCS(bench_workloads.DR5_SHAPE, bench_workloads.PIX_DEG, crval=mapDicts[0]['wcs'].crval)
    for m in mapDicts:
        m['wcs'] = wcs
    return mapDicts


BIG_CASES = {
    # configs[0]: quickstart-clusters.yml style single tile
    'arnaud_f148_1024': {'tile': _tileConfig0, 'filters': [_arnaud('Arnaud_M2e14_z0p4', 2e14, 0.4)], 'photFilter': None},
    # one tile of configs[3]: the photFilter first, then another scale of the DR5 list (reference loop order)
    'dr5_f150_f098_840x1440': {'tile': _tileDR5, 'filters': [_arnaud('Arnaud_M2e14_z0p4', 2e14, 0.4), _arnaud('Arnaud_M8e14_z0p2', 8e14, 0.2)],
                               'photFilter': 'Arnaud_M2e14_z0p4'},
}


def compact(prefix, data, SNMap, surveyMask, seg, signalNorm):
    data, SNMap = np.asarray(data, dtype=np.float64), np.asarray(SNMap, dtype=np.float64)
    seg = np.asarray(seg)
    idx = np.flatnonzero(seg.ravel() > 0).astype(np.int32)
    out = {'SN_sub': SNMap[::8, ::8].copy(), 'data_sub': data[::8, ::8].copy(),
           'SN_rowsum': SNMap.sum(axis=1), 'SN_colsum': SNMap.sum(axis=0), 'SN_sumsq': np.float64((SNMap ** 2).sum()),
           'data_rowsum': data.sum(axis=1), 'data_colsum': data.sum(axis=0), 'data_sumsq': np.float64((data ** 2).sum()),
           'mask_bits': np.packbits(np.asarray(surveyMask).ravel() != 0), 'seg_idx': idx, 'seg_lab': seg.ravel()[idx].astype(np.int32),
           'SN_at_seg': SNMap.ravel()[idx], 'signalNorm': np.float64(signalNorm)}
    return {prefix + k: v for k, v in out.items()}


def catalogColumns(prefix, cat):
    """Catalog (Table or list of row dictionaries) -> {prefix + column: array}."""
    out = {}
    if isinstance(cat, list):
        keys = list(cat[0].keys()) if len(cat) > 0 else []
        cols = {k: np.array([row[k] for row in cat]) for k in keys}
    else:
        cols = {k: np.asarray(cat[k]) for k in cat.keys()}
    for k, col in cols.items():
        out[prefix + k] = col.astype(str) if col.dtype == object or col.dtype.kind in 'US' else col.astype(np.float64)
    return out


def compareCompact(got, ref, prefix, rtolSN=1e-4, atolSN=1e-4, exact=False):
    """Asserts a compact record `got` against the fixture `ref` (both keyed prefix + name).  exact: bit for bit (oracle against
    reference run); else the north-star tolerances (S/N 1e-4, filtered map 1e-4 of its range, mask and labels bit-exact)."""
    g = lambda k: np.asarray(got[prefix + k])
    r = lambda k: np.asarray(ref[prefix + k])
    assert np.array_equal(g('mask_bits'), r('mask_bits')), "area masks differ"
    assert np.array_equal(g('seg_idx'), r('seg_idx')) and np.array_equal(g('seg_lab'), r('seg_lab')), "segmentation maps differ"
    if exact:
        for k in ('SN_sub', 'data_sub', 'SN_rowsum', 'SN_colsum', 'SN_sumsq', 'data_rowsum', 'data_colsum', 'data_sumsq', 'SN_at_seg', 'signalNorm'):
            assert np.array_equal(g(k), r(k)), k
        return
    assert np.allclose(g('SN_sub'), r('SN_sub'), rtol=rtolSN, atol=atolSN)
    assert np.allclose(g('SN_at_seg'), r('SN_at_seg'), rtol=rtolSN, atol=atolSN)
    scale = np.abs(r('data_sub')).max()
    assert np.abs(g('data_sub') - r('data_sub')).max() < 1e-4 * scale
    # full-resolution sums: a row of W pixels with per-pixel error e sums to at most W e
    n = max(len(r('SN_rowsum')), len(r('SN_colsum')))
    assert np.allclose(g('SN_rowsum'), r('SN_rowsum'), rtol=0, atol=1e-4 * n) and np.allclose(g('SN_colsum'), r('SN_colsum'), rtol=0, atol=1e-4 * n)
    assert np.allclose(g('data_rowsum'), r('data_rowsum'), rtol=0, atol=1e-4 * scale * n)
    assert np.allclose(g('data_colsum'), r('data_colsum'), rtol=0, atol=1e-4 * scale * n)
    assert abs(float(g('SN_sumsq')) / float(r('SN_sumsq')) - 1) < 1e-6 and abs(float(g('data_sumsq')) / float(r('data_sumsq')) - 1) < 1e-6
    assert abs(float(g('signalNorm')) / float(r('signalNorm')) - 1) < 1e-8


def assertNamesMatch(gotNames, refNames, refRADeg, refDecDeg, epsDeg=1e-9):
    """Names are positions with truncated digits (catalogs.py:259-274), so a position that sits exactly on a digit boundary
    (a centroid on a pixel row of a CAR map is a whole multiple of the pixel size) flips the last digit on a 1-ulp difference.
    Names must agree, except that for such a position the name of the reference position moved by +-epsDeg is accepted too."""
    from nemo_b200 import catalogs
    for got, want, ra, dec in zip(gotNames, refNames, np.asarray(refRADeg, dtype=float), np.asarray(refDecDeg, dtype=float)):
        if str(got) == str(want):
            continue
        prefix = str(want).split(' J')[0]
        near = {catalogs.makeName(ra + a, dec + b, prefix=prefix) for a in (-epsDeg, 0.0, epsDeg) for b in (-epsDeg, 0.0, epsDeg)}
        assert str(got) in near, "name %s differs from the reference's %s (RA %.12f, dec %.12f)" % (got, want, ra, dec)
    assert len(gotNames) == len(refNames)


def runOracle(name):
    """The oracle on a big case: {compact records and catalogs per filter, keyed 's<i>_...'} in the reference's loop order."""
    import copy
    from oracle import filters_ref, photometry_ref
    case = BIG_CASES[name]
    mapDicts = case['tile']()
    out, phot = {}, None
    for i, f in enumerate(case['filters']):
        filtered, obj = filters_ref.filterMaps(mapDicts, copy.deepcopy(f), 'T0', returnFilter=True)
        if f['label'] == case['photFilter']:
            phot = {'SNMap': filtered['SNMap'], 'data': filtered['data']}
        seg = photometry_ref.getObjectPositions(filtered['SNMap'], DETECT['threshold'])[3]
        cat = photometry_ref.findObjects(filtered, **DETECT)
        if phot is not None:
            photometry_ref.getSNRValues(cat, phot['SNMap'], filtered['wcs'], prefix='fixed_', useInterpolator=DETECT['useInterpolator'])
        photometry_ref.measureFluxes(cat, filtered, None, photFilteredMapDict=phot, useInterpolator=DETECT['useInterpolator'])
        out.update(compact('s%d_' % i, filtered['data'], filtered['SNMap'], filtered['surveyMask'], seg, obj.signalNorm))
        out.update(catalogColumns('s%d_cat_' % i, cat))
        if obj.fRelWeights:
            out['s%d_fRel' % i] = np.array([obj.fRelWeights[k] for k in sorted(obj.fRelWeights.keys())], dtype=np.float64)
    return mapDicts, out
