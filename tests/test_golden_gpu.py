"""
The CUDA path against the committed golden fixtures (tests/golden/*.npz, produced by the oracle with
tests/golden/make_golden.py): S/N maps within 1e-4, survey masks and segmentation labels bit-exact, catalog identical
above the S/N cut with positions within 0.01 px.  Inputs are regenerated from their seeds and checked against the
recorded checksums; the oracle's filtering code is not run here.  The same outputs are also checked against the fixtures
ref_<case>.npz, which hold what the REFERENCE's own source files computed for these tiles (tests/golden/make_reference_golden.py).
"""

import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import make_golden  # noqa: E402


@pytest.mark.parametrize("name", sorted(make_golden.CASES.keys()))
def test_cuda_path_matches_golden_fixture(name, report):
    from oracle import synth
    from nemo_b200 import filters, photometry
    gold = np.load(os.path.join(HERE, 'golden', name + '.npz'))
    case = make_golden.CASES[name]
    mapDicts, _ = synth.makeTile(**case['tile'])
    if name == 'beam_f148_128_bins':
        rng = np.random.default_rng(99)
        mapDicts[0]['weights'] = mapDicts[0]['weights'] * (0.6 + 0.8 * rng.random(mapDicts[0]['weights'].shape))
    shas = [make_golden.sha(m['data']) for m in mapDicts] + [make_golden.sha(m['weights']) for m in mapDicts]
    assert shas == list(gold['input_sha256'])
    got, obj = filters.filterMaps(mapDicts, make_golden.copyFilter(case['filter']), 'T0', verbose=False, returnFilter=True)
    sn = np.asarray(got['SNMap'])
    assert np.array_equal((np.asarray(got['surveyMask']) != 0).astype(np.uint8), gold['surveyMask'])
    errSN = float(np.abs(sn - gold['SNMap']).max())
    errData = float(np.abs(np.asarray(got['data']) - gold['data']).max() / np.abs(gold['data']).max())
    report("golden_" + name, maxabs_dSN=errSN, rel_dData=errData)
    assert np.allclose(sn, gold['SNMap'], rtol=1e-4, atol=1e-4)
    assert errData < 1e-4
    assert abs(obj.signalNorm / float(gold['signalNorm']) - 1) < 1e-8
    seg = photometry.getObjectPositions(got['SNMap'], make_golden.DETECT['threshold'])[3]
    assert np.array_equal(seg, gold['segmentation'])
    cat = photometry.findObjects(got, **make_golden.DETECT)
    photometry.measureFluxes(cat, got, None)
    assert len(cat) == int(gold['numObjects'])
    assert np.abs(np.asarray(cat['x']) - gold['cat_x']).max() < 0.01 and np.abs(np.asarray(cat['y']) - gold['cat_y']).max() < 0.01
    assert list(cat['name']) == list(gold['cat_name'])
    for k in gold.files:
        if k.startswith('cat_') and k not in ('cat_name', 'cat_x', 'cat_y'):
            np.testing.assert_allclose(np.asarray(cat[k[4:]], dtype=np.float64), gold[k], rtol=1e-4, atol=1e-9)
    if 'fRelWeights' in gold.files:
        for freq, wgt in gold['fRelWeights']:
            assert abs(obj.fRelWeights[float(freq)] - wgt) < 1e-8


@pytest.mark.parametrize("name", sorted(make_golden.CASES.keys()))
def test_cuda_path_matches_reference_run(name, report):
    """The CUDA path against runs of the reference's own code (nemo/filters.py, nemo/photometry.py imported from /root/reference
    when the fixture was made): S/N map within 1e-4, filtered map, area mask and labels, object positions within 0.01 px, the
    'smart' and 'percentile' noise maps."""
    from nemo_b200 import filters, photometry
    ref = np.load(os.path.join(HERE, 'golden', 'ref_' + name + '.npz'))
    case = make_golden.CASES[name]
    from oracle import synth
    mapDicts, _ = synth.makeTile(**case['tile'])
    if name == 'beam_f148_128_bins':
        rng = np.random.default_rng(99)
        mapDicts[0]['weights'] = mapDicts[0]['weights'] * (0.6 + 0.8 * rng.random(mapDicts[0]['weights'].shape))
    got, obj = filters.filterMaps(mapDicts, make_golden.copyFilter(case['filter']), 'T0', verbose=False, returnFilter=True)
    sn = np.asarray(got['SNMap'])
    assert np.array_equal(np.asarray(got['surveyMask']) != 0, ref['surveyMask'] != 0)
    assert np.allclose(sn, ref['SNMap'], rtol=1e-4, atol=1e-4)
    errData = float(np.abs(np.asarray(got['data']) - ref['data']).max() / np.abs(ref['data']).max())
    assert errData < 1e-4
    assert abs(obj.signalNorm / float(ref['signalNorm']) - 1) < 1e-8
    ids, pos, numPix, seg = photometry.getObjectPositions(got['SNMap'], 4.0)
    assert np.array_equal(seg, ref['segmentationMap']) and np.array_equal(ids, ref['objIDs'])
    assert np.array_equal(np.asarray(numPix, dtype=np.float64), np.asarray(ref['objNumPix'], dtype=np.float64))
    p = np.asarray(pos, dtype=np.float64).reshape(-1, 2)
    assert np.abs(p[1:] - ref['objPositions'][1:]).max() < 0.01          # entry 0 is the background label
    # the catalog of the reference's findObjects + measureFluxes on its own maps
    cat = photometry.findObjects(got, **make_golden.DETECT)
    photometry.measureFluxes(cat, got, None)
    assert len(cat) == len(ref['refcat_RADeg'])
    assert list(cat['name']) == list(ref['refcat_name'])
    pixDeg = mapDicts[0]['wcs'].getPixelSizeDeg()
    assert np.abs(np.asarray(cat['RADeg'], dtype=float) - ref['refcat_RADeg']).max() < 0.01 * pixDeg
    assert np.abs(np.asarray(cat['decDeg'], dtype=float) - ref['refcat_decDeg']).max() < 0.01 * pixDeg
    assert np.array_equal(np.asarray(cat['numSigPix'], dtype=float), ref['refcat_numSigPix'])
    for k in [k[7:] for k in ref.files if k.startswith('refcat_')]:
        if k not in ('name', 'template', 'tileName', 'RADeg', 'decDeg', 'numSigPix'):
            np.testing.assert_allclose(np.asarray(cat[k], dtype=np.float64), ref['refcat_' + k], rtol=1e-4, atol=1e-9, err_msg=k)
    # the other noise estimators, on the map the fixture used (3 x the reference's S/N map)
    for key, noiseParams in (('smart', {'method': 'dataMap', 'noiseGridArcmin': 'smart', 'numNoiseBins': 6}),
                             ('percentile', {'method': 'dataMap', 'noiseGridArcmin': 20., 'numNoiseBins': 2 if name.startswith('beam') else 1,
                                             'RMSEstimator': 'percentile'})):
        params = dict(make_golden.copyFilter(case['filter'])['params'], noiseParams=noiseParams)
        f = filters.MapFilter('t', mapDicts, params)
        rms = f.makeNoiseMap(ref['SNMap'] * 3.0 + 0.0)
        assert np.array_equal(rms == 0, ref['noise_' + key] == 0), key
        assert np.allclose(rms, ref['noise_' + key], rtol=1e-10, atol=0), key
    report("reference_run_" + name, maxabs_dSN=float(np.abs(sn - ref['SNMap']).max()), rel_dData=errData, labels=int(len(ids) - 1))


def test_real_space_apply_matches_reference_run(report):
    """RealSpaceMatchedFilter.applyFilter on the device (Gaussian high-pass K12 + kernel convolution K13) against what the
    reference's own applyFilter (nemo/filters.py:1172-1218, with maps.subtractBackground / smoothMap) computed for the seeded
    cube and kernel of ref_misc.npz."""
    import make_reference_golden as mrg
    from nemo_b200 import filters, device
    from nemo_b200.wcs import CARWCS
    ref = np.load(os.path.join(HERE, 'golden', 'ref_misc.npz'))
    inp = mrg.miscInputs()
    obj = object.__new__(filters.RealSpaceMatchedFilter)
    obj.ops = device.getOps('f64')
    obj.params = {'bckSub': True}
    obj.wcs = CARWCS(inp['shape'], 0.5 / 60.0, crval=(30.0, -20.0))
    obj.kern2d, obj.signalNorm = inp['kern'], inp['signalNorm']
    obj.applyRACentre, obj.applyDecCentre, obj.bckSubScaleArcmin = inp['applyRADeg'], inp['applyDecDeg'], inp['bckSubScaleArcmin']
    obj.unfilteredMapsDictList = [{'obsFreqGHz': 97.8}, {'obsFreqGHz': 149.6}]
    got = obj.applyFilter(inp['cube'].copy(), calcFRelWeights=True)
    scale = np.abs(ref['rs_filtered']).max()
    err = float(np.abs(got - ref['rs_filtered']).max() / scale)
    assert err < 1e-12
    assert np.allclose([obj.fRelWeights[97.8], obj.fRelWeights[149.6]], ref['rs_fRel'], rtol=1e-10, atol=0)
    obj.params = {'bckSub': False}
    errNo = float(np.abs(obj.applyFilter(inp['cube'].copy()) - ref['rs_filteredNoBck']).max() / np.abs(ref['rs_filteredNoBck']).max())
    assert errNo < 1e-12
    report("reference_run_real_space_apply", rel_err=err, rel_err_no_bck=errNo)


def test_phot_filter_flow_matches_reference_run(report):
    """The batched driver with a photFilter (reference filter first, fixed_SNR and fixed_ fluxes for the other scale) against the
    reference's own findObjects + getSNRValues + measureFluxes on that tile (ref_misc.npz phot_*)."""
    import copy
    import make_reference_golden as mrg
    from nemo_b200 import pipelines
    ref = np.load(os.path.join(HERE, 'golden', 'ref_misc.npz'))
    mapDicts, _ = make_golden.runCase('arnaud_f148_128')
    f0, f1 = copy.deepcopy(make_golden.CASES['arnaud_f148_128']['filter']), copy.deepcopy(mrg.PHOT_SECOND_SCALE)
    parDict = {'mapFilters': [f1, f0], 'photFilter': f0['label'], 'thresholdSigma': make_golden.DETECT['threshold'],
               'minObjPix': make_golden.DETECT['minObjPix'], 'findCenterOfMass': True, 'useInterpolator': True, 'rejectBorder': 0,
               'removeRings': True, 'ringThresholdSigma': make_golden.DETECT['ringThresholdSigma']}
    config = pipelines.SimpleConfig(parDict, {'T0': mapDicts})
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=1)
    cat = config.catalogDict[f1['label'] + '#T0']['catalog']
    assert len(cat) == len(ref['phot_SNR'])
    assert list(cat['name']) == list(ref['phot_name'])
    for k in [k[5:] for k in ref.files if k.startswith('phot_')]:
        if k in ('name', 'template', 'tileName'):
            continue
        np.testing.assert_allclose(np.asarray(cat[k], dtype=np.float64), ref['phot_' + k], rtol=1e-4, atol=1e-9, err_msg=k)
    report("reference_run_phot_filter", rows=len(cat), max_rel_fixed_SNR=float(np.abs(np.asarray(cat['fixed_SNR'], dtype=float) / ref['phot_fixed_SNR'] - 1).max()))


@pytest.mark.parametrize("name", ['arnaud_f148_1024', 'dr5_f150_f098_840x1440'])
def test_cuda_path_matches_full_size_reference_run(name, report):
    """BASELINE-size tiles against runs of the reference's own code (ref_big_<case>.npz, compact form, tests/golden/big_cases.py):
    configs[0] -- one 1024 x 1024 tile, ArnaudModelMatchedFilter -- and one 840 x 1440 two-frequency tile of configs[3] at
    declination -42.5 (x pixel scale 0.74 of y) with the photFilter flow of nemo/pipelines.py:238-284.  Class API, reference loop
    order; S/N within 1e-4, masks and labels bit-exact, catalogs identical with positions within 0.01 px."""
    import copy
    import big_cases
    from nemo_b200 import filters, photometry
    ref = np.load(os.path.join(HERE, 'golden', 'ref_big_' + name + '.npz'))
    case = big_cases.BIG_CASES[name]
    mapDicts = case['tile']()
    D = big_cases.DETECT
    phot = None
    for i, f in enumerate(case['filters']):
        got, obj = filters.filterMaps(mapDicts, copy.deepcopy(f), 'T0', verbose=False, returnFilter=True)
        if f['label'] == case['photFilter']:
            phot = {'SNMap': got['SNMap'], 'data': got['data']}
        seg = photometry.getObjectPositions(got['SNMap'], D['threshold'])[3]
        cat = photometry.findObjects(got, **D)
        if phot is not None:
            photometry.getSNRValues(cat, phot['SNMap'], got['wcs'], prefix='fixed_', useInterpolator=D['useInterpolator'])
        photometry.measureFluxes(cat, got, None, photFilteredMapDict=phot, useInterpolator=D['useInterpolator'])
        prefix = 's%d_' % i
        rec = big_cases.compact(prefix, got['data'], got['SNMap'], got['surveyMask'], seg, obj.signalNorm)
        big_cases.compareCompact(rec, ref, prefix)
        assert len(cat) == len(ref[prefix + 'cat_RADeg']) >= 20
        big_cases.assertNamesMatch(list(cat['name']), list(ref[prefix + 'cat_name']), ref[prefix + 'cat_RADeg'], ref[prefix + 'cat_decDeg'])
        pixDeg = mapDicts[0]['wcs'].getPixelSizeDeg()         # the reference's table keeps sky positions, not x / y
        assert np.abs(np.asarray(cat['RADeg'], dtype=float) - ref[prefix + 'cat_RADeg']).max() < 0.01 * pixDeg
        assert np.abs(np.asarray(cat['decDeg'], dtype=float) - ref[prefix + 'cat_decDeg']).max() < 0.01 * pixDeg
        assert np.array_equal(np.asarray(cat['numSigPix'], dtype=float), ref[prefix + 'cat_numSigPix'])
        for k in [k[len(prefix) + 4:] for k in ref.files if k.startswith(prefix + 'cat_')]:
            if k not in ('name', 'template', 'tileName', 'RADeg', 'decDeg', 'numSigPix'):
                np.testing.assert_allclose(np.asarray(cat[k], dtype=np.float64), ref[prefix + 'cat_' + k], rtol=1e-4, atol=1e-9, err_msg=k)
        if prefix + 'fRel' in ref.files:
            gotRel = np.array([obj.fRelWeights[k] for k in sorted(obj.fRelWeights.keys())])
            assert np.abs(gotRel - ref[prefix + 'fRel']).max() < 1e-8
        report("reference_run_big_%s_%d" % (name, i), maxabs_dSN_sub=float(np.abs(rec[prefix + 'SN_sub'] - ref[prefix + 'SN_sub']).max()), rows=len(cat))


def test_batched_driver_matches_full_size_reference_run(report):
    """The same configs[3] tile through the batched driver (pipelines.filterMapsAndMakeCatalogs: K1-K11 over tile x scale
    images, CUDA graphs): its per-(tile, scale) catalogs against the reference run's, fixed_ columns included."""
    import big_cases
    from nemo_b200 import pipelines
    name = 'dr5_f150_f098_840x1440'
    ref = np.load(os.path.join(HERE, 'golden', 'ref_big_' + name + '.npz'))
    case = big_cases.BIG_CASES[name]
    parDict = {'mapFilters': case['filters'], 'photFilter': case['photFilter'], 'thresholdSigma': 4.0, 'minObjPix': 1,
               'findCenterOfMass': True, 'useInterpolator': True, 'rejectBorder': 0}
    config = pipelines.SimpleConfig(parDict, {'T0': case['tile']()})
    opt = pipelines.filterMapsAndMakeCatalogs(config, verbose=False)
    for i, f in enumerate(case['filters']):
        prefix = 's%d_cat_' % i
        cat = config.catalogDict[f['label'] + '#T0']['catalog']
        assert len(cat) == len(ref[prefix + 'RADeg'])
        big_cases.assertNamesMatch(list(cat['name']), list(ref[prefix + 'name']), ref[prefix + 'RADeg'], ref[prefix + 'decDeg'])
        for k in [k[len(prefix):] for k in ref.files if k.startswith(prefix)]:
            if k in ('name', 'template', 'tileName'):
                continue
            tol = dict(rtol=0, atol=0.01 * 0.5 / 60) if k in ('RADeg', 'decDeg') else dict(rtol=1e-4, atol=1e-9)
            np.testing.assert_allclose(np.asarray(cat[k], dtype=np.float64), ref[prefix + k], err_msg=k, **tol)
    assert len(opt) >= 20
    report("batched_reference_run_big", optimal=len(opt))


def test_measure_shapes_matches_reference_run(report):
    """findObjects(measureShapes = True) on the device (nb_label_shapes: bounding-box corners and S/N-weighted moments per label,
    then the reference's formulas on the host) against the reference's own findObjects on the seeded S/N map of shape_case
    (ref_shapes.npz): same rows, the same -99 rows for small objects, ellipse columns within 1e-9 (the moment sums are fp64 atomics,
    so the last digits depend on the order).  The two exactly symmetric blobs, whose xy moment is rounding noise and whose
    position angle therefore a coin toss, are compared on A, B only.  Also through the batched entry point with two images, the
    second one inverted."""
    import shape_case
    from nemo_b200 import photometry, device
    ops64 = device.getOps('f64')
    ref = np.load(os.path.join(HERE, 'golden', 'ref_shapes.npz'))
    tile = shape_case.shapeTile()
    det = dict(shape_case.DETECT)
    cat = photometry.findObjects(tile, verbose=False, **det)
    assert len(cat) == len(ref['shape_SNR'])
    np.testing.assert_allclose(np.asarray(cat['SNR'], dtype=float), ref['shape_SNR'], rtol=1e-6)
    assert np.array_equal(np.asarray(cat['numSigPix'], dtype=float), ref['shape_numSigPix'])
    small = ref['shape_numSigPix'] <= 9
    knife = (~small) & ((ref['shape_ellipse_e'] < 1e-6) | (ref['shape_ellipse_A'] == -99))
    robust = ~small & ~knife
    assert robust.sum() >= 15 and small.sum() >= 2
    worst = 0.0
    for k in shape_case.SHAPE_KEYS:
        got = np.asarray(cat[k], dtype=np.float64)
        assert np.all(got[small] == -99), k
        np.testing.assert_allclose(got[robust], ref['shape_' + k][robust], rtol=1e-9, atol=1e-9, err_msg=k)
        worst = max(worst, float(np.abs(got[robust] - ref['shape_' + k][robust]).max()))
    for k in ('ellipse_A', 'ellipse_B'):
        got = np.asarray(cat[k], dtype=np.float64)[knife]
        want = ref['shape_' + k][knife]
        both = (got != -99) & (want != -99)
        np.testing.assert_allclose(got[both], want[both], rtol=1e-9, err_msg=k)
    # batched: image 0 as above, image 1 the negated map searched with invertMap (the same objects)
    sn = ops64.to_device(np.stack([tile['SNMap'], tile['SNMap']]), ops64.dtype)
    tabs = photometry.findObjectsBatch(ops64, sn[:1], [tile['wcs']], [tile['flagMask']], ['shapes'], ['T0'], threshold=det['threshold'],
                                       minObjPix=det['minObjPix'], rejectBorder=0, removeRings=True, ringThresholdSigma=det['ringThresholdSigma'],
                                       measureShapes=True)
    inv = photometry.findObjectsBatch(ops64, -sn[1:], [tile['wcs']], [tile['flagMask']], ['shapes'], ['T0'], threshold=det['threshold'],
                                      minObjPix=det['minObjPix'], rejectBorder=0, removeRings=True, ringThresholdSigma=det['ringThresholdSigma'],
                                      invertMap=True, measureShapes=True)
    for t in (tabs[0], inv[0]):
        assert len(t) == len(cat)
        for k in shape_case.SHAPE_KEYS:
            np.testing.assert_allclose(np.asarray(t[k], dtype=float)[robust], np.asarray(cat[k], dtype=float)[robust], rtol=1e-9, atol=1e-9, err_msg=k)
    report("measure_shapes_reference_run", rows=len(cat), ellipses=int(robust.sum()), max_abs_err=worst)
