"""
CPU tests of the oracle (oracle/): the committed golden fixtures (tests/golden/*.npz, made by
tests/golden/make_golden.py) and known answers for the restated third-party pieces.  The reference ships no
golden vectors for this path, so these pin the oracle itself: if its arithmetic drifts, they fail.
"""

import os
import sys

import numpy as np
import pytest
from scipy import ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))

import make_golden  # noqa: E402
from oracle import pixell_shim as enmap, filters_ref, photometry_ref, signals_ref, synth  # noqa: E402
from oracle.wcs_stub import SimpleCARWCS, calcAngSepDeg, j2000ToGalactic  # noqa: E402


@pytest.mark.parametrize("name", sorted(make_golden.CASES.keys()))
def test_oracle_reproduces_golden_fixture(name):
    gold = np.load(os.path.join(HERE, 'golden', name + '.npz'))
    mapDicts, out = make_golden.runCase(name)
    assert list(out['input_sha256']) == list(gold['input_sha256']), "the seeded input generator changed"
    assert np.array_equal(out['surveyMask'], gold['surveyMask'])
    assert np.array_equal(out['segmentation'], gold['segmentation'])
    assert int(out['numObjects']) == int(gold['numObjects'])
    # identical machine arithmetic reproduces bit for bit; a different BLAS/FFT build may differ in the last digits
    np.testing.assert_allclose(out['SNMap'], gold['SNMap'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(out['data'], gold['data'], rtol=1e-9, atol=1e-9 * np.abs(gold['data']).max())
    assert abs(float(out['signalNorm']) / float(gold['signalNorm']) - 1) < 1e-10
    for k in gold.files:
        if k.startswith('cat_') and k != 'cat_name':
            np.testing.assert_allclose(out[k], gold[k], rtol=1e-8, atol=1e-10)
    assert list(out['cat_name']) == list(gold['cat_name'])


def test_apod_profile_matches_pixell_semantics():
    """enmap.apod(m, 20): cosine taper over the first / last 20 samples, first sample 0, 20th exactly 1
    (filters.py:528, 749: `apod(ones) == 1` trims 19 pixels per edge)."""
    a = enmap.apod(np.ones((64, 80)), 20)
    assert a[0, 40] == 0.0 and a[40, 0] == 0.0
    assert a[19, 40] == 1.0 and a[40, 19] == 1.0 and a[18, 40] < 1.0
    inner = (a == 1)
    assert inner[19:-19, 19:-19].all() and inner.sum() == (64 - 38) * (80 - 38)
    assert np.allclose(a[5, 40], 0.5 * (1 - np.cos(np.pi * 5 / 19)))
    assert np.allclose(a, a[::-1, ::-1])


def test_fft_conventions():
    """enmap.fft normalises by sqrt(Npix) (Parseval holds); enmap.ifft(normalize=False) is the bare inverse sum."""
    rng = np.random.default_rng(0)
    m = rng.standard_normal((32, 48))
    fm = enmap.fft(m)
    assert np.allclose((np.abs(fm) ** 2).sum(), (m ** 2).sum())
    assert np.allclose(enmap.ifft(fm, normalize=True).real, m)
    assert np.allclose(enmap.ifft(fm, normalize=False).real, m * np.sqrt(32 * 48))


def test_apply_window_known_answer():
    """Order-0 pixel window: a single Fourier mode is scaled by sinc(fy) sinc(fx); pow=-1 undoes pow=+1."""
    H, W = 32, 64
    ky, kx = 3, 5
    yy, xx = np.mgrid[:H, :W]
    mode = np.cos(2 * np.pi * (ky * yy / H + kx * xx / W))
    out = enmap.apply_window(mode, pow=1.0)
    assert np.allclose(out, mode * np.sinc(ky / H) * np.sinc(kx / W))
    rng = np.random.default_rng(1)
    m = rng.standard_normal((H, W))
    assert np.allclose(enmap.apply_window(enmap.apply_window(m, 1.0), -1.0), m)


def test_wcs_stub_round_trip_and_separation():
    wcs = SimpleCARWCS((100, 200), 0.5 / 60, crval=(10.0, -2.0))
    x, y = wcs.wcs2pix(*wcs.pix2wcs(12.25, 80.5))
    assert abs(x - 12.25) < 1e-9 and abs(y - 80.5) < 1e-9
    ra, dec = wcs.getCentreWCSCoords()
    assert abs(ra - 10.0) < 1e-12 and abs(dec + 2.0) < 1e-12
    assert abs(calcAngSepDeg(10.0, 0.0, 10.0, 1.0) - np.degrees(np.tan(np.radians(1.0)))) < 1e-12
    # Galactic centre and north Galactic pole (J2000)
    l, b = j2000ToGalactic(266.40499, -28.93617)
    assert abs(b) < 0.01 and (abs(l) < 0.01 or abs(l - 360) < 0.01)
    l, b = j2000ToGalactic(192.85948, 27.12825)
    assert abs(b - 90) < 0.01


def test_names_follow_the_reference_string_slicing():
    """catalogs.py:259-346: digits are cut from printed values, never rounded."""
    assert photometry_ref.makeName(10.0, -1.0) == "ACT-CL J0040.0-0100"
    assert photometry_ref.makeName(359.9999, 45.9999) == "ACT-CL J2359.9+4559"
    assert photometry_ref.makeName(15.0 * 1.5 + 15.0 * 0.99 / 60, 0.5) == "ACT-CL J0130.9+0030"
    assert photometry_ref.makeName(150.0, -12.51, prefix='MOCK-S') == "MOCK-S J1000.0-1230"


def test_signal_conversions():
    assert abs(signals_ref.fSZ(217.0)) < 0.01                    # SZ null
    assert signals_ref.fSZ(148.0) < 0 < signals_ref.fSZ(280.0)
    assert np.isclose(signals_ref.convertToDeltaT(1e-4, 148.0), 2.7255e6 * 1e-4 * signals_ref.fSZ(148.0), rtol=1e-3)
    # Jy/sr conversion is linear in temperature
    assert np.isclose(signals_ref.deltaTToJyPerSr(2.0, 150.0), 2 * signals_ref.deltaTToJyPerSr(1.0, 150.0))


def test_make_noise_map_last_writer_wins_and_asymmetric_clip():
    """filters.py:409-481: cells extended by half a cell, later cells overwrite earlier ones in the overlaps, the clip
    test is |v| < |mean + 3 sigma| (not |v - mean|)."""
    rng = np.random.default_rng(3)
    H = W = 160
    pixDeg = 0.5 / 60
    data = rng.standard_normal((H, W))
    data[:, W // 2:] *= 4.0
    wcs = SimpleCARWCS((H, W), pixDeg)
    md = [{'data': data, 'weights': np.ones((H, W)), 'wcs': wcs, 'flagMask': np.zeros((H, W), np.uint8), 'obsFreqGHz': 150.0,
           'label': 'x', 'surveyMask': None, 'pointSourceMask': None}]
    f = filters_ref.MapFilter('t', md, {'noiseParams': {'noiseGridArcmin': 20.0}})
    rms = f.makeNoiseMap(data)
    def clipped(cell):
        mean, sig = cell.mean(), cell.std()
        for _ in range(10):
            sel = cell[np.abs(cell) < abs(mean + 3 * sig)]
            mean, sig = sel.mean(), sel.std()
        return sig

    # cells are [40 i - 20, 40 (i + 1) + 20) along each axis, clipped to the map
    # bottom-right pixel belongs only to the last cell
    assert np.isclose(rms[-1, -1], clipped(data[100:, 100:]), rtol=1e-12)
    # pixel (45, 45) lies in cells (0,0), (0,1), (1,0), (1,1): the LAST one in raster order wins
    assert np.isclose(rms[45, 45], clipped(data[20:100, 20:100]), rtol=1e-12)
    assert not np.isclose(rms[45, 45], clipped(data[0:60, 0:60]), rtol=1e-6)
    # pixel (45, 65) also lies in cells (0,2), (1,2): cell (1,2) = rows [20,100), cols [60,140) wins
    assert np.isclose(rms[45, 65], clipped(data[20:100, 60:140]), rtol=1e-12)
    assert len(np.unique(rms)) <= 16


def test_get_object_positions_matches_scipy_numbering():
    rng = np.random.default_rng(5)
    sn = ndimage.gaussian_filter(rng.standard_normal((96, 96)), 1.5) * 12
    ids, pos, npix, seg = photometry_ref.getObjectPositions(sn, 4.0)
    lab, n = ndimage.label(sn > 4.0)
    assert np.array_equal(seg, lab) and ids[0] == 0 and len(ids) == n + 1
    assert npix[0] == 0 and npix[1:].sum() == (sn > 4.0).sum()
    com = ndimage.center_of_mass(sn, lab, np.arange(1, n + 1))
    assert np.allclose(np.array(pos[1:]), np.array(com))


def test_synth_is_seeded():
    a, _ = synth.makeTile(shape=(64, 64), numObjects=2, seed=9, margin=20)
    b, _ = synth.makeTile(shape=(64, 64), numObjects=2, seed=9, margin=20)
    c, _ = synth.makeTile(shape=(64, 64), numObjects=2, seed=10, margin=20)
    assert np.array_equal(a[0]['data'], b[0]['data']) and not np.array_equal(a[0]['data'], c[0]['data'])


def test_smart_noise_branch_known_answers():
    """noiseGridArcmin = 'smart' (filters.py:366-407): numBins - 1 bins between the extreme weights, strict inequalities (the
    extreme-weight pixels and anything exactly on an edge stay unpainted), clipped standard deviation per bin."""
    from oracle import filters_ref
    from oracle.wcs_stub import SimpleCARWCS
    rng = np.random.default_rng(3)
    H, W = 64, 96
    w = np.empty((H, W))
    w[:, :32], w[:, 32:64], w[:, 64:] = 1.0, 2.0, 3.0
    w[0, 0], w[0, 95] = 0.5, 4.0                       # extremes: edges = linspace(0.5, 4, 3) = [0.5, 2.25, 4]
    m = rng.standard_normal((H, W)) * np.where(w <= 2.0, 1.0, 5.0)
    m[5, 5] = 0.0                                      # a zero pixel is not a sample but is painted
    md = [{'data': m, 'weights': w, 'wcs': SimpleCARWCS(m.shape, 0.5 / 60), 'flagMask': np.zeros(m.shape, np.uint8), 'obsFreqGHz': 150.0,
           'label': 'x', 'surveyMask': None, 'pointSourceMask': None}]
    f = filters_ref.MapFilter('t', md, {'noiseParams': {'noiseGridArcmin': 'smart', 'numNoiseBins': 3}})
    rms = f.makeNoiseMap(m)
    assert rms[0, 0] == 0 and rms[0, 95] == 0          # on the first / last edge: in no bin
    lowBin, highBin = (w > 0.5) & (w < 2.25), (w > 2.25) & (w < 4.0)
    assert len(np.unique(rms[lowBin])) == 1 and len(np.unique(rms[highBin])) == 1
    assert rms[5, 5] == rms[6, 6] > 0
    assert abs(rms[6, 6] - 1.0) < 0.1 and abs(rms[6, 80] - 5.0) < 0.5

    def clipped(v):
        mean, std = v.mean(), v.std()
        for _ in range(10):
            k = np.abs(v) < abs(mean + 3 * std)
            mean, std = v[k].mean(), v[k].std()
        return std
    assert np.isclose(rms[6, 6], clipped(m[lowBin & (m != 0)]), rtol=1e-14)
    with pytest.raises(Exception):
        filters_ref.MapFilter('t', md, {'noiseParams': {'noiseGridArcmin': 'smart'}}).makeNoiseMap(m)


@pytest.mark.parametrize("name", sorted(make_golden.CASES.keys()))
def test_oracle_matches_reference_run(name):
    """The fixtures ref_<case>.npz hold what the REFERENCE's own source files computed for the golden tiles
    (tests/golden/make_reference_golden.py: nemo/filters.py and nemo/photometry.py imported unmodified from /root/reference, the
    absent third-party packages replaced by the shims, the template maps given as inputs).  The oracle must reproduce them bit
    for bit: filtered map, S/N map, area mask, normalisation, relative frequency weights, labels, positions and pixel counts,
    the catalog of findObjects + measureFluxes (names, sky positions, S/N, pixel counts, flags, fluxes) and the 'smart' and
    'percentile' branches of makeNoiseMap."""
    from oracle import filters_ref, photometry_ref
    ref = np.load(os.path.join(HERE, 'golden', 'ref_' + name + '.npz'))
    mapDicts, out = make_golden.runCase(name)
    assert np.array_equal(out['data'], ref['data'])
    assert np.array_equal(out['SNMap'], ref['SNMap'])
    assert np.array_equal(out['surveyMask'] != 0, ref['surveyMask'] != 0)
    assert float(out['signalNorm']) == float(ref['signalNorm'])
    if 'fRelWeights' in out:
        assert np.array_equal(out['fRelWeights'][np.argsort(out['fRelWeights'][:, 0]), 1], ref['fRelValues'])
    ids, pos, numPix, seg = photometry_ref.getObjectPositions(out['SNMap'], 4.0)
    assert np.array_equal(seg, ref['segmentationMap']) and np.array_equal(ids, ref['objIDs'])
    assert np.array_equal(np.asarray(pos, dtype=np.float64).reshape(-1, 2), ref['objPositions'], equal_nan=True)
    assert np.array_equal(numPix, ref['objNumPix'])
    assert len(ids) - 1 >= 3
    # the reference's findObjects / measureFluxes / makeName on its own filtered maps: the catalog, column by column
    refKeys = [k[7:] for k in ref.files if k.startswith('refcat_')]
    assert len(ref['refcat_RADeg']) == int(out['numObjects']) >= 3
    for k in refKeys:
        if k in ('template', 'tileName'):
            continue
        if ref['refcat_' + k].dtype.kind in 'US':
            assert list(ref['refcat_' + k]) == list(out['cat_' + k]), k
        else:
            assert np.array_equal(ref['refcat_' + k], out['cat_' + k]), k
    assert {'name', 'RADeg', 'decDeg', 'SNR', 'numSigPix', 'flags', 'galacticLatDeg', 'deltaT_c', 'err_deltaT_c'} <= set(refKeys)
    case = make_golden.CASES[name]
    for key, noiseParams in (('smart', {'method': 'dataMap', 'noiseGridArcmin': 'smart', 'numNoiseBins': 6}),
                             ('percentile', {'method': 'dataMap', 'noiseGridArcmin': 20., 'numNoiseBins': 2 if name.startswith('beam') else 1,
                                             'RMSEstimator': 'percentile'})):   # constant weights leave bins empty, and np.percentile([]) raises
        params = dict(make_golden.copyFilter(case['filter'])['params'], noiseParams=noiseParams)
        f = filters_ref.MapFilter('t', mapDicts, params)
        got = f.makeNoiseMap(out['SNMap'] * 3.0 + 0.0)
        assert np.array_equal(got, ref['noise_' + key]), key
        if name.startswith('beam') or key == 'percentile':      # constant weights have no 'smart' bins: all zero, as the reference
            assert (got > 0).mean() > 0.3


def test_standalone_pieces_match_reference_run():
    """ref_misc.npz: the reference's RealSpaceMatchedFilter.applyFilter (Gaussian high-pass + kernel convolution,
    nemo/filters.py:1172-1218), catalogs.makeOptimalCatalog (nemo/catalogs.py:149-189) and maps.getPixelAreaArcmin2Map
    (nemo/maps.py:1461-1482), run from their own source files on seeded inputs; the oracle and the host-side merge reproduce
    them."""
    import make_reference_golden as mrg
    from oracle import filters_ref, completeness_ref
    from nemo_b200 import catalogs, completeness
    from nemo_b200.wcs import CARWCS
    ref = np.load(os.path.join(HERE, 'golden', 'ref_misc.npz'))
    inp = mrg.miscInputs()
    obj = object.__new__(filters_ref.RealSpaceMatchedFilter)
    obj.params = {'bckSub': True}
    obj.wcs, obj.kern2d, obj.signalNorm = inp['wcs'], inp['kern'], inp['signalNorm']
    obj.applyRACentre, obj.applyDecCentre, obj.bckSubScaleArcmin = inp['applyRADeg'], inp['applyDecDeg'], inp['bckSubScaleArcmin']
    obj.unfilteredMapsDictList = [{'obsFreqGHz': 97.8}, {'obsFreqGHz': 149.6}]
    assert np.array_equal(obj.applyFilter(inp['cube'].copy(), calcFRelWeights=True), ref['rs_filtered'])
    assert np.array_equal(np.array([obj.fRelWeights[97.8], obj.fRelWeights[149.6]]), ref['rs_fRel'])
    obj.params = {'bckSub': False}
    assert np.array_equal(obj.applyFilter(inp['cube'].copy()), ref['rs_filteredNoBck'])
    # pixel areas
    area = completeness_ref.getPixelAreaArcmin2Map(inp['shape'], inp['wcs'])
    assert bool(ref['pixAreaConstantAlongRows']) and np.array_equal(area[:, 0], ref['pixAreaArcmin2Rows'])
    wcs = CARWCS(inp['shape'], 0.5 / 60.0, crval=(30.0, -20.0))
    assert np.allclose(completeness.rowAreasDeg2(inp['shape'], wcs) * 3600.0, ref['pixAreaArcmin2Rows'], rtol=1e-13, atol=0)
    # cross-scale merge: same rows, same order
    table = lambda cols: catalogs.Table.fromColumns([(k, np.array(v)) for k, v in cols.items()])
    merged = catalogs.makeOptimalCatalog({'A#t': {'catalog': table(inp['catA'])}, 'B#t': {'catalog': table(inp['catB'])}})
    assert len(merged) == len(ref['opt_SNR']) and 30 < len(merged) < len(inp['catA']['SNR']) + len(inp['catB']['SNR'])
    for k in ('RADeg', 'decDeg', 'SNR'):
        assert np.array_equal(np.asarray(merged[k], dtype=np.float64), ref['opt_' + k]), k


def test_phot_filter_flow_matches_reference_run():
    """ref_misc.npz phot_*: the reference's findObjects + getSNRValues(prefix='fixed_') + measureFluxes(photFilteredMapDict=...) on a
    second filter scale of the golden Arnaud tile (the flow of nemo/pipelines.py:238-284); the oracle reproduces every column."""
    import copy
    import make_reference_golden as mrg
    from oracle import filters_ref, photometry_ref
    ref = np.load(os.path.join(HERE, 'golden', 'ref_misc.npz'))
    mapDicts, _ = make_golden.runCase('arnaud_f148_128')
    outs = [filters_ref.filterMaps(mapDicts, copy.deepcopy(f), 'T0') for f in (make_golden.CASES['arnaud_f148_128']['filter'], mrg.PHOT_SECOND_SCALE)]
    phot = {'SNMap': outs[0]['SNMap'], 'data': outs[0]['data']}
    cat = photometry_ref.findObjects(outs[1], **make_golden.DETECT)
    photometry_ref.getSNRValues(cat, phot['SNMap'], outs[1]['wcs'], prefix='fixed_', useInterpolator=True)
    photometry_ref.measureFluxes(cat, outs[1], None, photFilteredMapDict=phot, useInterpolator=True)
    assert len(cat) == len(ref['phot_SNR']) >= 3
    for k in [k[5:] for k in ref.files if k.startswith('phot_')]:
        if k in ('template', 'tileName'):
            continue
        col = [row[k] for row in cat]
        if ref['phot_' + k].dtype.kind in 'US':
            assert [str(v) for v in col] == list(ref['phot_' + k]), k
        else:
            assert np.array_equal(np.array(col, dtype=np.float64), ref['phot_' + k]), k


@pytest.mark.parametrize("name", ['arnaud_f148_1024', 'dr5_f150_f098_840x1440'])
def test_oracle_matches_full_size_reference_run(name):
    """BASELINE-size tiles: configs[0] (1024 x 1024, ArnaudModelMatchedFilter) and one 840 x 1440 two-frequency tile of configs[3]
    with the photFilter flow.  ref_big_<case>.npz holds, in compact form (tests/golden/big_cases.py), what the REFERENCE's own
    filters.py / photometry.py computed for them; the oracle must reproduce it bit for bit, catalogs included."""
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    import big_cases
    ref = np.load(os.path.join(HERE, 'golden', 'ref_big_' + name + '.npz'))
    _, out = big_cases.runOracle(name)
    for i, f in enumerate(big_cases.BIG_CASES[name]['filters']):
        prefix = 's%d_' % i
        big_cases.compareCompact(out, ref, prefix, exact=True)
        catKeys = [k for k in ref.files if k.startswith(prefix + 'cat_')]
        assert len(ref[prefix + 'cat_RADeg']) >= 20
        assert prefix + 'cat_y_c' in catKeys and (i == 0 or prefix + 'cat_fixed_SNR' in catKeys or big_cases.BIG_CASES[name]['photFilter'] is None)
        for k in catKeys:
            if k.endswith('cat_template') or k.endswith('cat_tileName'):
                continue
            if ref[k].dtype.kind in 'US':
                assert list(ref[k]) == list(out[k]), k
            else:
                assert np.array_equal(ref[k], out[k]), k
        if prefix + 'fRel' in ref.files:
            assert np.array_equal(ref[prefix + 'fRel'], out[prefix + 'fRel'])


def test_model_image_painter_matches_reference_run():
    """oracle/maps_ref.makeModelImage (point-source branch) against the reference's own maps.makeModelImage run on the same
    catalog (tests/golden/ref_paint.npz): identical arrays, with and without the pixel window, validAreaSection and S/N cut."""
    import make_reference_golden as gen
    from oracle import maps_ref, signals_ref
    inp = gen.paintInputs()
    ref = np.load(os.path.join(HERE, 'golden', 'ref_paint.npz'))
    beam = signals_ref.BeamProfile(profile1d=inp['beamProfile'], rDeg=inp['beamRDeg'])
    assert np.array_equal(maps_ref.makeModelImage(inp['shape'], inp['wcs'], inp['catalog'], beam), ref['paint_full'])
    assert np.array_equal(maps_ref.makeModelImage(inp['shape'], inp['wcs'], inp['catalog'], beam, applyPixelWindow=False), ref['paint_noWindow'])
    assert np.array_equal(maps_ref.makeModelImage(inp['shape'], inp['wcs'], inp['catalog'], beam, validAreaSection=inp['validAreaSection'],
                                                  minSNR=inp['minSNR']), ref['paint_cut'])
    far = dict(inp['catalog'], RADeg=inp['catalog']['RADeg'] + 90.0)
    assert maps_ref.makeModelImage(inp['shape'], inp['wcs'], far, beam) is None


def test_biweight_scale_restatement():
    """oracle.filters_ref.biweight_scale (astropy absent: restated from the published definition, parity unpinned): a sample
    evaluated by hand, location / scale behaviour, robustness to an outlier, the Gaussian limit and the MAD = 0 case."""
    from oracle.filters_ref import biweight_scale
    # [1, 2, 3, 4, 100]: M = 3, MAD = 1, u = d / 9; the outlier has |u| > 1; n' = 4
    # d = -2, -1, 0, 1: f1 = 4 (1 - 4/81)^4 + 2 (1 - 1/81)^4, f2 = (1 - 4/81)(1 - 20/81) + 2 (1 - 1/81)(1 - 5/81) + 1
    a, b = 1 - 4 / 81., 1 - 1 / 81.
    f1 = 4 * a ** 4 + 2 * b ** 4
    f2 = a * (1 - 20 / 81.) + 2 * b * (1 - 5 / 81.) + 1
    assert abs(biweight_scale(np.array([1., 2, 3, 4, 100]), 9.0, True) - np.sqrt(4 * f1) / f2) < 1e-14
    # without modify_sample_size the factor is the sample size
    assert abs(biweight_scale(np.array([1., 2, 3, 4, 100]), 9.0, False) - np.sqrt(5 * f1) / f2) < 1e-14
    rng = np.random.default_rng(12)
    x = rng.standard_normal(100000)
    s0 = biweight_scale(x, 9.0, True)
    assert abs(s0 - 1.0) < 0.02
    assert abs(biweight_scale(3.0 * x + 5.0, 9.0, True) - 3.0 * s0) < 1e-9
    assert abs(biweight_scale(np.r_[x, 1e9], 9.0, True) - s0) < 1e-4
    assert biweight_scale(np.array([2., 2, 2, 2, 7]), 9.0, True) == 0.0


def test_oracle_ellipses_match_reference_run():
    """findObjects(measureShapes = True): the oracle's restatement of the SExtractor-style ellipse block (photometry.py:126-178)
    against what the reference's own findObjects computed for the seeded S/N map of tests/golden/shape_case.py (ref_shapes.npz,
    made by make_reference_golden.py): every ellipse column bit for bit, the -99 rows (objects of at most 9 pixels, and the two
    exactly symmetric blobs whose xy moment is zero or rounding noise) included.  The host formulas the CUDA path uses
    (nemo_b200.photometry.ellipsesFromMoments) get the same numbers from moments summed here in numpy."""
    import shape_case
    from nemo_b200 import photometry
    ref = np.load(os.path.join(HERE, 'golden', 'ref_shapes.npz'))
    tile = shape_case.shapeTile()
    with np.errstate(divide='ignore', invalid='ignore'):
        cat = photometry_ref.findObjects(tile, verbose=False, **shape_case.DETECT)
    assert len(cat) == len(ref['shape_SNR']) >= 20
    assert np.array_equal(np.array([r['numSigPix'] for r in cat], dtype=np.float64), ref['shape_numSigPix'])
    for k in shape_case.SHAPE_KEYS:
        assert np.array_equal(np.array([r[k] for r in cat], dtype=np.float64), ref['shape_' + k]), k
    assert int((ref['shape_ellipse_A'] > 0).sum()) >= 15 and int((ref['shape_ellipse_A'] == -99).sum()) >= 2
    # moments of every kept object about its bounding-box corner -> the host formulas
    ids, pos, numPix, seg = photometry_ref.getObjectPositions(tile['SNMap'], shape_case.DETECT['threshold'])
    rows, nPix = [], []
    for i in range(len(ids)):
        if ids[i] == 0 or numPix[i] <= shape_case.DETECT['minObjPix']:
            continue
        ys, xs = np.where(seg == ids[i])
        d = tile['SNMap'][ys, xs]
        xs, ys = xs - xs.min(), ys - ys.min()
        yMin, xMin = np.where(seg == ids[i])[0].min(), np.where(seg == ids[i])[1].min()
        rows.append([d.sum(), (xs * d).sum(), (ys * d).sum(), (xs ** 2 * d).sum(), (ys ** 2 * d).sum(), (xs * ys * d).sum(), xMin, yMin])
        nPix.append(numPix[i])
    got = photometry.ellipsesFromMoments(np.array(rows), np.array(nPix))
    assert len(rows) == len(cat)          # nothing on a ring, every object above the S/N cut in this map
    robust = np.abs(ref['shape_ellipse_e']) > 1e-6      # away from the symmetric knife edge
    for k in shape_case.SHAPE_KEYS:
        np.testing.assert_allclose(got[k][robust], ref['shape_' + k][robust], rtol=1e-12, atol=1e-12, err_msg=k)
