"""
CPU tests of the host side of the boundary: the C-ABI library loads and exports what include/nemo_b200.h declares,
the product path refuses to run without a CUDA device, and the host-only logic (parameter handling, catalog tables,
names, cross-scale merge) matches the oracle / the reference's documented behaviour.  No kernel is launched here.
"""

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from nemo_b200 import _build, _lib
    _build.build(verbose=False)
    return _lib.load()


def _headerDeclarations():
    text = open(os.path.join(ROOT, 'include', 'nemo_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    decls = {}
    for m in re.finditer(r'\b(?:int|long long|const char\*)\s+(nb_\w+)\s*\(([^;]*?)\)\s*;', text, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ('', 'void') else len([a for a in args.split(',') if a.strip()])
        decls[m.group(1)] = n
    return decls


def test_library_exports_every_declared_symbol(lib):
    from nemo_b200 import _lib
    decls = _headerDeclarations()
    assert len(decls) >= 25
    for name in decls:
        assert hasattr(lib, name), "libnemo_b200.so lacks %s" % name
    # the ctypes prototypes bind exactly the declared entry points, with the declared number of arguments
    assert sorted(decls.keys()) == sorted(_lib.EXPORTED_SYMBOLS)
    for name, nargs in decls.items():
        assert len(_lib._PROTOTYPES[name]) == nargs, "%s: header has %d arguments, binding %d" % (name, nargs, len(_lib._PROTOTYPES[name]))


def test_library_is_built_for_sm_100a():
    import subprocess
    from nemo_b200 import _lib
    out = subprocess.run(['cuobjdump', '--list-elf', _lib.LIBPATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_host_only_entry_points(lib):
    assert lib.nb_version() >= 100
    # 1 = power of two (direct passes), 2 = any other side in 64..4096 (Bluestein passes), 0 = unsupported
    assert lib.nb_fft_size_supported(1024) == 1 and lib.nb_fft_size_supported(1000) == 2 and lib.nb_fft_size_supported(1031) == 2
    assert lib.nb_fft_size_supported(8192) == 0 and lib.nb_fft_size_supported(48) == 0
    assert lib.nb_label4_scratch_ints(2, 64, 64) >= 2 * 64 * 64
    assert lib.nb_launch_count() == 0      # nothing has been launched: no GPU work happens in the CPU suite


def test_argument_errors_are_reported_without_a_gpu(lib):
    rc = lib.nb_rfft2(0, None, None, 1, 64, 64, 0, 1.0, None)
    assert rc < 0 and b'bad arguments' in lib.nb_last_error()
    buf = ctypes.c_void_p(16)
    rc = lib.nb_rfft2(0, buf, buf, 1, 100, 63, 0, 1.0, None)       # a side below 64
    assert rc == -3 and b'not supported' in lib.nb_last_error()
    rc = lib.nb_rfft2(0, buf, buf, 1, 5000, 64, 0, 1.0, None)
    assert rc == -3
    assert lib.nb_fft_size_supported(1033) == 2 and lib.nb_fft_size_supported(1024) == 1 and lib.nb_fft_size_supported(4097) == 0


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    from nemo_b200 import _lib, device
    with pytest.raises(_lib.NemoB200Error):
        device.getOps('f64')
    from nemo_b200 import photometry
    with pytest.raises(_lib.NemoB200Error):
        photometry.getObjectPositions(np.zeros((64, 64)), 4.0)


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, 'nemo_b200')):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), "%s imports the oracle" % f


# -------------------------------------------------------------------------------------------------------------
def test_noise_grid_and_trim_follow_the_reference():
    from nemo_b200 import engine
    pixDeg = 0.5 / 60
    yc, xc, ov, nb, est = engine.noiseGrid((1024, 1024), {'method': 'dataMap', 'noiseGridArcmin': 40.0, 'numNoiseBins': 8}, pixDeg)
    assert ov == 40 and nb == 8 and len(yc) == 13 and yc[0] == 0 and yc[-1] == 1024 and est == 'clip'
    assert list(yc) == list(np.linspace(0, 1024, int(1024 / 80 + 1), dtype=int))
    yc, xc, ov, nb, est = engine.noiseGrid((300, 420), {'method': 'dataMap', 'RMSEstimator': 'percentile'}, pixDeg)
    assert list(yc) == [0, 300] and list(xc) == [0, 420] and ov == 0 and nb == 1 and est == 'percentile'
    assert engine.noiseGrid((64, 64), {'noiseGridArcmin': 'smart', 'numNoiseBins': 12}, pixDeg) == ('smart', 12)
    with pytest.raises(Exception):
        engine.noiseGrid((64, 64), {'noiseGridArcmin': 'smart'}, pixDeg)                 # numNoiseBins is required (filters.py:367-370)
    with pytest.raises(Exception):
        engine.noiseGrid((64, 64), {'noiseGridArcmin': 'smart', 'numNoiseBins': 4, 'RMSEstimator': 'percentile'}, pixDeg)
    assert engine.noiseGrid((64, 64), {'noiseGridArcmin': 20., 'RMSEstimator': 'biweight'}, pixDeg)[4] == 'biweight'
    with pytest.raises(Exception):
        engine.noiseGrid((64, 64), {'noiseGridArcmin': 'smart', 'numNoiseBins': 4, 'RMSEstimator': 'biweight'}, pixDeg)
    # filters.py:728-735: edgeTrimArcmin 0 means three noise cells
    p = {'noiseParams': {'noiseGridArcmin': 40.0}, 'edgeTrimArcmin': 0.0}
    assert engine.trimSizePix(p, pixDeg, realSpace=False) == 240
    assert engine.trimSizePix({'noiseParams': {'noiseGridArcmin': 40.0}, 'edgeTrimArcmin': 10.0}, pixDeg, False) == 20
    assert engine.trimSizePix(p, pixDeg, realSpace=True) == 0        # filters.py:1137-1141


def test_spectral_weights_match_oracle():
    from nemo_b200 import engine, signals
    from oracle import signals_ref
    maps = [{'units': 'uK', 'obsFreqGHz': 97.8}, {'units': 'uK', 'obsFreqGHz': 149.6}]
    w = engine.spectralWeights(maps, {'outputUnits': 'yc'})
    assert np.allclose(w, [signals_ref.fSZ(97.8), signals_ref.fSZ(149.6)], rtol=0, atol=0)
    assert list(engine.spectralWeights(maps, {'outputUnits': 'uK'})) == [1.0, 1.0]
    w = engine.spectralWeights(maps, {'outputUnits': 'uK', 'alpha': -0.7})
    assert np.isclose(w[1], (149.6 / 97.8) ** -0.7)
    with pytest.raises(Exception):
        engine.spectralWeights(maps, {'outputUnits': 'Jy'})
    assert signals.fSZ(148.0) == signals_ref.fSZ(148.0)


def test_templates_match_oracle():
    from nemo_b200 import signals
    from nemo_b200.wcs import CARWCS
    from oracle import signals_ref
    rDeg = np.linspace(0, 0.2, 50)[None, :] * np.ones((4, 1))
    b1, b2 = signals.gaussianBeam(1.4), signals_ref.gaussianBeam(1.4)
    assert np.array_equal(b1.profile1d, b2.profile1d)
    wcs = CARWCS((8, 50))
    assert np.array_equal(signals.makeBeamModelSignalMap(rDeg, wcs, b1), signals_ref.makeBeamModelSignalMap(rDeg, wcs, b2))
    p1 = signals.makeArnaudModelProfile(0.4, 2e14)
    p2 = signals_ref.makeArnaudModelProfile(0.4, 2e14)
    assert np.allclose(p1['tckP'][1], p2['tckP'][1], rtol=1e-12) and np.isclose(p1['theta500Arcmin'], p2['theta500Arcmin'])
    assert 2.0 < p1['theta500Arcmin'] < 3.0


def test_names_match_oracle_and_reference_semantics():
    from nemo_b200 import catalogs
    from oracle import photometry_ref
    rng = np.random.default_rng(2)
    for _ in range(2000):
        ra, dec = rng.uniform(0, 360), rng.uniform(-89, 89)
        if abs(dec) < 1e-3:
            continue
        assert catalogs.makeName(ra, dec, prefix='X') == photometry_ref.makeName(ra, dec, prefix='X')
    assert catalogs.makeName(10.0, -1.0) == "ACT-CL J0040.0-0100"
    assert catalogs.makeLongName(10.0, -1.0) == "ACT-CL J004000.0-010000.0"
    assert list(catalogs.makeNames([10.0, 150.0], [-1.0, -12.51], prefix='S')) == ["S J0040.0-0100", "S J1000.0-1230"]


def _table(n, off, snr0):
    from nemo_b200 import photometry
    cols = {'id': np.arange(n) + 1, 'x': np.arange(n) * 1.0, 'y': np.arange(n) * 2.0, 'RADeg': 10 + off + np.arange(n) * 0.1,
            'decDeg': -1 + np.arange(n) * 0.05, 'galacticLatDeg': np.zeros(n), 'numSigPix': np.ones(n), 'SNR': snr0 + np.arange(n) * 1.0,
            'flags': np.zeros(n, int), 'template': 'A%d' % off, 'tileName': 't', 'objIdent': 'X', 'longNames': False}
    return photometry.columnsToTable(cols, {'y_c': np.arange(n) * 1.0})


def test_table_operations_and_deferred_names(monkeypatch):
    from nemo_b200 import catalogs
    calls = []
    real = catalogs.makeNames
    monkeypatch.setattr(catalogs, 'makeNames', lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    a = _table(5, 0, 5.0)
    assert len(a) == 5 and 'name' in a.keys() and a.keys()[0] == 'name' and not calls
    sub = a[np.array([True, False, True, True, False])]
    sub.sort(['SNR'])
    both = catalogs.vstack([sub, _table(2, 3, 9.0), []])
    assert len(both) == 5 and not calls                      # names still not formatted
    assert list(both['name'][:2]) == ['X J0040.0-0100', 'X J0040.8-0054'] and len(calls) == 2
    row = a[1]
    assert row['template'] == 'A0' and row['id'] == 2
    row['SNR'] = 99.0
    assert a['SNR'][1] == 99.0
    assert a.toRows()[0]['name'] == 'X J0040.0-0100'
    with pytest.raises(ValueError):
        a.add_column(catalogs.Column(np.zeros(3), 'bad'))


def test_make_optimal_catalog_keeps_highest_snr_per_position():
    from nemo_b200 import catalogs
    a = _table(4, 0, 5.0)             # RA 10.0, 10.1, 10.2, 10.3
    b = _table(4, 0, 4.0)             # same positions, lower S/N, different template
    b['template'] = np.full(4, 'B', dtype=object)
    b['SNR'] = np.array([4.0, 20.0, 4.5, 3.0])
    opt = catalogs.makeOptimalCatalog({'A#t': {'catalog': a}, 'B#t': {'catalog': b}, 'C#t': {'catalog': []}})
    assert len(opt) == 4
    assert list(opt['RADeg']) == sorted(opt['RADeg'])
    assert list(opt['template']) == ['A0', 'B', 'A0', 'A0'] and opt['SNR'][1] == 20.0
    empty = catalogs.makeOptimalCatalog({'C#t': {'catalog': []}})
    assert len(empty) == 0
    # brute-force restatement of catalogs.py:171-183 on a denser set
    rng = np.random.default_rng(4)
    n = 300
    ra, dec, snr = rng.uniform(10, 10.5, n), rng.uniform(-0.25, 0.25, n), rng.uniform(4, 20, n)
    from nemo_b200 import wcs as nwcs
    keep = np.ones(n, bool)
    for i in range(n):
        r = nwcs.calcAngSepDeg(ra[i], dec[i], ra, dec)
        idx = np.where(r < catalogs.XMATCH_RADIUS_DEG)[0]
        if len(idx) > 1:
            best = idx[np.argmax(snr[idx])]
            keep[idx[idx != best]] = False
    assert np.array_equal(~catalogs._crossMatchKeepBest(ra, dec, snr, catalogs.XMATCH_RADIUS_DEG), keep)


def test_catalog_keys_and_filter_order():
    from nemo_b200 import pipelines
    keys = pipelines.catalogKeys(True, 'yc', True)
    assert keys[0] == 'name' and 'fixed_SNR' in keys and 'fixed_y_c' in keys and not any(k.startswith('ellipse') for k in keys)
    assert 'fluxJy' in pipelines.catalogKeys(False, 'uK', True) and 'fixed_SNR' not in pipelines.catalogKeys(False, 'uK', True)
    par = {'photFilter': 'b', 'mapFilters': [{'label': 'a'}, {'label': 'b'}, {'label': 'c'}]}
    assert [f['label'] for f in pipelines._orderedFilters(par)] == ['b', 'a', 'c']      # pipelines.py:153-167


def test_wcs_matches_oracle_stub():
    from nemo_b200.wcs import CARWCS, calcAngSepDeg
    from oracle import wcs_stub
    a, b = CARWCS((100, 200), 0.5 / 60, crval=(350.0, 3.0)), wcs_stub.SimpleCARWCS((100, 200), 0.5 / 60, crval=(350.0, 3.0))
    x, y = np.array([0.0, 10.5, 199.0]), np.array([0.0, 50.25, 99.0])
    assert np.array_equal(a.pix2wcs(x, y)[0], b.pix2wcs(x, y)[0]) and np.array_equal(a.pix2wcs(x, y)[1], b.pix2wcs(x, y)[1])
    ra, dec = a.pix2wcs(x, y)
    assert np.allclose(a.wcs2pix(ra, dec)[0], x) and np.array_equal(a.wcs2pix(ra, dec)[0], b.wcs2pix(ra, dec)[0])
    assert calcAngSepDeg(1.0, 2.0, 1.5, 2.5) == wcs_stub.calcAngSepDeg(1.0, 2.0, 1.5, 2.5)


def test_bench_byte_model_matches_survey_table():
    """B(F, S) = 8F/S + 18F + 85 (+16 with ring removal) bytes per pixel and scale at 4-byte reals."""
    import bench
    for F, S in [(1, 1), (2, 1), (2, 16)]:
        assert np.isclose(bench.endToEndBytesPerPixel(F, S, 4, rings=False), 8 * F / S + 18 * F + 85)
        assert np.isclose(bench.endToEndBytesPerPixel(F, S, 4, rings=True), 8 * F / S + 18 * F + 85 + 16)
    assert bench.endToEndBytesPerPixel(1, 1, 8) > bench.endToEndBytesPerPixel(1, 1, 4)


def test_batch_bounds_cover_all_tiles_in_order():
    """Batch sizes of the streaming driver: a single size, or a list mirrored at the end (small first and last batches)."""
    from nemo_b200.pipelines import _batchBounds, _batchKey
    cases = {(16, 8): [8, 8], (16, (2, 4)): [2, 4, 4, 4, 2], (5, (2, 4)): [2, 3], (1, (2, 4)): [1], (10, 4): [4, 4, 2],
             (7, (4,)): [4, 3], (3, (2, 4)): [1, 2]}
    for (n, tpb), sizes in cases.items():
        bounds = _batchBounds(n, list(tpb) if isinstance(tpb, tuple) else tpb)
        assert [j - i for i, j in bounds] == sizes
        assert bounds[0][0] == 0 and bounds[-1][1] == n
        assert all(bounds[k][1] == bounds[k + 1][0] for k in range(len(bounds) - 1))
    # geometry groups in the middle of a run: no small batches at their start / no mirrored tail at their end
    assert [j - i for i, j in _batchBounds(28, [2, 4, 8], rampDown=False)] == [2, 4, 8, 8, 6]
    assert [j - i for i, j in _batchBounds(7, [2, 4, 8], rampUp=False)] == [5, 2]
    assert [j - i for i, j in _batchBounds(20, [2, 4, 8], rampUp=False, rampDown=False)] == [8, 8, 4]
    assert _batchKey([2, 4]) == _batchKey((2, 4)) == (2, 4) and _batchKey(8) == 8
    with pytest.raises(Exception):
        _batchBounds(4, [0, 2])


def test_cross_match_equals_the_quadratic_rule():
    """The sweep-based cross-match removes exactly the rows the reference's O(n^2) loop removes (catalogs.py:171-183),
    including across RA = 0 and at high declination."""
    from nemo_b200 import catalogs, wcs

    def brute(ra, dec, snr, r):
        rem = np.zeros(len(ra), bool)
        for i in range(len(ra)):
            m = np.flatnonzero(wcs.calcAngSepDeg(ra[i], dec[i], ra, dec) < r)
            if len(m) > 1:
                best = m[np.argmax(snr[m])]
                rem[m[m != best]] = True
        return rem

    for seed, dec0, ra0 in [(0, 0.0, 0.0), (1, 60.0, 10.0), (2, -30.0, 359.0)]:
        rng = np.random.default_rng(seed)
        n = 500
        ra = (rng.uniform(0, 1.5, n) + ra0) % 360.0
        dec = rng.uniform(-0.8, 0.8, n) + dec0
        snr = rng.uniform(4, 50, n)
        got = catalogs._crossMatchKeepBest(ra, dec, snr, catalogs.XMATCH_RADIUS_DEG)
        want = brute(ra, dec, snr, catalogs.XMATCH_RADIUS_DEG)
        assert want.sum() > 20
        assert np.array_equal(got, want)
        # both candidate generators (declination sweep for small catalogs, sky grid hash for crowded ones) hold every close pair
        sep = lambda p: wcs.calcAngSepDeg(ra[p[:, 0]], dec[p[:, 0]], ra[p[:, 1]], dec[p[:, 1]])
        close = set()
        for i in range(n):
            for j in np.flatnonzero(wcs.calcAngSepDeg(ra[i], dec[i], ra, dec) < catalogs.XMATCH_RADIUS_DEG):
                if j != i:
                    close.add((min(i, int(j)), max(i, int(j))))
        for gen in (catalogs._candidatePairs, catalogs._candidatePairsGrid):
            pairs = gen(ra, dec, catalogs.XMATCH_RADIUS_DEG * 1.01)
            have = set((min(int(a), int(b)), max(int(a), int(b))) for a, b in pairs)
            assert close <= have and len(have) == len(pairs)


def test_row_areas_match_the_reference_pixel_area_map():
    """completeness.rowAreasDeg2 (vectorised) against the oracle's restatement of maps.getPixelAreaArcmin2Map, and the
    oracle's getRMSTab on a hand-made map."""
    from nemo_b200 import completeness
    from nemo_b200.wcs import CARWCS
    from oracle import completeness_ref
    from oracle.wcs_stub import SimpleCARWCS
    shape = (48, 80)
    for crval in [(10.0, 0.0), (200.0, -52.0)]:
        ref = completeness_ref.getPixelAreaArcmin2Map(shape, SimpleCARWCS(shape, 0.5 / 60, crval=crval))
        got = completeness.rowAreasDeg2(shape, CARWCS(shape, 0.5 / 60, crval=crval))
        assert ref.shape == shape and np.array_equal(ref[:, 0], ref[:, -1])
        assert np.allclose(got * 3600.0, ref[:, 0], rtol=1e-13, atol=0)
        assert np.allclose(completeness.getPixelAreaArcmin2Map(shape, CARWCS(shape, 0.5 / 60, crval=crval)), ref, rtol=1e-13, atol=0)
    rms = np.zeros(shape)
    rms[:, :40], rms[:, 40:] = 2.0, 1.0
    rms[:4] = 0.0
    area = (rms > 0).astype(np.uint8)
    values, areas = completeness_ref.getRMSTab(rms, area, SimpleCARWCS(shape, 0.5 / 60))
    assert list(values) == [1.0, 2.0] and abs(areas[0] - areas[1]) < 1e-15
    assert np.isclose(areas.sum(), 44 * 80 * (0.5 / 60) ** 2, rtol=1e-4)
    area[:4] = 1                                         # masked-in area without a noise level: the reference refuses (tol 0.003 deg^2)
    rmsBig, areaBig = np.zeros((600, 600)), np.ones((600, 600), np.uint8)
    rmsBig[:300] = 1.0
    with pytest.raises(Exception):
        completeness_ref.getRMSTab(rmsBig, areaBig, SimpleCARWCS((600, 600), 0.5 / 60))


def test_rows_of_is_a_lazy_independent_copy_of_a_row_range():
    """catalogs.RowsOf: the per-image catalogs of a tile batch are row ranges of the batch's one stacked table, cut out (copied) the
    first time they are looked at; deferred columns (object names) stay deferred; stacking such views gives the parent's rows."""
    from nemo_b200 import catalogs
    calls = []

    def names():
        calls.append(1)
        return np.array(['a', 'b', 'c', 'd', 'e'], dtype=object)

    parent = catalogs.Table.fromColumns([('SNR', np.array([5., 6., 7., 8., 9.])), ('x', np.arange(5.)), ('name', catalogs.deferred(names))])
    parts = [catalogs.RowsOf(parent, 0, 2), catalogs.RowsOf(parent, 2, 5)]
    assert [len(p) for p in parts] == [2, 3] and not calls          # lengths without cutting anything out
    assert list(parts[1]['SNR']) == [7., 8., 9.] and not calls       # numeric columns do not touch the deferred one
    assert list(parts[1]['name']) == ['c', 'd', 'e'] and len(calls) == 1
    parts[0]['SNR'][0] = -1.0                                        # a copy: the parent keeps its values
    assert parent['SNR'][0] == 5.0
    both = catalogs.vstack(parts)
    assert list(both['x']) == [0., 1., 2., 3., 4.] and list(both['name']) == ['a', 'b', 'c', 'd', 'e']
    sub = parts[1][np.array([True, False, True])]
    assert list(sub['x']) == [2., 4.] and parts[1].keys() == ['SNR', 'x', 'name']
