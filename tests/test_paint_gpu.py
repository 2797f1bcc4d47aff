"""
The model-image painter (reference nemo/maps.py:1734-1899 makeModelImage) on the device.  Point-source catalogs against
a run of the reference's own maps.makeModelImage (tests/golden/ref_paint.npz, make_reference_golden.py::runPaint) -- the
unwindowed image to one ulp in a handful of pixels, the windowed one to FFT rounding; cluster catalogs against the oracle's restated painter.
"""

import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))


@pytest.fixture(scope="module")
def ops64():
    from nemo_b200.ops import Ops
    return Ops('f64')


def _inputs():
    import make_reference_golden as gen
    from nemo_b200 import signals
    from nemo_b200.wcs import CARWCS
    inp = gen.paintInputs()
    wcs = CARWCS(inp['shape'], inp['wcs'].pixDeg, crval=inp['wcs'].crval)
    beam = signals.BeamProfile(profile1d=inp['beamProfile'], rDeg=inp['beamRDeg'])
    return inp, wcs, beam


def test_point_source_model_image_matches_reference_run(report):
    from nemo_b200 import maps
    inp, wcs, beam = _inputs()
    ref = np.load(os.path.join(HERE, 'golden', 'ref_paint.npz'))
    got = maps.makeModelImage(inp['shape'], wcs, inp['catalog'], beam, applyPixelWindow=False)
    # same operations in the same order.  One of them is not reproducible bit for bit: the reference squares the row offsets as
    # numpy scalars (yDeg[i]**2, maps.py:2468), which calls libm's pow -- within an ulp of, but not always equal to, the
    # correctly rounded product the device (and numpy on arrays) computes.  On this catalog 10 of 126 000 pixels differ, by one ulp.
    differing = int((got != ref['paint_noWindow']).sum())
    assert np.array_equal(got == 0, ref['paint_noWindow'] == 0) and differing < 1e-3 * got.size
    assert np.abs(got - ref['paint_noWindow']).max() <= 4e-16 * np.abs(ref['paint_noWindow']).max()
    full = maps.makeModelImage(inp['shape'], wcs, inp['catalog'], beam, applyPixelWindow=True)
    cut = maps.makeModelImage(inp['shape'], wcs, inp['catalog'], beam, applyPixelWindow=True, validAreaSection=inp['validAreaSection'],
                              minSNR=inp['minSNR'])
    scale = np.abs(ref['paint_full']).max()
    errFull, errCut = np.abs(full - ref['paint_full']).max() / scale, np.abs(cut - ref['paint_cut']).max() / scale
    report("paint_sources_reference_run", pixels_differing_without_window=differing, relerr_windowed=errFull, relerr_cut=errCut)
    assert errFull < 1e-13 and errCut < 1e-13
    far = dict(inp['catalog'], RADeg=inp['catalog']['RADeg'] + 90.0)
    assert maps.makeModelImage(inp['shape'], wcs, far, beam) is None          # nothing inside the map (maps.py:1812-1813)


def test_batch_of_model_images_and_injection(ops64, report):
    """Several images in one launch, each with its own catalog, beam and WCS, added to the maps they are injected into."""
    from nemo_b200 import maps, signals
    from nemo_b200.wcs import CARWCS
    from oracle import maps_ref, signals_ref
    rng = np.random.default_rng(8)
    shape = (200, 260)
    wcsList = [CARWCS(shape, 0.5 / 60, crval=(20.0 + 3 * i, -50.0 + 9 * i)) for i in range(3)]
    beams = [signals.gaussianBeam(f) for f in (1.4, 2.2, 1.4)]
    cats, refs = [], []
    base = rng.standard_normal((3,) + shape)
    for i in range(3):
        n = [25, 0, 70][i]
        ra, dec = wcsList[i].pix2wcs(rng.uniform(0, shape[1], n), rng.uniform(0, shape[0], n))
        cats.append({'RADeg': np.asarray(ra), 'decDeg': np.asarray(dec), 'deltaT_c': rng.uniform(-300, 800, n)})
        ob = signals_ref.BeamProfile(profile1d=beams[i].profile1d, rDeg=beams[i].rDeg)
        m = maps_ref.makeModelImage(shape, wcsList[i], cats[i], ob) if n > 0 else None
        refs.append(base[i] + (m if m is not None else 0.0))
    got, counts = maps.paintModelImages(ops64, shape, wcsList, cats, beams, base=ops64.to_device(base))
    got = got.cpu().numpy()
    err = max(np.abs(got[i] - refs[i]).max() for i in range(3)) / 800.0
    report("paint_batch_injection", relerr=err, counts=counts)
    assert counts[1] == 0 and np.allclose(got[1], base[1], rtol=0, atol=1e-12) and err < 1e-13


def test_cluster_model_image_matches_oracle_painter(report):
    """y_c catalogs: override = one (M500, z) for all objects, as the source-injection test paints them (maps.py:1834-1849), and
    per-object template names (:1866-1872), converted to delta T at 150 GHz."""
    from nemo_b200 import maps, signals
    from nemo_b200.wcs import CARWCS
    from oracle import maps_ref, signals_ref
    rng = np.random.default_rng(9)
    shape = (256, 300)
    wcs = CARWCS(shape, 0.5 / 60, crval=(40.0, -30.0))
    beam = signals.gaussianBeam(1.4)
    obeam = signals_ref.BeamProfile(profile1d=beam.profile1d, rDeg=beam.rDeg)
    n = 12
    ra, dec = wcs.pix2wcs(rng.uniform(5, shape[1] - 5, n), rng.uniform(5, shape[0] - 5, n))
    cat = {'RADeg': np.asarray(ra), 'decDeg': np.asarray(dec), 'y_c': rng.uniform(0.3, 3.0, n),
           'template': np.array(['Arnaud_M2e14_z0p4', 'Arnaud_M4e14_z0p2', 'Arnaud_M1e14_z0p8'] * 4)}
    for kw in ({'override': {'M500': 2e14, 'redshift': 0.4}}, {}):
        ref = maps_ref.makeModelImage(shape, wcs, cat, obeam, obsFreqGHz=149.6, **kw)
        got = maps.makeModelImage(shape, wcs, cat, beam, obsFreqGHz=149.6, **kw)
        err = np.abs(got - ref).max() / np.abs(ref).max()
        report("paint_clusters_oracle", override=int(bool(kw)), relerr=err)
        assert err < 1e-11


def test_device_injection_equals_host_injection(report):
    """The source-injection loop (maps.py:2093) both ways: models painted by the oracle into the host maps, re-staged and
    re-filtered with the cached filters and RMS maps -- and the same catalogs handed to the driver, which paints them into the
    resident maps on the device (TileBatch.inject).  Same recovered catalogs; a second realisation starts from the original maps
    again, and an empty dict restores them."""
    from oracle import synth, maps_ref, signals_ref
    from nemo_b200 import pipelines
    tileMaps = {}
    for t in range(3):
        md, _ = synth.makeTile(shape=(256, 320), numObjects=3, kind='sources', seed=500 + t, freqsGHz=(148.0,), noiseUK=(80.0,))
        tileMaps['tile%d' % t] = md
    f = {'label': 'Beam', 'class': 'BeamMatchedFilter',
         'params': {'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': 20.}, 'outputUnits': 'uK', 'edgeTrimArcmin': 0.}}
    parDict = {'mapFilters': [f], 'photFilter': 'Beam', 'thresholdSigma': 5.0, 'minObjPix': 1, 'rejectBorder': 0}
    config = pipelines.SimpleConfig(parDict, tileMaps)
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=2, cacheFilters=True)
    original = {k: md[0]['data'].copy() for k, md in tileMaps.items()}
    rng = np.random.default_rng(77)
    keys = ['RADeg', 'decDeg', 'SNR', 'deltaT_c', 'x', 'y']

    def realisation():
        cats = {}
        for k, md in tileMaps.items():
            x, y = rng.uniform(40, 280, 12), rng.uniform(40, 216, 12)
            ra, dec = md[0]['wcs'].pix2wcs(x, y)
            cats[k] = {'RADeg': np.asarray(ra), 'decDeg': np.asarray(dec), 'deltaT_c': rng.uniform(900, 2500, 12)}
        return cats
    for it in range(2):
        cats = realisation()
        got = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=2, useCachedFilters=True, useCachedRMSMap=True,
                                                  injectCatalogs=cats)
        for k, md in tileMaps.items():
            beam = md[0]['beamFileName']
            ob = signals_ref.BeamProfile(profile1d=beam.profile1d, rDeg=beam.rDeg)
            md[0]['data'] = original[k] + maps_ref.makeModelImage(original[k].shape, md[0]['wcs'], cats[k], ob)
        want = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=2, useCachedFilters=True, useCachedRMSMap=True,
                                                   restage=True)
        for b in config._staged['batches']:       # the re-staged upload replaced the resident maps: drop them for the next round
            b.dev, b.devOriginal = None, None
        assert len(got) == len(want) >= 30
        for key in keys:
            np.testing.assert_allclose(np.asarray(got[key], dtype=float), np.asarray(want[key], dtype=float), rtol=1e-9, atol=1e-9, err_msg=key)
        report("device_injection", realisation=it, rows=len(got))
    for k, md in tileMaps.items():
        md[0]['data'] = original[k]
    plain = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=2, useCachedFilters=True, useCachedRMSMap=True, restage=True)
    for b in config._staged['batches']:
        b.dev, b.devOriginal = None, None
    restored = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=2, useCachedFilters=True, useCachedRMSMap=True,
                                                   injectCatalogs={})
    assert len(restored) == len(plain)
    np.testing.assert_allclose(np.asarray(restored['SNR'], dtype=float), np.asarray(plain['SNR'], dtype=float), rtol=1e-12)
