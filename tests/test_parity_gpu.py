"""
End-to-end parity of the drop-in boundary (nemo_b200.filters / photometry / pipelines) against the oracle on
identical seeded inputs.  Tolerances are those of BASELINE.json's north_star: detection catalog identical above the
S/N cut with positions within 0.01 px, S/N maps within 1e-4 (relative, with an absolute floor of 1e-4 because S/N
crosses zero), segmentation labels bit-exact.  fp64 typically agrees to ~1e-10.
"""

import os
import numpy as np
import pytest

from oracle import synth, filters_ref, photometry_ref

pytestmark = pytest.mark.gpu

DETECT = dict(threshold=4.0, minObjPix=1, findCenterOfMass=True, removeRings=True, ringThresholdSigma=3,
              useInterpolator=True, rejectBorder=0)


def _arnaud(label='Arnaud_M2e14_z0p4', M=2e14, z=0.4, grid=40., **extra):
    p = {'M500MSun': M, 'z': z, 'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': grid}, 'outputUnits': 'yc',
         'edgeTrimArcmin': 0.}
    p.update(extra)
    return {'label': label, 'class': 'ArnaudModelMatchedFilter', 'params': p}


def _beam(label='Beam', grid=20., numBins=None, **extra):
    p = {'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': grid}, 'outputUnits': 'uK', 'edgeTrimArcmin': 0.}
    if numBins is not None:
        p['noiseParams']['numNoiseBins'] = numBins
    p.update(extra)
    return {'label': label, 'class': 'BeamMatchedFilter', 'params': p}


def _copyFilter(f):
    import copy
    return copy.deepcopy(f)


def _compareMaps(got, ref, report, name):
    sn, snRef = np.asarray(got['SNMap']), ref['SNMap']
    d, dRef = np.asarray(got['data']), ref['data']
    assert np.array_equal(np.asarray(got['surveyMask']) != 0, ref['surveyMask'] != 0), "survey masks differ"
    errSN = float(np.abs(sn - snRef).max())
    errData = float(np.abs(d - dRef).max() / np.abs(dRef).max())
    report(name, maxabs_dSN=errSN, rel_dData=errData, zeros_equal=int(np.array_equal(d == 0, dRef == 0)))
    assert np.allclose(sn, snRef, rtol=1e-4, atol=1e-4), "S/N maps differ: %g" % errSN
    assert errData < 1e-4
    assert np.array_equal(d == 0, dRef == 0)
    assert got['mapUnits'] == ref['mapUnits'] and got['obsFreqGHz'] == ref['obsFreqGHz']
    assert got['beamSolidAngle_nsr'] == ref['beamSolidAngle_nsr']
    assert np.array_equal(got['flagMask'], ref['flagMask'])


def _compareCatalogs(gotTab, refRows, keys):
    n = len(gotTab) if not isinstance(gotTab, list) else 0
    assert n == len(refRows), "catalog sizes differ: %d vs %d" % (n, len(refRows))
    if n == 0:
        return
    gotRows = gotTab.toRows()
    for a, b in zip(gotRows, refRows):
        assert abs(a['x'] - b['x']) < 0.01 and abs(a['y'] - b['y']) < 0.01
        assert a['numSigPix'] == b['numSigPix'] and a['id'] == b['id'] and a['flags'] == b['flags']
        assert a['template'] == b['template'] and a['tileName'] == b['tileName']
        # names truncate the coordinates, so an object sitting exactly on an arcminute boundary may differ by one
        # digit for a 1e-13 px difference: accept the name of either side's coordinates
        assert a['name'] in (b['name'], photometry_ref.makeName(a['RADeg'], a['decDeg'], prefix=b['name'].split(' J')[0]))
        for k in keys:
            assert abs(a[k] - b[k]) <= 1e-4 * abs(b[k]) + 1e-12, "%s: %g vs %g" % (k, a[k], b[k])


@pytest.mark.parametrize("shape,seed", [((256, 256), 3), ((512, 512), 1)])
def test_arnaud_matched_filter_single_frequency(report, shape, seed):
    """BASELINE config 1 (reduced tile): 148 GHz white noise + clusters, Arnaud theta500 ~ 2.4', S/N > 4."""
    from nemo_b200 import filters, photometry
    mapDicts, truth = synth.makeTile(shape=shape, numObjects=10, kind='clusters', seed=seed)
    f = _arnaud()
    ref, refObj = filters_ref.filterMaps(mapDicts, _copyFilter(f), 'PRIMARY', returnFilter=True)
    got, gotObj = filters.filterMaps(mapDicts, _copyFilter(f), 'PRIMARY', verbose=False, returnFilter=True)
    _compareMaps(got, ref, report, "arnaud_F1_%dx%d" % shape)
    assert abs(gotObj.signalNorm / refObj.signalNorm - 1) < 1e-9
    assert np.abs(gotObj.filt - refObj.filt).max() <= 1e-9 * np.abs(refObj.filt).max()
    for k in refObj.fRelWeights:
        assert abs(gotObj.fRelWeights[k] - refObj.fRelWeights[k]) < 1e-9
    refCat = photometry_ref.findObjects(ref, **DETECT)
    photometry_ref.measureFluxes(refCat, ref, None)
    gotCat = photometry.findObjects(got, **DETECT)
    photometry.measureFluxes(gotCat, got, None)
    assert len(refCat) >= 3
    _compareCatalogs(gotCat, refCat, ['SNR', 'y_c', 'err_y_c', 'deltaT_c', 'err_deltaT_c', 'RADeg', 'decDeg', 'galacticLatDeg'])
    # segmentation labels bit-exact on the GPU-made S/N map
    ids, pos, npix, seg = photometry.getObjectPositions(got['SNMap'], 4.0)
    rids, rpos, rnpix, rseg = photometry_ref.getObjectPositions(np.asarray(got['SNMap']), 4.0)
    assert np.array_equal(seg, rseg) and np.array_equal(ids, rids) and np.array_equal(npix, rnpix)
    assert np.abs(np.array(pos[1:]) - np.array(rpos[1:])).max() < 1e-9


def test_beam_matched_filter_point_sources_with_masks(report):
    """BASELINE config 2 (one reduced tile): point sources, uK output, numNoiseBins, zero border, point-source holes."""
    from nemo_b200 import filters, photometry
    mapDicts, truth = synth.makeTile(shape=(512, 512), freqsGHz=(148.0,), noiseUK=(100.0,), numObjects=40, kind='sources',
                                     seed=1234, borderPix=30, psHoles=3)
    rng = np.random.default_rng(5)
    mapDicts[0]['weights'] = mapDicts[0]['weights'] * (0.6 + 0.8 * rng.random(mapDicts[0]['weights'].shape))
    mapDicts[0]['flagMask'][200:260, 100:180] = 1
    f = _beam(numBins=4, edgeTrimArcmin=10.0)
    ref = filters_ref.filterMaps(mapDicts, _copyFilter(f), 'PRIMARY')
    got = filters.filterMaps(mapDicts, _copyFilter(f), 'PRIMARY', verbose=False)
    _compareMaps(got, ref, report, "beam_F1_bins4")
    refCat = photometry_ref.findObjects(ref, **DETECT)
    photometry_ref.measureFluxes(refCat, ref, None)
    gotCat = photometry.findObjects(got, **DETECT)
    photometry.measureFluxes(gotCat, got, None)
    assert len(refCat) >= 10
    _compareCatalogs(gotCat, refCat, ['SNR', 'deltaT_c', 'err_deltaT_c', 'fluxJy', 'err_fluxJy'])


def test_two_frequency_matched_filter_with_cmb(report):
    """BASELINE config 4 (one reduced tile): 90 + 150 GHz, CMB + noise, two Arnaud scales, photFilter columns."""
    from nemo_b200 import filters, photometry
    mapDicts, truth = synth.makeTile(shape=(512, 512), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4),
                                     numObjects=12, kind='clusters', cmb=True, seed=11)
    fPhot = _arnaud('Arnaud_M2e14_z0p4', 2e14, 0.4, grid=40.)
    fBig = _arnaud('Arnaud_M8e14_z0p2', 8e14, 0.2, grid=40.)
    refPhot, refObj = filters_ref.filterMaps(mapDicts, _copyFilter(fPhot), 'T0', returnFilter=True)
    gotPhot, gotObj = filters.filterMaps(mapDicts, _copyFilter(fPhot), 'T0', verbose=False, returnFilter=True)
    _compareMaps(gotPhot, refPhot, report, "arnaud_F2_phot")
    for k in refObj.fRelWeights:
        assert abs(gotObj.fRelWeights[k] - refObj.fRelWeights[k]) < 1e-8
    assert np.abs(gotObj.filt - refObj.filt).max() <= 1e-8 * np.abs(refObj.filt).max()
    ref = filters_ref.filterMaps(mapDicts, _copyFilter(fBig), 'T0')
    got = filters.filterMaps(mapDicts, _copyFilter(fBig), 'T0', verbose=False)
    _compareMaps(got, ref, report, "arnaud_F2_big")
    refCat = photometry_ref.findObjects(ref, **DETECT)
    photometry_ref.measureFluxes(refCat, ref, None, photFilteredMapDict={'SNMap': refPhot['SNMap'], 'data': refPhot['data']})
    gotCat = photometry.findObjects(got, **DETECT)
    photometry.measureFluxes(gotCat, got, None, photFilteredMapDict={'SNMap': gotPhot['SNMap'], 'data': gotPhot['data']})
    assert len(refCat) >= 3
    _compareCatalogs(gotCat, refCat, ['SNR', 'y_c', 'fixed_SNR', 'fixed_y_c', 'fixed_err_y_c', 'fixed_deltaT_c'])


def test_cached_filter_round_trip(tmp_path, report):
    """saveFilter then useCachedFilter (the source-injection path, maps.py:2093): same maps as a fresh build."""
    from nemo_b200 import filters
    mapDicts, truth = synth.makeTile(shape=(256, 256), numObjects=5, kind='clusters', seed=21)
    f = _arnaud(saveFilter=True)
    d = str(tmp_path)
    got1, obj1 = filters.filterMaps(mapDicts, _copyFilter(f), 'T0', diagnosticsDir=d, selFnDir=d, verbose=False, returnFilter=True)
    assert obj1._cachePath() is not None
    got2, obj2 = filters.filterMaps(mapDicts, _copyFilter(f), 'T0', diagnosticsDir=d, selFnDir=d, verbose=False,
                                    useCachedFilter=True, returnFilter=True)
    assert obj2.signalNorm == obj1.signalNorm
    # the filtered map is bit-identical; the noise map of the one-bin grid kernel sums its clipped samples in the order in
    # which warps appended them to a list (one atomic per warp), so two runs agree to rounding, not to the bit
    assert np.array_equal(np.asarray(got1['data']), np.asarray(got2['data']))
    assert np.allclose(np.asarray(got1['SNMap']), np.asarray(got2['SNMap']), rtol=1e-12, atol=1e-12)
    # applyFilter on real maps == what buildAndApply filtered
    again = obj2.applyFilter(np.array([m['data'] for m in mapDicts]))
    ref = filters_ref.filterMaps(mapDicts, _copyFilter(_arnaud()), 'T0', undoPixelWindow=False)
    inner = ref['surveyMask'] > 0
    assert np.abs(again[inner] - ref['data'][inner]).max() < 1e-9 * np.abs(ref['data']).max()
    report("cached_filter", ok=1)


def test_files_are_written_in_the_reference_layout(tmp_path, report):
    """The FITS files other parts of Nemo read back (nemo_b200.fitsio): the filter cache with SIGNORM / RW<n> keywords
    (filters.py:767-794) holds exactly the filter and normalisation of the object that wrote it; the class API's RMSMap file
    (filters.py:760-762) and the batched driver's RMSMap / areaMask files (pipelines.py:266-267, 336) hold the maps the filter
    step returned, with the tile's WCS keywords."""
    from nemo_b200 import filters, fitsio, pipelines
    mapDicts, truth = synth.makeTile(shape=(256, 320), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4),
                                     numObjects=6, kind='clusters', cmb=True, seed=23)
    f = _arnaud(saveFilter=True, saveRMSMap=True)
    d = str(tmp_path)
    got, obj = filters.filterMaps(mapDicts, _copyFilter(f), 'T0', diagnosticsDir=d, selFnDir=d, verbose=False, returnFilter=True)
    assert obj.filterFileName.endswith('.fits') and os.path.exists(obj.filterFileName)
    filt, header = fitsio.readImage(obj.filterFileName)
    assert filt.dtype == np.float64 and filt.shape == (2, 256, 320) and np.array_equal(filt, np.asarray(obj.filt))
    assert header['SIGNORM'] == obj.signalNorm
    assert {header['RW1_GHZ']: header['RW1'], header['RW2_GHZ']: header['RW2']} == {float(k): float(v) for k, v in obj.fRelWeights.items()}
    rms, h = fitsio.readImage(os.path.join(d, 'T0', 'RMSMap_%s#T0.fits' % f['label']))
    assert rms.dtype == np.float32 and h['CTYPE1'] == 'RA---CAR' and h['NAXIS1'] == 320
    sn, data = np.asarray(got['SNMap']), np.asarray(got['data'])
    # the class API's file holds the masked RMS map of the filter step: S/N x RMS gives back the (window-convolved) map wherever
    # the S/N is non-zero, so check it through the batched driver's copy, which also exposes that map
    parDict = {'mapFilters': [_copyFilter(_arnaud())], 'photFilter': None, 'thresholdSigma': 4.0, 'minObjPix': 1, 'rejectBorder': 0}
    config = pipelines.SimpleConfig(parDict, {'T0': mapDicts})
    config.selFnDir = os.path.join(d, 'batched')
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, writeAreaMask=True, writeRMSMaps=True)
    rmsB, hB = fitsio.readImage(os.path.join(config.selFnDir, 'T0', 'RMSMap_%s#T0.fits' % f['label']))
    assert np.array_equal(rmsB, rms) and hB['CRVAL1'] == h['CRVAL1']
    area, hA = fitsio.readImage(os.path.join(config.selFnDir, 'areaMask.fits'), extName='T0')
    assert area.dtype == np.uint8 and np.array_equal(area != 0, np.asarray(got['surveyMask']) != 0) and hA['EXTNAME'] == 'T0'
    assert np.array_equal(rms > 0, (area != 0) & (rms > 0)) and (rms > 0).sum() > 0.5 * rms.size
    report("reference_layout_files", ok=1)


def test_reshape_filter_and_apply_to_another_shape(report):
    """reshapeFilter (filters.py:797-821): onto its own grid it is the identity; onto a grid twice as fine in l_x it
    interpolates linearly between neighbouring modes; applyFilter on a map of another shape uses it (filters.py:843-844), and
    for a filter that is smooth in l the result agrees with filtering at the original shape where the two maps overlap."""
    from nemo_b200 import filters
    mapDicts, truth = synth.makeTile(shape=(128, 160), numObjects=3, kind='clusters', seed=31)
    got, obj = filters.filterMaps(mapDicts, _copyFilter(_arnaud(grid=20.)), 'T0', verbose=False, returnFilter=True)
    filt = np.asarray(obj.filt)
    assert np.array_equal(obj.reshapeFilter(filt.shape), filt) and np.array_equal(obj.reshapeFilter(filt.shape[1:]), filt)
    wide = obj.reshapeFilter((1, 128, 320))
    assert wide.shape == (1, 128, 320)
    assert np.allclose(wide[0][:, 0:160:2], filt[0][:, 0:80], rtol=0, atol=1e-12 * np.abs(filt).max())          # the shared modes
    assert np.allclose(wide[0][:, 1:159:2], 0.5 * (filt[0][:, 0:79] + filt[0][:, 1:80]), rtol=0, atol=1e-12 * np.abs(filt).max())
    # a smooth filter: a Gaussian low-pass in |l|
    ly = 2 * np.pi * np.fft.fftfreq(128, d=np.radians(obj.degPerPixY))
    lx = 2 * np.pi * np.fft.fftfreq(160, d=np.radians(obj.degPerPixX))
    obj.filt = np.exp(-0.5 * (ly[:, None] ** 2 + lx[None, :] ** 2) / 2500.0 ** 2)[None]
    obj.signalNorm = 1.0
    rng = np.random.default_rng(3)
    big = rng.standard_normal((1, 192, 240))
    small = big[:, 32:160, 40:200]
    a = obj.applyFilter(small.copy())
    b = obj.applyFilter(big.copy())
    inner = (slice(40, 88), slice(50, 110))
    # enmap.fft carries 1 / sqrt(HW) and the inverse is un-normalised (filters.py:851-857): the same filter applied on a grid
    # of another size comes out sqrt(H'W' / HW) larger -- the reference's own note next to its padding code (signals.py:987)
    b = b / np.sqrt(192 * 240 / (128 * 160.0))
    err = float(np.abs(a[inner] - b[32:160, 40:200][inner]).max() / np.abs(a[inner]).max())
    report("reshape_filter", rel_diff_inner=err)
    assert err < 2e-2


def test_real_space_matched_filter_two_frequency(report):
    """BASELINE config 3 (one reduced tile): 90 + 150 GHz real-space kernel, symmetrised, bckSub."""
    from nemo_b200 import filters, photometry
    mapDicts, truth = synth.makeTile(shape=(512, 512), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4),
                                     numObjects=10, kind='clusters', cmb=True, seed=31)
    p = {'M500MSun': 2e14, 'z': 0.4, 'outputUnits': 'yc', 'bckSub': True, 'edgeTrimArcmin': 0.,
         'noiseParams': {'method': 'dataMap', 'matchedFilterClass': 'ArnaudModelMatchedFilter', 'RADecSection': None,
                         'kernelMaxArcmin': 7., 'symmetrize': True, 'noiseGridArcmin': 20.}}
    f = {'label': 'ArnaudRS', 'class': 'ArnaudModelRealSpaceMatchedFilter', 'params': p}
    ref, refObj = filters_ref.filterMaps(mapDicts, _copyFilter(f), 'T0', returnFilter=True)
    got, gotObj = filters.filterMaps(mapDicts, _copyFilter(f), 'T0', verbose=False, returnFilter=True)
    assert gotObj.kern2d.shape == refObj.kern2d.shape
    assert np.abs(gotObj.kern2d - refObj.kern2d).max() <= 1e-8 * np.abs(refObj.kern2d).max()
    assert abs(gotObj.bckSubScaleArcmin - refObj.bckSubScaleArcmin) < 1e-12
    assert abs(gotObj.signalNorm / refObj.signalNorm - 1) < 1e-8
    _compareMaps(got, ref, report, "realspace_F2")
    refCat = photometry_ref.findObjects(ref, **DETECT)
    photometry_ref.measureFluxes(refCat, ref, None)
    gotCat = photometry.findObjects(got, **DETECT)
    photometry.measureFluxes(gotCat, got, None)
    _compareCatalogs(gotCat, refCat, ['SNR', 'y_c'])


def test_batched_pipeline_matches_per_tile_oracle(report):
    """Tile x scale batching: 3 tiles x 3 scales, photFilter first, optimal catalog."""
    from nemo_b200 import pipelines, catalogs
    tileMaps = {}
    for t in range(3):
        md, _ = synth.makeTile(shape=(256, 256), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4),
                               numObjects=6, kind='clusters', cmb=True, seed=100 + t)
        tileMaps['tile%d' % t] = md
    filtersList = [_arnaud('Arnaud_M1e14_z0p8', 1e14, 0.8), _arnaud('Arnaud_M2e14_z0p4', 2e14, 0.4), _arnaud('Arnaud_M8e14_z0p2', 8e14, 0.2)]
    parDict = {'mapFilters': filtersList, 'photFilter': 'Arnaud_M2e14_z0p4', 'thresholdSigma': 4.0, 'minObjPix': 1,
               'findCenterOfMass': True, 'useInterpolator': True, 'rejectBorder': 0}
    config = pipelines.SimpleConfig(parDict, tileMaps)
    opt = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=[1, 2])     # batches of 1 and 2 tiles
    # oracle: the reference's loop order (pipelines.py:179-286)
    ordered = pipelines._orderedFilters(config.parDict)
    assert ordered[0]['label'] == 'Arnaud_M2e14_z0p4'
    nObj = 0
    for tileName, md in tileMaps.items():
        phot = None
        for f in ordered:
            ref = filters_ref.filterMaps(md, _copyFilter(f), tileName)
            if f['label'] == parDict['photFilter']:
                phot = {'SNMap': ref['SNMap'], 'data': ref['data']}
            refCat = photometry_ref.findObjects(ref, threshold=4.0, minObjPix=1, findCenterOfMass=True, removeRings=True,
                                                ringThresholdSigma=3, rejectBorder=0, useInterpolator=True)
            photometry_ref.measureFluxes(refCat, ref, None, photFilteredMapDict=phot)
            gotTab = config.catalogDict[f['label'] + "#" + tileName]['catalog']
            _compareCatalogs(gotTab, refCat, ['SNR', 'y_c', 'err_y_c', 'fixed_SNR', 'fixed_y_c', 'deltaT_c'])
            nObj += len(refCat)
    assert nObj >= 6
    assert len(opt) > 0 and len(opt) <= nObj
    # the driver merges the batches' stacked rows; stacking the per-image tables (the reference's vstack) gives the same catalog
    optPerImage = catalogs.makeOptimalCatalog(config.catalogDict)
    assert len(optPerImage) == len(opt)
    for k in optPerImage.keys():
        a, b = np.asarray(opt[k]), np.asarray(optPerImage[k])
        assert np.array_equal(a, b), k
    report("batched_pipeline", objects=nObj, optimal=len(opt))


def test_batched_real_space_driver(report, monkeypatch):
    """Real-space matched filters through the batched driver (nemo_b200/realspace.py: kernels of all tiles of a batch built and
    applied together; BASELINE configs[2] shape, reduced): catalogs against the oracle's per-tile loop
    (filters.py:862-1218) and against this library's own per-tile class loop; two filter scales with a photFilter."""
    from nemo_b200 import pipelines, realspace
    tileMaps = {}
    for t in range(3):
        md, _ = synth.makeTile(shape=(256, 256), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4),
                               numObjects=6, kind='clusters', cmb=True, seed=170 + t)
        tileMaps['tile%d' % t] = md

    def rs(label, M, z):
        p = {'M500MSun': M, 'z': z, 'outputUnits': 'yc', 'bckSub': True, 'bckSubScaleArcmin': 14.0, 'edgeTrimArcmin': 0.,
             'noiseParams': {'method': 'dataMap', 'matchedFilterClass': 'ArnaudModelMatchedFilter', 'RADecSection': None,
                             'kernelMaxArcmin': 7., 'symmetrize': True, 'noiseGridArcmin': 20.}}
        return {'label': label, 'class': 'ArnaudModelRealSpaceMatchedFilter', 'params': p}
    filtersList = [rs('RS_M2e14_z0p4', 2e14, 0.4), rs('RS_M4e14_z0p3', 4e14, 0.3)]
    parDict = {'mapFilters': filtersList, 'photFilter': 'RS_M2e14_z0p4', 'thresholdSigma': 4.0, 'minObjPix': 1, 'findCenterOfMass': True,
               'useInterpolator': True, 'rejectBorder': 0}
    assert realspace.canBatch(filtersList, tileMaps)
    config = pipelines.SimpleConfig(parDict, tileMaps)
    opt = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=2)
    batchedDict = config.catalogDict
    # the oracle, in the reference's loop order
    nObj = 0
    for tileName, md in tileMaps.items():
        phot = None
        for f in pipelines._orderedFilters(config.parDict):
            ref = filters_ref.filterMaps(md, _copyFilter(f), tileName)
            if f['label'] == parDict['photFilter']:
                phot = {'SNMap': ref['SNMap'], 'data': ref['data']}
            refCat = photometry_ref.findObjects(ref, **DETECT)
            photometry_ref.measureFluxes(refCat, ref, None, photFilteredMapDict=phot)
            _compareCatalogs(batchedDict[f['label'] + "#" + tileName]['catalog'], refCat, ['SNR', 'y_c', 'err_y_c', 'fixed_SNR', 'fixed_y_c'])
            nObj += len(refCat)
    assert nObj >= 4 and 0 < len(opt) <= nObj
    # the per-tile class loop of this library gives the same catalogs
    monkeypatch.setattr(realspace, 'canBatch', lambda *a, **k: False)
    config2 = pipelines.SimpleConfig(parDict, tileMaps)
    opt2 = pipelines.filterMapsAndMakeCatalogs(config2, verbose=False)
    assert len(opt2) == len(opt)
    for key, entry in config2.catalogDict.items():
        a, b = entry['catalog'], batchedDict[key]['catalog']
        assert len(a) == len(b)
        if len(a) > 0:
            for col in ['SNR', 'y_c', 'fixed_SNR', 'x', 'y']:
                assert np.allclose(np.asarray(a[col], dtype=float), np.asarray(b[col], dtype=float), rtol=1e-8, atol=1e-8), (key, col)
    report("batched_real_space", objects=nObj, optimal=len(opt))


@pytest.mark.parametrize("findCenterOfMass", [True, False])
def test_find_objects_ring_removal_and_options(report, findCenterOfMass):
    """photometry.findObjects on hand-made S/N maps: detections sitting on > 30-pixel rings at ringThresholdSigma are
    dropped (photometry.py:60-70, 104-105), others kept; centre-of-mass and maximum-position modes, invertMap, minObjPix."""
    from scipy import ndimage
    from nemo_b200 import photometry
    from oracle.wcs_stub import SimpleCARWCS
    rng = np.random.default_rng(8)
    H, W = 256, 320
    sn = ndimage.gaussian_filter(rng.standard_normal((H, W)), 1.2)
    sn /= sn.std()
    yy, xx = np.mgrid[:H, :W]
    for cy, cx in [(60, 70), (180, 240)]:                       # two rings (artefacts around bright sources)
        r = np.hypot(yy - cy, xx - cx)
        sn[(r > 10) & (r < 14)] = 3.6
        # a "detection" sitting on each ring (a peaked bump: exact ties would make scipy's maximum_position arbitrary)
        sn = np.maximum(sn, 7.0 * np.exp(-0.5 * ((yy - cy - 0.3) ** 2 + (xx - cx - 12.2) ** 2) / 1.5 ** 2))
    for cy, cx, a in [(120, 100, 9.0), (40, 250, 6.0), (200, 60, 12.0)]:
        sn += a * np.exp(-0.5 * ((yy - cy) ** 2 + (xx - cx) ** 2) / 2.0 ** 2)
    # as in a real filtered map (apodisation mask, filters.py:749-750) nothing survives near the tile edge; the GPU spline
    # differs from FITPACK's not-a-knot spline only there
    sn[:24, :] = 0
    sn[-24:, :] = 0
    sn[:, :24] = 0
    sn[:, -24:] = 0
    wcs = SimpleCARWCS((H, W), 0.5 / 60)
    wcs.header['BUNIT'] = 'yc'
    fmd = {'SNMap': sn, 'data': sn * 1e-5, 'wcs': wcs, 'surveyMask': np.ones((H, W)), 'flagMask': np.zeros((H, W), dtype=int),
           'label': 'F', 'tileName': 'T', 'mapUnits': 'yc', 'obsFreqGHz': 148.0, 'beamSolidAngle_nsr': 0.0}
    for kw in [dict(), dict(removeRings=False), dict(minObjPix=6), dict(invertMap=True, threshold=2.5, ringThresholdSigma=2.0)]:
        opts = dict(threshold=4.0, minObjPix=1, findCenterOfMass=findCenterOfMass, removeRings=True, ringThresholdSigma=3,
                    useInterpolator=True, rejectBorder=0)
        opts.update(kw)
        ref = photometry_ref.findObjects(fmd, **opts)
        got = photometry.findObjects(fmd, **opts)
        _compareCatalogs(got, ref, ['SNR'])
        if not kw:
            plain = photometry_ref.findObjects(fmd, **dict(opts, removeRings=False))
            assert len(plain) == len(ref) + 2, "the two detections on the rings should be the only ones removed"
    report("find_objects_rings", com=int(findCenterOfMass), objects=len(ref))


def test_injection_path_cached_filters_and_rms(tmp_path, report):
    """BASELINE config 5 (reduced): sources are injected into the maps and the tiles re-filtered with the cached filters
    and the photFilter's cached RMS map (reference maps.py:2093 -> pipelines.py:216-232), against the oracle doing the
    reference's sequence with its own cached filter files."""
    from nemo_b200 import pipelines
    from oracle import pixell_shim as enmap
    d = str(tmp_path)
    tileMaps = {}
    for t in range(2):
        md, _ = synth.makeTile(shape=(256, 256), numObjects=4, kind='clusters', seed=300 + t)
        tileMaps['tile%d' % t] = md
    filtersList = [_arnaud('Arnaud_M2e14_z0p4', 2e14, 0.4, saveFilter=True), _arnaud('Arnaud_M6e14_z0p3', 6e14, 0.3, saveFilter=True)]
    parDict = {'mapFilters': filtersList, 'photFilter': 'Arnaud_M2e14_z0p4', 'thresholdSigma': 4.0, 'minObjPix': 1,
               'findCenterOfMass': True, 'useInterpolator': True, 'rejectBorder': 0}
    config = pipelines.SimpleConfig(parDict, tileMaps)
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=2, cacheFilters=True)
    # oracle, first pass: filters and RMS maps of the real maps
    ordered = pipelines._orderedFilters(config.parDict)
    refRMS = {}
    for tileName, md in tileMaps.items():
        for f in ordered:
            _, obj = filters_ref.filterMaps(md, _copyFilter(f), tileName, diagnosticsDir=d, selFnDir=d, returnFilter=True)
            if f['label'] == parDict['photFilter']:
                refRMS[tileName] = obj.RMSMap
    # inject: a few beam-convolved decrements per tile, directly into the map arrays
    rng = np.random.default_rng(4)
    yy, xx = np.mgrid[:256, :256]
    for tileName, md in tileMaps.items():
        for _ in range(3):
            cy, cx = rng.uniform(60, 196, 2)
            md[0]['data'] = md[0]['data'] - 450.0 * np.exp(-0.5 * ((yy - cy) ** 2 + (xx - cx) ** 2) / 2.5 ** 2)
    opt = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=2, useCachedFilters=True, useCachedRMSMap=True,
                                              restage=True)
    nObj = 0
    for tileName, md in tileMaps.items():
        phot = None
        for f in ordered:
            ref = filters_ref.filterMaps(md, _copyFilter(f), tileName, diagnosticsDir=d, selFnDir=d, useCachedFilter=True,
                                         undoPixelWindow=False)
            RMSMap = refRMS[tileName]
            validMask = np.greater(RMSMap, 0)
            SNMap = np.zeros(ref['data'].shape) + ref['data']
            SNMap[validMask] = SNMap[validMask] / RMSMap[validMask]
            ref['SNMap'] = SNMap
            mask = np.equal(ref['data'], 0)
            ref['data'] = enmap.apply_window(ref['data'], pow=-1.0)
            ref['data'][mask] = 0
            if f['label'] == parDict['photFilter']:
                phot = {'SNMap': ref['SNMap'], 'data': ref['data']}
            refCat = photometry_ref.findObjects(ref, threshold=4.0, minObjPix=1, findCenterOfMass=True, removeRings=True,
                                                ringThresholdSigma=3, rejectBorder=0, useInterpolator=True)
            photometry_ref.measureFluxes(refCat, ref, None, photFilteredMapDict=phot)
            gotTab = config.catalogDict[f['label'] + "#" + tileName]['catalog']
            _compareCatalogs(gotTab, refCat, ['SNR', 'y_c', 'err_y_c', 'fixed_SNR', 'fixed_y_c', 'deltaT_c'])
            nObj += len(refCat)
    assert nObj >= 8 and len(opt) >= 6
    report("injection_cached", objects=nObj, optimal=len(opt))


def test_fp32_storage_mode_against_fp64_oracle(report):
    """precision='f32' (fp32 maps / spectra / FFT arithmetic; covariance, solve, reductions and splines in fp64) on maps that
    are exactly representable in fp32, against the fp64 oracle on the same values: the north-star tolerances (S/N maps 1e-4,
    positions 0.01 px) must still hold, and this test records by how much."""
    from nemo_b200 import filters, photometry
    mapDicts, truth = synth.makeTile(shape=(512, 512), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4),
                                     numObjects=12, kind='clusters', cmb=True, seed=11)
    for m in mapDicts:
        m['data'] = m['data'].astype(np.float32).astype(np.float64)
        m['weights'] = m['weights'].astype(np.float32).astype(np.float64)
    f = _arnaud('Arnaud_M2e14_z0p4', 2e14, 0.4, grid=40.)
    ref = filters_ref.filterMaps(mapDicts, _copyFilter(f), 'T0')
    f32 = _copyFilter(f)
    f32['params']['precision'] = 'f32'
    got = filters.filterMaps(mapDicts, f32, 'T0', verbose=False)
    sn, snRef = np.asarray(got['SNMap'], dtype=np.float64), ref['SNMap']
    assert np.array_equal(np.asarray(got['surveyMask']) != 0, ref['surveyMask'] != 0)
    errSN = float(np.abs(sn - snRef).max())
    relData = float(np.abs(np.asarray(got['data'], dtype=np.float64) - ref['data']).max() / np.abs(ref['data']).max())
    report("fp32_mode", maxabs_dSN=errSN, rel_dData=relData)
    assert np.allclose(sn, snRef, rtol=1e-4, atol=1e-4), "S/N maps differ: %g" % errSN
    refCat = photometry_ref.findObjects(ref, **DETECT)
    gotCat = photometry.findObjects(got, **DETECT)
    assert len(gotCat) == len(refCat) and len(refCat) >= 3
    for a, b in zip(gotCat.toRows(), refCat):
        assert abs(a['x'] - b['x']) < 0.01 and abs(a['y'] - b['y']) < 0.01 and abs(a['SNR'] - b['SNR']) < 1e-3 * abs(b['SNR'])


def test_non_power_of_two_tile_end_to_end(report):
    """A 300 x 420 tile (real Nemo tiles are not powers of two, e.g. 1440 x 840): the whole path against the oracle."""
    from nemo_b200 import filters, photometry
    mapDicts, truth = synth.makeTile(shape=(300, 420), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4),
                                     numObjects=8, kind='clusters', cmb=True, seed=41, margin=50)
    f = _arnaud('Arnaud_M2e14_z0p4', 2e14, 0.4, grid=20.)
    ref, refObj = filters_ref.filterMaps(mapDicts, _copyFilter(f), 'T0', returnFilter=True)
    got, gotObj = filters.filterMaps(mapDicts, _copyFilter(f), 'T0', verbose=False, returnFilter=True)
    _compareMaps(got, ref, report, "arnaud_F2_300x420")
    assert abs(gotObj.signalNorm / refObj.signalNorm - 1) < 1e-8
    refCat = photometry_ref.findObjects(ref, **DETECT)
    photometry_ref.measureFluxes(refCat, ref, None)
    gotCat = photometry.findObjects(got, **DETECT)
    photometry.measureFluxes(gotCat, got, None)
    assert len(refCat) >= 3
    _compareCatalogs(gotCat, refCat, ['SNR', 'y_c', 'err_y_c', 'deltaT_c'])


def test_odd_sided_tile_end_to_end(report):
    """A 301 x 421 tile (the reference's autotiler produces arbitrary sides, maps.py:691-791): odd width -> (W + 1) / 2 spectrum
    columns and no Nyquist column.  Beam filter on two frequencies (its template is painted on radiansMap by the oracle and the
    product alike; the oracle's simplified Arnaud painter centres on radiansMap too, while the product follows the reference and
    paints at wcs.getCentreWCSCoords(), half a pixel away on an odd side).  Class API against the oracle, and the batched driver
    against the same catalog."""
    from nemo_b200 import filters, photometry, pipelines
    mapDicts, truth = synth.makeTile(shape=(301, 421), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4),
                                     numObjects=14, kind='sources', cmb=True, seed=43, margin=50)
    f = _beam(edgeTrimArcmin=5.0)
    ref, refObj = filters_ref.filterMaps(mapDicts, _copyFilter(f), 'T0', returnFilter=True)
    got, gotObj = filters.filterMaps(mapDicts, _copyFilter(f), 'T0', verbose=False, returnFilter=True)
    _compareMaps(got, ref, report, "beam_F2_301x421")
    assert abs(gotObj.signalNorm / refObj.signalNorm - 1) < 1e-8
    refCat = photometry_ref.findObjects(ref, **DETECT)
    photometry_ref.measureFluxes(refCat, ref, None)
    gotCat = photometry.findObjects(got, **DETECT)
    photometry.measureFluxes(gotCat, got, None)
    assert len(refCat) >= 3
    _compareCatalogs(gotCat, refCat, ['SNR', 'deltaT_c', 'err_deltaT_c'])
    parDict = {'mapFilters': [_copyFilter(f)], 'photFilter': None, 'thresholdSigma': DETECT['threshold'], 'minObjPix': DETECT['minObjPix'],
               'findCenterOfMass': DETECT['findCenterOfMass'], 'useInterpolator': DETECT['useInterpolator'],
               'rejectBorder': DETECT['rejectBorder'], 'removeRings': DETECT['removeRings'],
               'ringThresholdSigma': DETECT['ringThresholdSigma']}
    config = pipelines.SimpleConfig(parDict, {'T0': mapDicts})
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False)
    _compareCatalogs(config.catalogDict[f['label'] + '#T0']['catalog'], refCat, ['SNR', 'deltaT_c', 'err_deltaT_c'])


def test_bench_workload_parity_at_full_tile_size(report):
    """The workload bench.py measures (BASELINE configs[1]: 1024 x 1024 tiles, Beam filter, 8 weight bins, S/N > 4) through
    the batched, graph-replayed pipeline, against the oracle on two of its tiles: catalogs identical, noise tables equal."""
    import bench_workloads
    from nemo_b200 import pipelines
    parDict, tileMaps, truth = bench_workloads.pointSourceConfig(numTiles=4)
    config = pipelines.SimpleConfig(parDict, tileMaps)
    for _ in range(2):      # the second pass replays the captured graphs
        opt = pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=2)
    f = parDict['mapFilters'][0]
    nObj = 0
    for tileName in ['0_0', '0_3']:
        ref = filters_ref.filterMaps(tileMaps[tileName], _copyFilter(f), tileName)
        refCat = photometry_ref.findObjects(ref, threshold=4.0, minObjPix=1, findCenterOfMass=True, removeRings=True,
                                            ringThresholdSigma=3, rejectBorder=0, useInterpolator=True, objIdent='MOCK-S')
        photometry_ref.measureFluxes(refCat, ref, None)
        gotTab = config.catalogDict[f['label'] + "#" + tileName]['catalog']
        _compareCatalogs(gotTab, refCat, ['SNR', 'deltaT_c', 'err_deltaT_c', 'fluxJy', 'err_fluxJy'])
        nObj += len(refCat)
        # the injected sources are recovered: every bright one has a detection within 1.5 px
        t = truth[tileName]
        bright = t['amp'] > 1500
        got = gotTab.toRows()
        for x, y in zip(t['x'][bright], t['y'][bright]):
            assert min(np.hypot(r['x'] - x, r['y'] - y) for r in got) < 1.5
    assert nObj > 100 and len(opt) > 200
    report("bench_workload_parity", objects=nObj, optimal=len(opt))


def test_three_frequencies_half_masked_tile_and_empty_catalog(report):
    """Three frequency maps (3 x 3 solve per mode), half of the tile masked out and zeroed, and a threshold nothing reaches."""
    from nemo_b200 import filters, photometry
    mapDicts, truth = synth.makeTile(shape=(256, 256), freqsGHz=(97.8, 149.6, 220.0), noiseUK=(60.0, 40.0, 90.0),
                                     beamFWHMArcmin=(2.2, 1.4, 1.0), numObjects=6, kind='clusters', cmb=True, seed=61)
    for m in mapDicts:
        m['surveyMask'] = m['surveyMask'].copy()
        m['surveyMask'][:, 128:] = 0
        m['data'] = m['data'] * m['surveyMask']
        m['weights'] = m['weights'] * m['surveyMask']
    f = _arnaud('Arnaud_M2e14_z0p4', 2e14, 0.4, grid=20., edgeTrimArcmin=8.0)
    ref = filters_ref.filterMaps(mapDicts, _copyFilter(f), 'T0')
    got = filters.filterMaps(mapDicts, _copyFilter(f), 'T0', verbose=False)
    _compareMaps(got, ref, report, "arnaud_F3_half_masked")
    assert not np.asarray(got['surveyMask'])[:, 128:].any()
    refCat = photometry_ref.findObjects(ref, **DETECT)
    gotCat = photometry.findObjects(got, **DETECT)
    _compareCatalogs(gotCat, refCat, ['SNR'])
    assert photometry.findObjects(got, **dict(DETECT, threshold=1e6)) == []
    assert photometry_ref.findObjects(ref, **dict(DETECT, threshold=1e6)) == []


def test_pipeline_driver_with_real_space_filters(report):
    """filterMapsAndMakeCatalogs with a real-space filter set (BASELINE config 3 shape, reduced): the driver falls back to the
    reference's per-tile loop over the class API; catalogs against the oracle's loop."""
    from nemo_b200 import pipelines
    tileMaps = {}
    for t in range(2):
        md, _ = synth.makeTile(shape=(256, 256), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4),
                               numObjects=6, kind='clusters', cmb=True, seed=70 + t)
        tileMaps['tile%d' % t] = md
    p = {'M500MSun': 2e14, 'z': 0.4, 'outputUnits': 'yc', 'bckSub': True, 'edgeTrimArcmin': 0.,
         'noiseParams': {'method': 'dataMap', 'matchedFilterClass': 'ArnaudModelMatchedFilter', 'RADecSection': None,
                         'kernelMaxArcmin': 7., 'symmetrize': True, 'noiseGridArcmin': 20.}}
    f = {'label': 'ArnaudRS', 'class': 'ArnaudModelRealSpaceMatchedFilter', 'params': p}
    parDict = {'mapFilters': [f], 'photFilter': None, 'thresholdSigma': 4.0, 'minObjPix': 1, 'findCenterOfMass': True,
               'useInterpolator': True, 'rejectBorder': 0}
    config = pipelines.SimpleConfig(parDict, tileMaps)
    opt = pipelines.filterMapsAndMakeCatalogs(config, verbose=False)
    nObj = 0
    for tileName, md in tileMaps.items():
        ref = filters_ref.filterMaps(md, _copyFilter(f), tileName)
        refCat = photometry_ref.findObjects(ref, **DETECT)
        photometry_ref.measureFluxes(refCat, ref, None)
        _compareCatalogs(config.catalogDict['ArnaudRS#' + tileName]['catalog'], refCat, ['SNR', 'y_c', 'err_y_c'])
        nObj += len(refCat)
    assert nObj >= 2 and len(opt) == nObj
    report("pipeline_real_space", objects=nObj)


def test_float32_host_tiles_are_widened_exactly(report):
    """Tiles handed over as float32 (the on-disk format) are staged and uploaded as float32 and widened on the device: the
    catalogs equal those of the same values handed over as float64."""
    from nemo_b200 import pipelines
    import torch
    tiles32, tiles64 = {}, {}
    for t in range(3):
        md, _ = synth.makeTile(shape=(256, 256), numObjects=6, kind='clusters', seed=300 + t)
        m32 = [dict(m, data=np.asarray(m['data'], dtype=np.float32), weights=np.asarray(m['weights'], dtype=np.float32)) for m in md]
        tiles32['tile%d' % t] = m32
        tiles64['tile%d' % t] = [dict(m, data=np.asarray(m['data'], dtype=np.float64), weights=np.asarray(m['weights'], dtype=np.float64))
                                 for m in m32]
    parDict = {'mapFilters': [_arnaud('Arnaud_M2e14_z0p4', 2e14, 0.4)], 'photFilter': None, 'thresholdSigma': 4.0, 'minObjPix': 1,
               'findCenterOfMass': True, 'useInterpolator': True, 'rejectBorder': 0}
    c32 = pipelines.SimpleConfig(parDict, tiles32)
    c64 = pipelines.SimpleConfig(parDict, tiles64)
    opt32 = pipelines.filterMapsAndMakeCatalogs(c32, verbose=False, tilesPerBatch=2)
    opt64 = pipelines.filterMapsAndMakeCatalogs(c64, verbose=False, tilesPerBatch=2)
    assert c32._staged['batches'][0].host['data'].dtype == torch.float32
    assert c64._staged['batches'][0].host['data'].dtype == torch.float64
    assert c32._staged['batches'][0].h2dBytes < 0.6 * c64._staged['batches'][0].h2dBytes
    assert len(opt32) == len(opt64) and len(opt64) >= 3
    # the widening is exact, so the two runs see the same doubles; the per-object moments are accumulated with atomics (their
    # order varies from run to run), hence a last-digits tolerance rather than bit equality
    for k in ['x', 'y', 'SNR', 'y_c', 'RADeg', 'decDeg']:
        assert np.allclose(np.asarray(opt32[k], dtype=float), np.asarray(opt64[k], dtype=float), rtol=1e-12, atol=1e-12), k
    assert np.array_equal(np.asarray(opt32['numSigPix'], dtype=float), np.asarray(opt64['numSigPix'], dtype=float))
    report("float32-staged tiles: %d objects equal (1e-12) to the float64-staged run" % len(opt64))


def test_smart_noise_option_end_to_end(report):
    """noiseGridArcmin: 'smart' with numNoiseBins (the reference's examples/sources/PS_S18d_f220_202006.yml set-up, 'dataMap'
    noise): whole-map weight bins instead of grid cells, through filterMaps and through the batched driver."""
    from nemo_b200 import filters, photometry, pipelines
    tileMaps = {}
    for t in range(2):
        md, _ = synth.makeTile(shape=(256, 256), freqsGHz=(148.0,), noiseUK=(100.0,), numObjects=12, kind='sources', seed=700 + t,
                               borderPix=20)
        yy, xx = np.mgrid[0:256, 0:256]
        md[0]['weights'] = md[0]['weights'] * (0.7 + 0.3 * np.sin(yy / 40.0) * np.cos(xx / 55.0))
        tileMaps['t%d' % t] = md
    f = _beam(grid='smart', numBins=12, edgeTrimArcmin=0.0)
    for name, md in tileMaps.items():
        ref = filters_ref.filterMaps(md, _copyFilter(f), name)
        got = filters.filterMaps(md, _copyFilter(f), name, verbose=False)
        _compareMaps(got, ref, report, "beam_smart_" + name)
        assert (ref['SNMap'] != 0).mean() > 0.5
    parDict = {'mapFilters': [f], 'photFilter': None, 'thresholdSigma': 4.0, 'minObjPix': 1, 'findCenterOfMass': True,
               'useInterpolator': True, 'rejectBorder': 0}
    config = pipelines.SimpleConfig(parDict, tileMaps)
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=2)
    nObj = 0
    for name, md in tileMaps.items():
        ref = filters_ref.filterMaps(md, _copyFilter(f), name)
        refCat = photometry_ref.findObjects(ref, **DETECT)
        photometry_ref.measureFluxes(refCat, ref, None)
        _compareCatalogs(config.catalogDict[f['label'] + "#" + name]['catalog'], refCat, ['SNR', 'deltaT_c'])
        nObj += len(refCat)
    assert nObj >= 6


def test_rms_area_tables_match_reference_getRMSTab(report):
    """completeness.getRMSTab (nemo/completeness.py:981-999): noise level versus area of a filtered tile, from the class API
    (grid and 'smart' noise) and from the batched driver (makeRMSTables), against the oracle's np.unique / boolean-sum version
    on the oracle's own RMS map and area mask."""
    from nemo_b200 import filters, pipelines, completeness
    from oracle import completeness_ref
    tileMaps = {}
    for t in range(3):
        md, _ = synth.makeTile(shape=(256, 256), freqsGHz=(148.0,), noiseUK=(100.0,), numObjects=8, kind='sources', seed=900 + t,
                               borderPix=24)
        yy, xx = np.mgrid[0:256, 0:256]
        md[0]['weights'] = md[0]['weights'] * (0.7 + 0.3 * np.sin(yy / 33.0 + t) * np.cos(xx / 47.0))
        tileMaps['t%d' % t] = md
    for f in [_beam(grid=20., numBins=4), _beam(grid='smart', numBins=9)]:
        for name, md in tileMaps.items():
            ref, refObj = filters_ref.filterMaps(md, _copyFilter(f), name, returnFilter=True)
            got, gotObj = filters.filterMaps(md, _copyFilter(f), name, verbose=False, returnFilter=True)
            # the reference makes its table from the masked RMS map and the area mask the filter step wrote
            values, areas = completeness_ref.getRMSTab(refObj.RMSMap, ref['surveyMask'], md[0]['wcs'])
            tab = completeness.getRMSTab(gotObj.RMSMap, got['surveyMask'], md[0]['wcs'], gotObj.noiseTable[0])
            assert len(tab) == len(values)
            assert np.allclose(np.asarray(tab['y0RMS']), values, rtol=1e-11, atol=0)
            assert np.allclose(np.asarray(tab['areaDeg2']), areas, rtol=1e-11, atol=0)
            assert len(values) >= 3
        parDict = {'mapFilters': [f], 'photFilter': None, 'thresholdSigma': 4.0, 'minObjPix': 1, 'findCenterOfMass': True,
                   'useInterpolator': True, 'rejectBorder': 0}
        config = pipelines.SimpleConfig(parDict, tileMaps)
        pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=2, makeRMSTables=True)
        assert sorted(config.RMSTabs.keys()) == sorted(f['label'] + "#" + n for n in tileMaps)
        for name, md in tileMaps.items():
            ref, refObj = filters_ref.filterMaps(md, _copyFilter(f), name, returnFilter=True)
            values, areas = completeness_ref.getRMSTab(refObj.RMSMap, ref['surveyMask'], md[0]['wcs'])
            bv, ba = config.RMSTabs[f['label'] + "#" + name]
            assert len(bv) == len(values)
            assert np.allclose(bv, values, rtol=1e-11, atol=0) and np.allclose(ba, areas, rtol=1e-11, atol=0)
        report("rms_area_tables %s: %d levels in the last tile, %.3f deg2" % (f['params']['noiseParams']['noiseGridArcmin'], len(values), areas.sum()))


@pytest.mark.parametrize("estimator", ['percentile', 'biweight'])
def test_rms_estimator_end_to_end(report, estimator):
    """noiseParams RMSEstimator: 'percentile' / 'biweight' (filters.py:460-465) through filterMaps and the batched driver."""
    from nemo_b200 import filters, pipelines
    tileMaps = {}
    for t in range(2):
        md, _ = synth.makeTile(shape=(256, 256), freqsGHz=(148.0,), noiseUK=(100.0,), numObjects=10, kind='sources', seed=800 + t,
                               borderPix=20)
        tileMaps['t%d' % t] = md
    f = _beam(grid=20., numBins=2, edgeTrimArcmin=0.0)
    f['params']['noiseParams']['RMSEstimator'] = estimator
    for name, md in tileMaps.items():
        ref = filters_ref.filterMaps(md, _copyFilter(f), name)
        got = filters.filterMaps(md, _copyFilter(f), name, verbose=False)
        _compareMaps(got, ref, report, "beam_" + estimator + "_" + name)
    parDict = {'mapFilters': [f], 'photFilter': None, 'thresholdSigma': 4.0, 'minObjPix': 1, 'findCenterOfMass': True,
               'useInterpolator': True, 'rejectBorder': 0}
    config = pipelines.SimpleConfig(parDict, tileMaps)
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=2)
    nObj = 0
    for name, md in tileMaps.items():
        ref = filters_ref.filterMaps(md, _copyFilter(f), name)
        refCat = photometry_ref.findObjects(ref, **DETECT)
        photometry_ref.measureFluxes(refCat, ref, None)
        _compareCatalogs(config.catalogDict[f['label'] + "#" + name]['catalog'], refCat, ['SNR', 'deltaT_c'])
        nObj += len(refCat)
    assert nObj >= 4


def test_dr5_tile_all_sixteen_scales_against_oracle(report):
    """BASELINE configs[3] at full size: one 840 x 1440 two-frequency tile through the batched driver at all 16 Arnaud_M*_z*
    scales with photFilter Arnaud_M2e14_z0p4, every per-scale catalog against the oracle in the reference's loop order."""
    import bench_workloads
    from nemo_b200 import pipelines
    parDict, tileMaps, _ = bench_workloads.dr5Config(numTiles=1, firstIndex=57)       # third declination row
    config = pipelines.SimpleConfig(parDict, tileMaps)
    opt = pipelines.filterMapsAndMakeCatalogs(config, verbose=False)
    ordered = pipelines._orderedFilters(config.parDict)
    assert len(ordered) == 16 and ordered[0]['label'] == 'Arnaud_M2e14_z0p4'
    nObj = 0
    for tileName, md in tileMaps.items():
        phot = None
        for f in ordered:
            ref = filters_ref.filterMaps(md, _copyFilter(f), tileName)
            if f['label'] == parDict['photFilter']:
                phot = {'SNMap': ref['SNMap'], 'data': ref['data']}
            refCat = photometry_ref.findObjects(ref, **dict(DETECT, objIdent='ACT-CL'))
            photometry_ref.measureFluxes(refCat, ref, None, photFilteredMapDict=phot)
            _compareCatalogs(config.catalogDict[f['label'] + "#" + tileName]['catalog'], refCat,
                             ['SNR', 'y_c', 'err_y_c', 'fixed_SNR', 'fixed_y_c', 'deltaT_c'])
            nObj += len(refCat)
    assert nObj >= 16 * 15 and len(opt) >= 15
    report("dr5_sixteen_scales", objects=nObj, optimal=len(opt))


def test_tiles_at_two_declinations_get_their_own_filters(report):
    """CAR tiles at different declinations have different x pixel scales, hence different radiansMap, templates, filters and
    normalisation (filters.py:214-239, 617, 642-646).  Four tiles, two at dec 0 and two at dec -50, interleaved: the batched driver
    groups them by geometry and every catalog matches the per-tile oracle."""
    from nemo_b200 import pipelines
    from nemo_b200.wcs import CARWCS
    tileMaps = {}
    for t, dec in enumerate([0.0, -50.0, 0.0, -50.0]):
        md, _ = synth.makeTile(shape=(256, 320), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4),
                               numObjects=6, kind='clusters', seed=300 + t)
        wcs = CARWCS((256, 320), 0.5 / 60, crval=(20.0 + t, dec))
        for m in md:
            m['wcs'] = wcs
        tileMaps['tile%d' % t] = md
    filtersList = [_arnaud('Arnaud_M2e14_z0p4', 2e14, 0.4, grid=20.), _arnaud('Arnaud_M8e14_z0p2', 8e14, 0.2, grid=20.)]
    parDict = {'mapFilters': filtersList, 'photFilter': 'Arnaud_M2e14_z0p4', 'thresholdSigma': 4.0, 'minObjPix': 1,
               'findCenterOfMass': True, 'useInterpolator': True, 'rejectBorder': 0}
    config = pipelines.SimpleConfig(parDict, tileMaps)
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=4)
    assert [b.tileNames for b in config._staged['batches']] == [['tile0', 'tile2'], ['tile1', 'tile3']]
    assert len(config._templates) == 2
    norms = {}
    nObj = 0
    for tileName, md in tileMaps.items():
        phot = None
        for f in pipelines._orderedFilters(config.parDict):
            ref, refObj = filters_ref.filterMaps(md, _copyFilter(f), tileName, returnFilter=True)
            norms[(tileName, f['label'])] = refObj.signalNorm
            if f['label'] == parDict['photFilter']:
                phot = {'SNMap': ref['SNMap'], 'data': ref['data']}
            refCat = photometry_ref.findObjects(ref, **DETECT)
            photometry_ref.measureFluxes(refCat, ref, None, photFilteredMapDict=phot)
            _compareCatalogs(config.catalogDict[f['label'] + "#" + tileName]['catalog'], refCat, ['SNR', 'y_c', 'fixed_SNR', 'fixed_y_c'])
            nObj += len(refCat)
    # the two geometries really differ: the normalisation moves by far more than the comparison tolerance
    assert abs(norms[('tile1', 'Arnaud_M2e14_z0p4')] / norms[('tile0', 'Arnaud_M2e14_z0p4')] - 1) > 0.02
    # a batch that mixes geometries is refused instead of silently taking tile 0's
    with pytest.raises(Exception):
        pipelines.TileBatch(pipelines.device.getOps(None), ['tile0', 'tile1'], tileMaps)
    report("two_declinations", objects=nObj)


def test_graph_slots_do_not_leak_between_batches(report):
    """Seven batches through three pipeline slots with CUDA graphs: the noise tables of every tile equal those of a run
    without graphs (a slot's graph outputs are overwritten by the batch that reuses it, so they are copied out per batch)."""
    from nemo_b200 import pipelines
    tileMaps = {}
    for t in range(7):
        md, _ = synth.makeTile(shape=(128, 160), freqsGHz=(148.0,), noiseUK=(40.0 + 7 * t,), numObjects=3, kind='sources', seed=400 + t, margin=30)
        tileMaps['tile%d' % t] = md
    parDict = {'mapFilters': [_beam('Beam', grid=10., numBins=2)], 'photFilter': None, 'thresholdSigma': 4.0, 'minObjPix': 1,
               'findCenterOfMass': True, 'useInterpolator': True, 'rejectBorder': 0}
    config = pipelines.SimpleConfig(parDict, tileMaps)
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=1)
    withGraphs = {k: np.array(v) for k, v in config.noiseTables.items()}
    ops = pipelines.device.getOps(None)
    batches = config._staged['batches']
    templates = [pipelines.templatesFor(config, ops, pipelines._orderedFilters(config.parDict), b) for b in batches]
    for b in batches:
        b.dev = None
    results = pipelines.processBatches(ops, batches, pipelines._orderedFilters(config.parDict), templates, config.parDict, useGraphs=False)
    assert len(withGraphs) == 7
    for b, (_, extras) in zip(batches, results):
        key = 'Beam#' + b.tileNames[0]
        assert np.array_equal(withGraphs[key], np.asarray(extras['noiseTable'])[0]), key
    tabs = [withGraphs['Beam#tile%d' % t] for t in range(7)]
    assert all(not np.array_equal(tabs[0], t) for t in tabs[1:])
    report("graph_slots", batches=7)


def test_batched_driver_reads_solid_angle_from_beam_file(tmp_path, report):
    """outputUnits uK with beams given as ACT-style profile files that carry a 'solid angle = ... nsr' line and no
    solidAngle_nsr key in the map dictionary: the batched driver and the class API report the same fluxJy (filters.py:180-196,
    photometry.py:336-339)."""
    from nemo_b200 import pipelines, filters, photometry, signals
    md, _ = synth.makeTile(shape=(192, 192), freqsGHz=(148.0,), noiseUK=(80.0,), numObjects=5, kind='sources', seed=77, margin=40)
    beam = signals.gaussianBeam(1.4)
    path = str(tmp_path / "beam_profile.txt")
    with open(path, "w") as fh:
        fh.write("# solid angle = 188.1 nsr\n")
        for r, p in zip(beam.rDeg, beam.profile1d):
            fh.write("%.8e %.8e\n" % (r, p))
    for m in md:
        m.pop('solidAngle_nsr', None)
        m['beamFileName'] = path
        m['beam'] = path
    f = _beam('Beam', grid=20.)
    parDict = {'mapFilters': [f], 'photFilter': None, 'thresholdSigma': 4.0, 'minObjPix': 1, 'findCenterOfMass': True,
               'useInterpolator': True, 'rejectBorder': 0}
    config = pipelines.SimpleConfig(parDict, {'T0': md})
    pipelines.filterMapsAndMakeCatalogs(config, verbose=False)
    batched = config.catalogDict['Beam#T0']['catalog']
    got = filters.filterMaps(md, _copyFilter(f), 'T0', verbose=False)
    assert got['beamSolidAngle_nsr'] == 188.1
    cat = photometry.findObjects(got, **DETECT)
    photometry.measureFluxes(cat, got, None)
    assert len(cat) == len(batched) >= 3
    assert np.all(np.asarray(cat['fluxJy'], dtype=float) != 0)
    np.testing.assert_allclose(np.asarray(batched['fluxJy'], dtype=float), np.asarray(cat['fluxJy'], dtype=float), rtol=1e-9)
    np.testing.assert_allclose(np.asarray(batched['err_fluxJy'], dtype=float), np.asarray(cat['err_fluxJy'], dtype=float), rtol=1e-9)
    report("beam_file_solid_angle", objects=len(cat))


def test_batched_driver_with_background_subtraction(report):
    """bckSub in the batched driver (Gaussian high-pass of the filtered map and of the filtered template, filters.py:655-659,
    853-856), with CUDA graphs: catalogs against the per-tile oracle."""
    from nemo_b200 import pipelines
    tileMaps = {}
    for t in range(2):
        md, _ = synth.makeTile(shape=(256, 256), freqsGHz=(148.0,), noiseUK=(50.0,), numObjects=5, kind='clusters', cmb=True, seed=500 + t)
        tileMaps['tile%d' % t] = md
    f = _arnaud('Arnaud_M2e14_z0p4', 2e14, 0.4, grid=20., bckSub=True, bckSubScaleArcmin=6.0)
    parDict = {'mapFilters': [f], 'photFilter': None, 'thresholdSigma': 4.0, 'minObjPix': 1, 'findCenterOfMass': True,
               'useInterpolator': True, 'rejectBorder': 0}
    config = pipelines.SimpleConfig(parDict, tileMaps)
    for _ in range(2):          # the second call replays the captured graphs
        pipelines.filterMapsAndMakeCatalogs(config, verbose=False, tilesPerBatch=1)
    nObj = 0
    for tileName, md in tileMaps.items():
        ref = filters_ref.filterMaps(md, _copyFilter(f), tileName)
        refCat = photometry_ref.findObjects(ref, **DETECT)
        photometry_ref.measureFluxes(refCat, ref, None)
        _compareCatalogs(config.catalogDict[f['label'] + "#" + tileName]['catalog'], refCat, ['SNR', 'y_c', 'deltaT_c'])
        nObj += len(refCat)
    assert nObj >= 4
    report("batched_bcksub", objects=nObj)
