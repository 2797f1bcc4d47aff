"""
Runs the REFERENCE's own code (``/root/reference/nemo/filters.py`` and ``photometry.py``, unmodified, imported from where they
lie) on the seeded tiles of the golden cases and stores its outputs in ``ref_<case>.npz``.  These fixtures pin the oracle
(``tests/test_oracle.py::test_oracle_matches_reference_run``) and the CUDA path (``tests/test_golden_gpu.py``) against what the
reference computes -- not against a restatement of it.

    python tests/golden/make_reference_golden.py        # needs /root/reference; the fixtures travel, the reference does not

What is the reference's and what is not.  The reference cannot be imported as a package here: astLib, astropy, pixell, pyccl,
mahotas, mpi4py, matplotlib and its own compiled ``nemoCython`` are absent, and this fork assigns ``enmap = None``
(nemo/filters.py:22-27), so even with every dependency present its Fourier filter could not run.  The generator therefore

  * puts stand-in modules (``unittest.mock``) in ``sys.modules`` for the absent third-party packages, so that the reference's
    source files import unchanged;
  * hands the reference the things the hot path needs from those packages: ``enmap.fft / ifft / apod / apply_window``
    (``oracle/pixell_shim.py``), a CAR ``astWCS.WCS`` and ``astCoords.calcAngSepDeg`` (``oracle/wcs_stub.py``) -- the published
    behaviour of pixell / astLib, restated -- and the five CODATA constants ``signals.fSZ`` reads from ``astropy.constants``;
  * gives its filter classes the signal TEMPLATE maps as inputs (``makeSignalTemplateMap`` is overridden to return the arrays the
    oracle's painter makes): template painting needs pixell / nemoCython and is an input of the path.

Everything else that runs is the reference's own code: ``filters.filterMaps``, ``MapFilter.__init__ / makeRadiansMap /
makeNoiseMap``, ``MatchedFilter.buildAndApply / applyFilter`` (noise covariance, smoothing, the per-mode solve, normalisation,
S/N, edge trimming, masks), ``photometry.getObjectPositions / findObjects / measureFluxes`` and ``catalogs.catalogListToTab / makeName`` (with a plain column table
standing in for ``astropy.table``), with the real numpy / scipy underneath.
"""

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest import mock

import numpy as np
import scipy, scipy.ndimage, scipy.interpolate, scipy.stats, scipy.optimize, scipy.integrate, scipy.special, scipy.signal  # noqa: F401,E401

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF = '/root/reference/nemo'


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Any import of an absent third-party package resolves to a MagicMock module."""
    NAMES = ('astLib', 'astropy', 'pixell', 'pyccl', 'mahotas', 'mpi4py', 'pylab', 'matplotlib', 'colorcet', 'nemoCython', 'PIL',
             'healpy', 'numba', 'cmocean', 'shapely', 'requests', 'zCluster', 'IPython', 'cycler', 'pyximport')

    def find_spec(self, name, path, target=None):
        if name.split('.')[0] in self.NAMES or name == 'nemo.nemoCython':
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__, m.__path__, m.__spec__, m.__all__ = spec.name, [], spec, []
        return m

    def exec_module(self, module):
        pass


def loadReference():
    """(filters, photometry) modules of the reference, executed from their own source files."""
    from oracle import pixell_shim, wcs_stub
    sys.meta_path.insert(0, _StubFinder())
    pkg = types.ModuleType('nemo')
    pkg.__path__ = [REF]
    pkg.__version__ = 'reference'
    sys.modules['nemo'] = pkg
    refFilters = importlib.import_module('nemo.filters')
    refPhotometry = importlib.import_module('nemo.photometry')
    refMaps = importlib.import_module('nemo.maps')
    shim = types.SimpleNamespace(fft=pixell_shim.fft, ifft=pixell_shim.ifft, apod=pixell_shim.apod, apply_window=pixell_shim.apply_window)
    coords = types.SimpleNamespace(calcAngSepDeg=wcs_stub.calcAngSepDeg,
                                   convertCoords=lambda a, b, ra, dec, epoch: wcs_stub.j2000ToGalactic(ra, dec))
    from nemo_b200 import catalogs as hostTables        # a column table with astropy's add_column / row assignment (pure Python)
    tables = types.SimpleNamespace(Table=hostTables.Table, Column=hostTables.Column)
    refCatalogs = importlib.import_module('nemo.catalogs')
    for mod in (refFilters, refPhotometry, refMaps, refCatalogs):
        mod.enmap = shim
        mod.astCoords = coords
        mod.atpy = tables
    # astropy.constants (CODATA 2018), read by the reference's signals.fSZ (nemo/signals.py:364-369)
    refSignals = importlib.import_module('nemo.signals')
    val = lambda v, cgs=None: types.SimpleNamespace(value=v, cgs=types.SimpleNamespace(value=cgs))
    consts = types.SimpleNamespace(h=val(6.62607015e-34, 6.62607015e-27), k_B=val(1.380649e-23, 1.380649e-16), sigma_T=val(6.6524587321e-29),
                                   m_e=val(9.1093837015e-31), c=val(299792458.0, 29979245800.0), pc=val(3.0856775814913674e16),
                                   M_sun=val(1.988409870698051e30))
    refSignals.constants = consts
    refPhotometry.constants = consts          # photometry.deltaTToJyPerSr reads the cgs values (nemo/photometry.py:465-467)
    return refFilters, refPhotometry


class RefMapDict(dict):
    """What the reference's filter classes expect of maps.MapDict: a dict with copy() and preprocess() (FITS I/O there; the tiles
    are already in memory here)."""

    def copy(self):
        return RefMapDict(self)

    def preprocess(self, tileName='PRIMARY', diagnosticsDir=None):
        return None


def runCase(refFilters, refPhotometry, name, tmpDir):
    from oracle import filters_ref
    import copy
    import make_golden
    case = make_golden.CASES[name]
    mapDicts, _ = make_golden.runCase(name)          # the same seeded tile the oracle fixture was made from
    f = copy.deepcopy(case['filter'])
    # the oracle object only paints the templates (inputs of the path)
    oracleObj = getattr(filters_ref, f['class'])(f['label'], mapDicts, copy.deepcopy(f['params']), tileName='T0')

    class RefFilter(refFilters.MatchedFilter):
        def makeSignalTemplateMap(self, beamFileName, amplitude=None):
            return oracleObj.makeSignalTemplateMap(beamFileName, amplitude=amplitude)

    refFilters.RefFilter = RefFilter
    refMaps = []
    for m in mapDicts:
        d = RefMapDict(m)
        d['wcs'].AWCS = None            # MapFilter.__init__ keeps it for pixell; nothing on this path reads it
        refMaps.append(d)
    fRef = {'label': f['label'], 'class': 'RefFilter', 'params': copy.deepcopy(f['params'])}
    os.makedirs(os.path.join(tmpDir, 'T0'), exist_ok=True)
    out, obj = refFilters.filterMaps(refMaps, fRef, 'T0', diagnosticsDir=tmpDir, selFnDir=tmpDir, verbose=False,
                                     undoPixelWindow=True, returnFilter=True)
    ids, pos, numPix, seg = refPhotometry.getObjectPositions(out['SNMap'], 4.0, findCenterOfMass=True)
    # the reference's findObjects / measureFluxes on its own filtered maps (ring removal, spline S/N, names, fluxes)
    import make_golden as mg
    cat = refPhotometry.findObjects(out, threshold=mg.DETECT['threshold'], minObjPix=mg.DETECT['minObjPix'],
                                    rejectBorder=mg.DETECT['rejectBorder'], findCenterOfMass=mg.DETECT['findCenterOfMass'],
                                    removeRings=mg.DETECT['removeRings'], ringThresholdSigma=mg.DETECT['ringThresholdSigma'],
                                    useInterpolator=mg.DETECT['useInterpolator'], verbose=False)
    refPhotometry.measureFluxes(cat, out, tmpDir, useInterpolator=mg.DETECT['useInterpolator'])
    catalog = {}
    for k in cat.keys():
        col = np.asarray(cat[k])
        catalog['refcat_' + k] = col.astype(str) if col.dtype == object or col.dtype.kind in 'US' else col.astype(np.float64)
    # the other branches of the reference's makeNoiseMap (filters.py:366-407 'smart', :464-465 'percentile'), on the filtered map
    noise = {}
    for key, noiseParams in (('smart', {'method': 'dataMap', 'noiseGridArcmin': 'smart', 'numNoiseBins': 6}),
                             ('percentile', {'method': 'dataMap', 'noiseGridArcmin': 20., 'numNoiseBins': 2 if name.startswith('beam') else 1,
                                             'RMSEstimator': 'percentile'})):   # constant weights leave bins empty, and np.percentile([]) raises
        saved = obj.params['noiseParams']
        obj.params['noiseParams'] = noiseParams
        noise['noise_' + key] = np.asarray(obj.makeNoiseMap(np.asarray(out['SNMap']) * 3.0 + 0.0))
        obj.params['noiseParams'] = saved
    return {'data': np.asarray(out['data']), 'SNMap': np.asarray(out['SNMap']), 'surveyMask': np.asarray(out['surveyMask']),
            'signalNorm': np.float64(obj.signalNorm), 'filt': np.asarray(obj.filt),
            'objIDs': np.asarray(ids), 'objPositions': np.asarray(pos, dtype=np.float64).reshape(-1, 2), 'objNumPix': np.asarray(numPix),
            'segmentationMap': np.asarray(seg, dtype=np.int32),
            'fRelKeys': np.array(sorted(obj.fRelWeights.keys()), dtype=np.float64),
            'fRelValues': np.array([obj.fRelWeights[k] for k in sorted(obj.fRelWeights.keys())], dtype=np.float64), **noise, **catalog}


def miscInputs():
    """Seeded inputs of the stand-alone pieces (also read by the tests): a two-frequency cube and kernel for the real-space
    filter's applyFilter, two overlapping catalogs for the cross-scale merge."""
    from oracle.wcs_stub import SimpleCARWCS
    rng = np.random.default_rng(2024)
    shape = (96, 120)
    wcs = SimpleCARWCS(shape, 0.5 / 60.0, crval=(30.0, -20.0))
    cube = rng.standard_normal((2,) + shape) * np.array([60.0, 40.0])[:, None, None]
    yy, xx = np.mgrid[-4:5, -4:5]
    kern = np.stack([np.exp(-(xx ** 2 + yy ** 2) / (2 * s ** 2)) - 0.4 * np.exp(-(xx ** 2 + yy ** 2) / (2 * (2 * s) ** 2)) for s in (1.2, 1.8)])
    kern[1] *= -0.3
    n = 40
    ra = 30.0 + rng.uniform(-0.4, 0.4, n)
    dec = -20.0 + rng.uniform(-0.3, 0.3, n)
    snr = rng.uniform(4, 30, n)
    catA = {'RADeg': ra, 'decDeg': dec, 'SNR': snr}
    keep = rng.random(n) < 0.7                                   # the second scale re-detects most objects, slightly shifted
    catB = {'RADeg': ra[keep] + rng.normal(0, 0.002, keep.sum()), 'decDeg': dec[keep] + rng.normal(0, 0.002, keep.sum()),
            'SNR': snr[keep] * rng.uniform(0.6, 1.5, keep.sum())}
    extra = {'RADeg': 30.0 + rng.uniform(-0.4, 0.4, 6), 'decDeg': -20.0 + rng.uniform(-0.3, 0.3, 6), 'SNR': rng.uniform(4, 9, 6)}
    catB = {k: np.concatenate([catB[k], extra[k]]) for k in catB}
    return {'shape': shape, 'wcs': wcs, 'cube': cube, 'kern': kern, 'catA': catA, 'catB': catB,
            'applyRADeg': 30.0, 'applyDecDeg': -20.0, 'bckSubScaleArcmin': 3.0, 'signalNorm': 0.37}


PHOT_SECOND_SCALE = {'label': 'Arnaud_M8e14_z0p2', 'class': 'ArnaudModelMatchedFilter',
                     'params': {'M500MSun': 8e14, 'z': 0.2, 'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': 20.},
                                'outputUnits': 'yc', 'edgeTrimArcmin': 0.}}


def runPhotFilter(refFilters, refPhotometry, tmpDir):
    """The photFilter flow of the inner loop (nemo/pipelines.py:238-284) with the reference's own functions: the golden Arnaud
    tile filtered at its own scale (the reference filter) and at a second scale; the second scale's catalog gets fixed_SNR from
    the reference filter's S/N map (getSNRValues, prefix 'fixed_') and fixed_ fluxes (measureFluxes with photFilteredMapDict)."""
    from oracle import filters_ref
    import copy
    import make_golden
    mapDicts, _ = make_golden.runCase('arnaud_f148_128')
    outs = []
    for f in (make_golden.CASES['arnaud_f148_128']['filter'], PHOT_SECOND_SCALE):
        f = copy.deepcopy(f)
        oracleObj = getattr(filters_ref, f['class'])(f['label'], mapDicts, copy.deepcopy(f['params']), tileName='T0')

        class RefFilter(refFilters.MatchedFilter):
            def makeSignalTemplateMap(self, beamFileName, amplitude=None, _o=oracleObj):
                return _o.makeSignalTemplateMap(beamFileName, amplitude=amplitude)
        refFilters.RefFilter = RefFilter
        refMaps = []
        for m in mapDicts:
            d = RefMapDict(m)
            d['wcs'].AWCS = None
            refMaps.append(d)
        os.makedirs(os.path.join(tmpDir, 'T0'), exist_ok=True)
        outs.append(refFilters.filterMaps(refMaps, {'label': f['label'], 'class': 'RefFilter', 'params': copy.deepcopy(f['params'])}, 'T0',
                                          diagnosticsDir=tmpDir, selFnDir=tmpDir, verbose=False, undoPixelWindow=True))
    phot = {'SNMap': outs[0]['SNMap'], 'data': outs[0]['data']}
    mg = make_golden
    cat = refPhotometry.findObjects(outs[1], threshold=mg.DETECT['threshold'], minObjPix=mg.DETECT['minObjPix'],
                                    rejectBorder=mg.DETECT['rejectBorder'], findCenterOfMass=mg.DETECT['findCenterOfMass'],
                                    removeRings=mg.DETECT['removeRings'], ringThresholdSigma=mg.DETECT['ringThresholdSigma'],
                                    useInterpolator=mg.DETECT['useInterpolator'], verbose=False)
    refPhotometry.getSNRValues(cat, phot['SNMap'], outs[1]['wcs'], prefix='fixed_', useInterpolator=mg.DETECT['useInterpolator'])
    refPhotometry.measureFluxes(cat, outs[1], tmpDir, photFilteredMapDict=phot, useInterpolator=mg.DETECT['useInterpolator'])
    res = {}
    for k in cat.keys():
        col = np.asarray(cat[k])
        res['phot_' + k] = col.astype(str) if col.dtype == object or col.dtype.kind in 'US' else col.astype(np.float64)
    return res


def runMisc(refFilters):
    """Stand-alone pieces of the reference: RealSpaceMatchedFilter.applyFilter (maps.subtractBackground / smoothMap +
    ndimage.convolve, nemo/filters.py:1172-1218), catalogs.makeOptimalCatalog (nemo/catalogs.py:149-189) and
    maps.getPixelAreaArcmin2Map (nemo/maps.py:1461-1482)."""
    from nemo_b200 import catalogs as hostTables
    refCatalogs = importlib.import_module('nemo.catalogs')
    refMaps = importlib.import_module('nemo.maps')
    refCatalogs.atpy.vstack = hostTables.vstack
    inp = miscInputs()
    obj = object.__new__(refFilters.RealSpaceMatchedFilter)
    obj.params = {'bckSub': True}
    obj.wcs, obj.kern2d, obj.signalNorm = inp['wcs'], inp['kern'], inp['signalNorm']
    obj.applyRACentre, obj.applyDecCentre, obj.bckSubScaleArcmin = inp['applyRADeg'], inp['applyDecDeg'], inp['bckSubScaleArcmin']
    obj.unfilteredMapsDictList = [{'obsFreqGHz': 97.8}, {'obsFreqGHz': 149.6}]
    filtered = obj.applyFilter(inp['cube'].copy(), calcFRelWeights=True)
    obj.params = {'bckSub': False}
    filteredNoBck = obj.applyFilter(inp['cube'].copy())

    def table(cols):
        return hostTables.Table.fromColumns([(k, np.array(v)) for k, v in cols.items()])
    merged = refCatalogs.makeOptimalCatalog({'A#t': {'catalog': table(inp['catA'])}, 'B#t': {'catalog': table(inp['catB'])}})
    area = refMaps.getPixelAreaArcmin2Map(inp['shape'], inp['wcs'])
    return {'rs_filtered': filtered, 'rs_filteredNoBck': filteredNoBck,
            'rs_fRel': np.array([obj.fRelWeights[97.8], obj.fRelWeights[149.6]]),
            'opt_RADeg': np.asarray(merged['RADeg'], dtype=np.float64), 'opt_decDeg': np.asarray(merged['decDeg'], dtype=np.float64),
            'opt_SNR': np.asarray(merged['SNR'], dtype=np.float64), 'pixAreaArcmin2Rows': np.asarray(area)[:, 0].copy(),
            'pixAreaConstantAlongRows': np.array(np.array_equal(np.asarray(area)[:, :1] * np.ones((1, inp['shape'][1])), area))}


def paintInputs():
    """Seeded inputs of the model-image painter (also read by the tests): a 300 x 420 tile at declination -35, a Gaussian beam
    written to a two-column text file (the ACT beam-profile format BeamProfile reads, nemo/signals.py:107-127), and a catalog of
    60 sources of either sign, some outside the tile and some on its rim, with S/N values for the minSNR cut."""
    from oracle.wcs_stub import SimpleCARWCS
    from oracle import signals_ref
    rng = np.random.default_rng(4242)
    shape = (300, 420)
    wcs = SimpleCARWCS(shape, 0.5 / 60.0, crval=(52.0, -35.0))
    n = 60
    x = rng.uniform(-15, shape[1] + 15, n)
    y = rng.uniform(-15, shape[0] + 15, n)
    x[:4] = [0.3, shape[1] - 1.4, 200.5, 17.0]
    y[:4] = [150.2, 40.0, shape[0] - 1.2, 0.1]
    ra, dec = wcs.pix2wcs(x, y)
    beam = signals_ref.gaussianBeam(1.4)
    return {'shape': shape, 'wcs': wcs, 'beamRDeg': beam.rDeg, 'beamProfile': beam.profile1d,
            'catalog': {'RADeg': np.asarray(ra), 'decDeg': np.asarray(dec), 'deltaT_c': rng.uniform(-400, 900, n), 'SNR': rng.uniform(2, 40, n)},
            'validAreaSection': [30, 390, 20, 270], 'minSNR': 5.0}


def writeBeamFile(path, rDeg, profile1d):
    with open(path, 'w') as fh:
        for r, p in zip(rDeg, profile1d):
            fh.write("%.17g %.17g\n" % (r, p))


def runPaint(tmpDir):
    """maps.makeModelImage of the reference (nemo/maps.py:1734-1899), point-source branch: makeDegreesDistanceMap per source,
    signals.makeBeamModelSignalMap, sum in catalog order, enmap.apply_window; once over the whole tile and once with
    validAreaSection and a minSNR cut."""
    from nemo_b200 import catalogs as hostTables
    refMaps = importlib.import_module('nemo.maps')
    refSignals = importlib.import_module('nemo.signals')
    # BeamProfile also derives the beam transform with pixell's curvedsky (None in this fork, nemo/signals.py:125); the
    # point-source painter never reads it
    refSignals.curvedsky = types.SimpleNamespace(profile2harm=lambda prof, r: np.ones(2))
    inp = paintInputs()
    beamPath = os.path.join(tmpDir, 'beam_f150.txt')
    writeBeamFile(beamPath, inp['beamRDeg'], inp['beamProfile'])
    tab = hostTables.Table.fromColumns([(k, np.array(v)) for k, v in inp['catalog'].items()])

    class AstWCSLike(type(inp['wcs'])):
        """astWCS.WCS.wcs2pix returns a list of [x, y] pairs when given sequences (which is how catalogs.getCatalogWithinImage
        and the validAreaSection cut index its result); the stub returns the two coordinate arrays."""

        def wcs2pix(self, RADeg, decDeg):
            x, y = super().wcs2pix(RADeg, decDeg)
            return (x, y) if np.ndim(RADeg) == 0 else np.stack([np.asarray(x), np.asarray(y)], axis=1).tolist()
    wcs = AstWCSLike(inp['shape'], inp['wcs'].pixDeg, crval=inp['wcs'].crval)
    full = refMaps.makeModelImage(inp['shape'], wcs, tab, beamPath, applyPixelWindow=True)
    noWindow = refMaps.makeModelImage(inp['shape'], wcs, tab, beamPath, applyPixelWindow=False)
    cut = refMaps.makeModelImage(inp['shape'], wcs, tab, beamPath, applyPixelWindow=True,
                                 validAreaSection=inp['validAreaSection'], minSNR=inp['minSNR'])
    return {'paint_full': np.asarray(full), 'paint_noWindow': np.asarray(noWindow), 'paint_cut': np.asarray(cut)}


def runBigCase(refFilters, refPhotometry, name, tmpDir):
    """A full-size case (big_cases.BIG_CASES) through the reference's filterMaps / getObjectPositions / findObjects / getSNRValues
    / measureFluxes in its loop order (nemo/pipelines.py:179-286), stored in compact form."""
    from oracle import filters_ref
    import copy
    import big_cases
    case = big_cases.BIG_CASES[name]
    mapDicts = case['tile']()
    res, phot = {}, None
    for i, f in enumerate(case['filters']):
        f = copy.deepcopy(f)
        oracleObj = getattr(filters_ref, f['class'])(f['label'], mapDicts, copy.deepcopy(f['params']), tileName='T0')

        class RefFilter(refFilters.MatchedFilter):
            def makeSignalTemplateMap(self, beamFileName, amplitude=None, _o=oracleObj):
                return _o.makeSignalTemplateMap(beamFileName, amplitude=amplitude)
        refFilters.RefFilter = RefFilter
        refMaps = []
        for m in mapDicts:
            d = RefMapDict(m)
            d['wcs'].AWCS = None
            refMaps.append(d)
        os.makedirs(os.path.join(tmpDir, 'T0'), exist_ok=True)
        out, obj = refFilters.filterMaps(refMaps, {'label': f['label'], 'class': 'RefFilter', 'params': copy.deepcopy(f['params'])}, 'T0',
                                         diagnosticsDir=tmpDir, selFnDir=tmpDir, verbose=False, undoPixelWindow=True, returnFilter=True)
        if f['label'] == case['photFilter']:
            phot = {'SNMap': out['SNMap'], 'data': out['data']}
        D = big_cases.DETECT
        seg = refPhotometry.getObjectPositions(out['SNMap'], D['threshold'], findCenterOfMass=True)[3]
        cat = refPhotometry.findObjects(out, threshold=D['threshold'], minObjPix=D['minObjPix'], rejectBorder=D['rejectBorder'],
                                        findCenterOfMass=D['findCenterOfMass'], removeRings=D['removeRings'],
                                        ringThresholdSigma=D['ringThresholdSigma'], useInterpolator=D['useInterpolator'], verbose=False)
        if phot is not None:
            refPhotometry.getSNRValues(cat, phot['SNMap'], out['wcs'], prefix='fixed_', useInterpolator=D['useInterpolator'])
        refPhotometry.measureFluxes(cat, out, tmpDir, photFilteredMapDict=phot, useInterpolator=D['useInterpolator'])
        res.update(big_cases.compact('s%d_' % i, out['data'], out['SNMap'], out['surveyMask'], seg, obj.signalNorm))
        res.update(big_cases.catalogColumns('s%d_cat_' % i, cat))
        if obj.fRelWeights:
            res['s%d_fRel' % i] = np.array([obj.fRelWeights[k] for k in sorted(obj.fRelWeights.keys())], dtype=np.float64)
    return res


def runShapes(refPhotometry):
    """The reference's findObjects with measureShapes = True (nemo/photometry.py:126-178) on the seeded S/N map of shape_case."""
    import shape_case
    import big_cases
    tile = shape_case.shapeTile()
    with np.errstate(divide='ignore', invalid='ignore'):
        cat = refPhotometry.findObjects(tile, verbose=False, **shape_case.DETECT)
    return big_cases.catalogColumns('shape_', cat)


def main():
    import tempfile
    sys.path.insert(0, HERE)
    import make_golden
    refFilters, refPhotometry = loadReference()
    shapes = runShapes(refPhotometry)
    np.savez_compressed(os.path.join(HERE, 'ref_shapes.npz'), **shapes)
    print("shapes: findObjects(measureShapes=True) stored (%d rows, %d with an ellipse)" % (len(shapes['shape_SNR']), int((shapes['shape_ellipse_A'] > 0).sum())))
    if '--only-shapes' in sys.argv:
        return
    tmpDir = tempfile.mkdtemp()
    paint = runPaint(tmpDir)
    np.savez_compressed(os.path.join(HERE, 'ref_paint.npz'), **paint)
    print("paint: maps.makeModelImage stored (%d non-zero pixels)" % int((paint['paint_full'] != 0).sum()))
    if '--only-paint' in sys.argv:
        return
    for name in make_golden.CASES:
        res = runCase(refFilters, refPhotometry, name, tmpDir)
        # compact: S/N and data as float64 (they are the contract), the filter as a checksum plus a few samples
        filt = res.pop('filt')
        res['filtSample'] = filt[:, ::16, ::16].copy()
        np.savez_compressed(os.path.join(HERE, 'ref_%s.npz' % name), **res)
        print("%s: reference run stored (%d labels above S/N 4, signalNorm %.6g)" % (name, len(res['objIDs']) - 1, float(res['signalNorm'])))
    import big_cases
    for name in big_cases.BIG_CASES:
        res = runBigCase(refFilters, refPhotometry, name, tmpDir)
        path = os.path.join(HERE, 'ref_big_%s.npz' % name)
        np.savez_compressed(path, **res)
        print("%s: reference run stored (%d rows in the first catalog, %.0f kB)" % (name, len(res['s0_cat_RADeg']), os.path.getsize(path) / 1e3))
    misc = runMisc(refFilters)
    misc.update(runPhotFilter(refFilters, refPhotometry, tmpDir))
    np.savez_compressed(os.path.join(HERE, 'ref_misc.npz'), **misc)
    print("misc: real-space applyFilter, makeOptimalCatalog (%d rows kept), pixel areas stored" % len(misc['opt_SNR']))


if __name__ == '__main__':
    main()
