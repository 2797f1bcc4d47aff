"""
signals.fitQ (reference nemo/signals.py:864-1125) on the device against the oracle's restatement of its per-tile body:
model clusters of several angular sizes filtered with the tile's reference filter, peak of the filtered signal over y0.
"""

import os

import numpy as np
import pytest

from oracle import synth, filters_ref, signals_ref

pytestmark = pytest.mark.gpu


def _filter(**extra):
    p = {'M500MSun': 2e14, 'z': 0.4, 'noiseParams': {'method': 'dataMap', 'noiseGridArcmin': 20.}, 'outputUnits': 'yc', 'edgeTrimArcmin': 0.}
    p.update(extra)
    return {'label': 'Arnaud_M2e14_z0p4', 'class': 'ArnaudModelMatchedFilter', 'params': p}


def _scales():
    """The reference scale first, then every sixth size of the reference's grid (0.1' ... 50'): the profile integrals of all 55
    take minutes on the host, in the oracle and in the product alike."""
    zRange, MRange = signals_ref.fitQScales(2e14, 0.4)
    keep = [0] + list(range(1, len(zRange), 6))
    return [zRange[i] for i in keep], [MRange[i] for i in keep]


def test_scale_grid_follows_the_reference():
    from nemo_b200 import signals
    zRange, MRange = signals.fitQScales(2e14, 0.4)
    zRef, MRef = signals_ref.fitQScales(2e14, 0.4)
    assert len(zRange) == 55 and zRange == zRef and np.allclose(MRange, MRef, rtol=1e-12)
    assert zRange[0] == 0.4 and MRange[0] == 2e14 and zRange[1:11] == [2.0] * 10 and zRange[-4:] == [0.07] * 4
    theta = [signals.calcTheta500Arcmin(z, M) for z, M in zip(zRange[1:], MRange[1:])]
    assert np.allclose(theta, np.power(10, np.arange(np.log10(0.1), np.log10(50), 0.05055349)), rtol=1e-9)


def test_fitQ_matches_oracle(tmp_path, report):
    from nemo_b200 import pipelines, signals, fitsio
    tileMaps, refs = {}, {}
    scales = _scales()
    f = _filter()
    for t in range(2):
        md, _ = synth.makeTile(shape=(256, 256), freqsGHz=(97.8, 149.6), noiseUK=(60.0, 40.0), beamFWHMArcmin=(2.2, 1.4), numObjects=4,
                               kind='clusters', cmb=True, seed=610 + t)
        tileMaps['tile%d' % t] = md
        _, obj = filters_ref.filterMaps(md, {'label': f['label'], 'class': f['class'], 'params': dict(f['params'])}, 'tile%d' % t, returnFilter=True)
        refs['tile%d' % t] = signals_ref.fitQTile(obj, md, scales[0], scales[1])
    parDict = {'mapFilters': [f, _filter(M500MSun=6e14, z=0.3)], 'photFilter': f['label'], 'thresholdSigma': 4.0, 'minObjPix': 1}
    parDict['mapFilters'][1]['label'] = 'Arnaud_M6e14_z0p3'
    config = pipelines.SimpleConfig(parDict, tileMaps)
    config.selFnDir = str(tmp_path)
    got = signals.fitQ(config, scales=scales)
    for name, (Q, theta, z) in refs.items():
        tab = got[name]
        assert len(tab) == len(Q) == len(scales[0])
        assert np.allclose(np.asarray(tab['theta500Arcmin']), theta, rtol=1e-12) and np.array_equal(np.asarray(tab['z']), z)
        err = float(np.abs(np.asarray(tab['Q']) / Q - 1).max())
        report("fitQ_oracle", tile=name, max_rel_dQ=err, Qmin=float(Q.min()), Qmax=float(Q.max()))
        assert err < 1e-8
        assert abs(np.asarray(tab['Q'])[np.argmin(np.abs(theta - signals.calcTheta500Arcmin(0.4, 2e14)))] - 1) < 1e-6     # Q = 1 at the reference scale
        cols, header = fitsio.readTable(os.path.join(str(tmp_path), 'QFit.fits'), name)
        assert np.array_equal(cols['Q'], np.asarray(tab['Q'])) and header['ZDEPQ'] == 0 and header['TILENAME'] == name
    assert not np.allclose(np.asarray(got['tile0']['Q']), np.asarray(got['tile1']['Q']), rtol=1e-6)      # Q follows each tile's own noise
