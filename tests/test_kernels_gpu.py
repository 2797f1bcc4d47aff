"""
Kernel-level parity: every C-ABI entry point against the numpy / scipy call it replaces, on seeded inputs.
Integer and index results (labels, argmax, masks) must be bit-exact; fp64 results must agree to the tolerance
written beside each check.  Runs on the GPU box:  pytest tests -m gpu
"""

import numpy as np
import pytest
from scipy import ndimage, interpolate

from oracle import pixell_shim as enmap

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops64():
    from nemo_b200.ops import Ops
    return Ops('f64')


@pytest.fixture(scope="module")
def ops32():
    from nemo_b200.ops import Ops
    return Ops('f32')


def _np(t):
    return t.detach().cpu().numpy()


def _relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(64, 64), (128, 256), (256, 128), (512, 512), (1024, 1024), (2048, 512), (64, 4096), (4096, 64)])
@pytest.mark.parametrize("apod", [0, 20])
def test_rfft2_f64(ops64, report, shape, apod):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2,) + shape)
    ref = np.fft.rfft2(enmap.apod(x, apod) if apod else x) / np.sqrt(shape[0] * shape[1])
    got = _np(ops64.rfft2(ops64.to_device(x), apodPix=apod))
    err = _relerr(got, ref)
    report("rfft2_f64", shape=shape, apod=apod, relerr=err)
    assert err < 1e-13


@pytest.mark.parametrize("shape", [(64, 64), (256, 512), (1024, 1024)])
def test_rfft2_f32(ops32, report, shape):
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2,) + shape).astype(np.float32)
    ref = np.fft.rfft2(enmap.apod(x.astype(np.float64), 20)) / np.sqrt(shape[0] * shape[1])
    got = _np(ops32.rfft2(ops32.to_device(x), apodPix=20))
    err = _relerr(got, ref)
    report("rfft2_f32", shape=shape, relerr=err)
    assert err < 2e-6


@pytest.mark.parametrize("shape", [(64, 64), (128, 512), (1024, 1024), (512, 2048)])
def test_irfft2_f64(ops64, report, shape):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3,) + shape)
    spec = np.fft.rfft2(x)
    mask = (rng.random((3,) + shape) > 0.2).astype(np.uint8)
    norm = rng.random(3) + 0.5
    got = _np(ops64.irfft2(ops64.to_device(spec, ops64.cdtype), shape[1], scale=0.5, norm=ops64.to_device(norm),
                           mask=ops64.to_device(mask, ops64.torch.uint8)))
    ref = x * shape[0] * shape[1] * 0.5 * norm[:, None, None] * mask
    err = _relerr(got, ref)
    report("irfft2_f64", shape=shape, relerr=err)
    assert err < 1e-13


@pytest.mark.parametrize("F,S,shape", [(1, 1, (128, 128)), (2, 3, (256, 128)), (3, 2, (64, 256)), (2, 1, (1024, 1024))])
def test_apply_filter_f64(ops64, report, F, S, shape):
    rng = np.random.default_rng(4)
    nT = 2
    H, W = shape
    maps = rng.standard_normal((nT, F, H, W))
    filt = rng.random((nT * S, F, H, W))            # full plane, NOT symmetric under k -> -k
    fm = np.fft.fft2(maps) / np.sqrt(H * W)
    norm = rng.random(nT * S) + 0.5
    mask = (rng.random((nT, H, W)) > 0.1).astype(np.uint8)
    ref = np.empty((nT * S, H, W))
    for o in range(nT * S):
        ref[o] = np.real(np.fft.ifft2(fm[o // S] * filt[o]) * H * W).sum(axis=0) * norm[o] * mask[o // S]
    fmaps = ops64.rfft2(ops64.to_device(maps.reshape(nT * F, H, W))).reshape(nT, F, H, W // 2 + 1)
    filtH = ops64.fold_filter(ops64.to_device(filt))
    got = _np(ops64.apply_filter(fmaps, filtH, S, norm=ops64.to_device(norm), mask=ops64.to_device(mask, ops64.torch.uint8)))
    err = _relerr(got, ref)
    report("apply_filter_f64", F=F, S=S, shape=shape, relerr=err)
    assert err < 1e-12


@pytest.mark.parametrize("shape", [(128, 128), (256, 512), (1024, 1024)])
def test_window_deconv_f64(ops64, report, shape):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2,) + shape)
    x[:, :10, :] = 0
    x[:, 50:60, 30:40] = 0
    ref = enmap.apply_window(x, pow=-1.0)
    ref[x == 0] = 0
    got = _np(ops64.apply_window(ops64.to_device(x), -1.0, restoreZeros=True))
    err = _relerr(got, ref)
    assert np.array_equal(got == 0, x == 0)
    ref2 = enmap.apply_window(x, pow=1.0)
    got2 = _np(ops64.apply_window(ops64.to_device(x), 1.0, restoreZeros=False))
    err2 = _relerr(got2, ref2)
    report("window_f64", shape=shape, relerr=err, relerr_pow1=err2)
    assert err < 1e-12 and err2 < 1e-12


# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F,shape", [(1, (64, 128)), (2, (128, 64)), (3, (64, 64)), (2, (512, 512)), (2, (97, 131)), (1, (300, 421))])
def test_xpower_smooth_f64(ops64, report, F, shape):
    rng = np.random.default_rng(6)
    H, W = shape
    maps = rng.standard_normal((2, F, H, W)) * np.array([1.0, 3.0, 0.5])[:F][None, :, None, None]
    maps[:, 1:] += maps[:, :1] * 5  # correlated frequencies
    fm = np.fft.fft2(maps) / np.sqrt(H * W)
    fmaps = ops64.rfft2(ops64.to_device(maps.reshape(2 * F, H, W))).reshape(2, F, H, W // 2 + 1)
    cov = _np(ops64.xpower_smooth(fmaps, 3.0, W=W))
    p = 0
    worst = 0.0
    for i in range(F):
        for j in range(i, F):
            for b in range(2):
                ref = ndimage.gaussian_filter(np.real(fm[b, i] * fm[b, j].conj()), (3, 3))
                worst = max(worst, _relerr(cov[b, p], ref))
            p += 1
    report("xpower_smooth_f64", F=F, shape=shape, relerr=worst)
    assert worst < 1e-12


@pytest.mark.parametrize("F,S,shape", [(1, 2, (64, 64)), (2, 2, (128, 64)), (3, 1, (64, 128)), (2, 2, (97, 131))])
def test_build_filter_f64(ops64, report, F, S, shape):
    rng = np.random.default_rng(7)
    H, W = shape
    nT = 2
    maps = rng.standard_normal((nT, F, H, W))
    maps[:, 1:] += maps[:, :1] * 3
    fm = np.fft.fft2(maps) / np.sqrt(H * W)
    fmaps = ops64.rfft2(ops64.to_device(maps.reshape(nT * F, H, W))).reshape(nT, F, H, W // 2 + 1)
    covT = ops64.xpower_smooth(fmaps, 3.0, W=W)
    templ = rng.standard_normal((S, F, H, W))
    fsigT = ops64.rfft2(ops64.to_device(templ.reshape(S * F, H, W))).reshape(S, F, H, W // 2 + 1)
    w = rng.standard_normal(F)
    filtH, filtFull = ops64.build_filter(covT, fsigT, w, S, sigPerTile=False, wantFull=True)
    filtFull = _np(filtFull)
    fsig = np.fft.fft2(templ) / np.sqrt(H * W)
    worst = 0.0
    for t in range(nT):
        noiseCov = np.array([[ndimage.gaussian_filter(np.real(fm[t, i] * fm[t, j].conj()), (3, 3)) for j in range(F)] for i in range(F)])
        cov = np.ascontiguousarray(np.moveaxis(noiseCov, (0, 1), (2, 3)))
        for s in range(S):
            rhs = np.moveaxis(w[:, None, None] * abs(fsig[s]), 0, 2)[..., None]
            ref = np.moveaxis((np.linalg.inv(cov) @ rhs)[..., 0], 2, 0)
            worst = max(worst, _relerr(filtFull[t * S + s], ref))
    refH = 0.5 * (filtFull + np.roll(np.flip(filtFull, axis=(-2, -1)), (1, 1), axis=(-2, -1)))[..., :W // 2 + 1]
    errH = _relerr(_np(filtH), refH)
    report("build_filter_f64", F=F, S=S, shape=shape, relerr=worst, relerr_fold=errH)
    assert worst < 1e-9 and errH < 1e-15


@pytest.mark.parametrize("F,S,shape", [(1, 2, (64, 64)), (2, 2, (128, 64)), (3, 1, (64, 128)), (2, 3, (300, 420)), (4, 1, (96, 100)), (1, 1, (1024, 1024))])
def test_build_filter_fused_matches_two_step(ops64, report, F, S, shape):
    """nb_build_filter_fused (covariance kept in shared memory) against nb_xpower_smooth + nb_build_filter, both the folded
    half-plane filter and the full plane, for tiles that do and do not divide into 32 x 32 blocks."""
    rng = np.random.default_rng(23)
    H, W = shape
    nT = 2
    maps = rng.standard_normal((nT, F, H, W))
    maps[:, 1:] += maps[:, :1] * 3
    fmaps = ops64.rfft2(ops64.to_device(maps.reshape(nT * F, H, W))).reshape(nT, F, H, W // 2 + 1)
    templ = rng.standard_normal((S, F, H, W))
    fsig = ops64.rfft2(ops64.to_device(templ.reshape(S * F, H, W))).reshape(S, F, H, W // 2 + 1)
    w = rng.standard_normal(F)
    cov = ops64.xpower_smooth(fmaps, 3.0)
    refH, refFull = ops64.build_filter(cov, fsig, w, S, sigPerTile=False, wantFull=True)
    gotH, gotFull = ops64.build_filter_fused(fmaps, fsig, w, S, False, sigma=3.0, wantFull=True)
    errH = _relerr(_np(gotH), _np(refH))
    errFull = _relerr(_np(gotFull), _np(refFull))
    report("build_filter_fused", F=F, S=S, shape=shape, relerr_half=errH, relerr_full=errFull)
    assert errH < 1e-11 and errFull < 1e-11


def test_build_filter_singular(ops64):
    H, W, F = 64, 64, 2
    cov = np.zeros((1, 3, H, W))
    cov[0, 0] = 2.0
    cov[0, 1] = 1.0
    cov[0, 2] = 0.5        # [[2, 1], [1, 0.5]] is exactly singular
    cov[0, :, 5, 7] = [2.0, 0.0, 4.0]
    templ = np.zeros((1, F, H, W))
    templ[0, :, 0, 0] = 1.0
    fsig = ops64.rfft2(ops64.to_device(templ.reshape(F, H, W))).reshape(1, F, H, W // 2 + 1)
    filtH, filtFull = ops64.build_filter(ops64.to_device(cov), fsig, [1.0, 1.0], 1, False, wantFull=True)
    ff = _np(filtFull)[0]
    assert ff[0, 0, 0] == 0.0 and ff[1, 3, 3] == 0.0
    np.testing.assert_allclose(ff[:, 5, 7], np.array([0.5, 0.25]) / np.sqrt(H * W), rtol=1e-14)


@pytest.mark.parametrize("mode", [0, 1])
def test_template_norm_f64(ops64, report, mode):
    rng = np.random.default_rng(8)
    H, W = 256, 512
    m = ndimage.gaussian_filter(rng.standard_normal((3, H, W)), 2.0)
    m[1, 100, 200] = m[1].max() + 1.0
    m[1, 150, 10] = m[1, 100, 200]      # tie: first in raster order wins
    cy, cx = H / 2.0 - 0.3, W / 2.0 + 0.45
    norm, argmax, maxVal, centre = ops64.template_norm(ops64.to_device(m), mode, cy, cx, 2e-4)
    norm, argmax, maxVal, centre = _np(norm), _np(argmax), _np(maxVal), _np(centre)
    for b in range(3):
        assert argmax[b] == np.argmax(m[b])
        assert maxVal[b] == m[b].max()
        if mode == 1:
            sp = interpolate.RectBivariateSpline(np.arange(H), np.arange(W), m[b], kx=3, ky=3)(cy, cx)[0][0]
            assert abs(centre[b] - sp) < 1e-12 * max(1, abs(sp))
            assert abs(norm[b] - 2e-4 / sp) <= 1e-10 * abs(2e-4 / sp)
        else:
            assert norm[b] == 2e-4 / m[b].max()
    report("template_norm", mode=mode, ok=1)


@pytest.mark.parametrize("H,W", [(64, 128), (97, 131)])
def test_plane_values_at_f64(ops64, report, H, W):
    rng = np.random.default_rng(9)
    B, F = 2, 2
    maps = rng.standard_normal((B, F, H, W))
    filt = rng.random((B, F, H, W))
    fm = np.fft.fft2(maps) / np.sqrt(H * W)
    cube = np.real(np.fft.ifft2(fm * filt) * H * W)
    idx = np.array([np.argmax(cube[b].sum(axis=0)) for b in range(B)])
    ref = np.array([[cube[b, f].flatten()[idx[b]] for f in range(F)] for b in range(B)])
    fmaps = ops64.rfft2(ops64.to_device(maps.reshape(B * F, H, W))).reshape(B, F, H, W // 2 + 1)
    filtH = ops64.fold_filter(ops64.to_device(filt))
    got = _np(ops64.plane_values_at(fmaps, filtH, ops64.to_device(idx, ops64.torch.int64), W=W))
    err = _relerr(got, ref)
    # with the total given, the last plane is the remainder: the same numbers to rounding
    total = ops64.to_device(ref.sum(axis=1))
    got2 = _np(ops64.plane_values_at(fmaps, filtH, ops64.to_device(idx, ops64.torch.int64), W=W, total=total))
    err2 = _relerr(got2, ref)
    report("plane_values_at", relerr=err, relerr_with_total=err2)
    assert err < 1e-11 and err2 < 1e-11


# ----------------------------------------------------------------------------------------------------------
def _oracle_noise(mapData, weightsList, pixDeg, gridArcmin, numBins, estimator=None):
    from oracle import filters_ref
    from oracle.wcs_stub import SimpleCARWCS
    wcs = SimpleCARWCS(mapData.shape, pixDeg)
    md = [{'data': mapData, 'weights': w, 'wcs': wcs, 'flagMask': np.zeros(mapData.shape, dtype=np.uint8), 'obsFreqGHz': 150.0,
           'label': 'x', 'surveyMask': None, 'pointSourceMask': None} for w in weightsList]
    params = {'noiseParams': {'noiseGridArcmin': gridArcmin}}
    if numBins is not None:
        params['noiseParams']['numNoiseBins'] = numBins
    if estimator is not None:
        params['noiseParams']['RMSEstimator'] = estimator
    f = filters_ref.MapFilter('t', md, params)
    return f.makeNoiseMap(mapData)


def _chunks(n, gridSize):
    return np.linspace(0, n, int(n / gridSize + 1), dtype=int)


def _noise_case(shape, seed, constantWeights=False):
    rng = np.random.default_rng(seed)
    H, W = shape
    sig = 1.0 + 2.0 * (np.arange(W)[None, :] > W // 2) + 0.2 * rng.random((H, W))
    m = rng.standard_normal((H, W)) * sig
    m[rng.random((H, W)) < 0.002] += 30.0           # outliers for the clip
    m[:12, :] = 0
    m[:, -9:] = 0
    m[100:130, 50:200] = 0                          # a hole covering whole rows of some cells
    w1 = 1.0 / sig ** 2
    if constantWeights:
        w1 = np.full((H, W), 0.25)
    w2 = w1 * (0.5 + rng.random((H, W)))
    w1[m == 0] = 0
    return m, [w1, w2] if not constantWeights else [w1]


def _run_cell_rms(ops, m, weights, gridArcmin, numBins):
    H, W = m.shape
    pixDeg = 0.5 / 60
    wm = ops.median_weights(ops.to_device(np.array(weights)[None]))
    if gridArcmin is None:
        yc, xc, ov = np.array([0, H]), np.array([0, W]), 0
    else:
        gridSize = int(round((gridArcmin / 60.) / pixDeg))
        yc, xc, ov = _chunks(H, gridSize), _chunks(W, gridSize), gridSize // 2
    rms, table, edges = ops.cell_rms(ops.to_device(m[None]), wm, 1, yc, xc, ov, numBins)
    return _np(rms)[0], _np(wm)[0]


@pytest.mark.parametrize("numBins", [1, 4, 8])
@pytest.mark.parametrize("shape,gridArcmin", [((256, 256), 20.0), ((300, 420), 20.0), ((512, 640), 40.0), ((200, 256), None)])
def test_cell_rms_f64(ops64, report, numBins, shape, gridArcmin):
    """Cells of 80 x 80 (registers only), 160 x 160 (registers + shared memory) and a single whole-map cell (streaming
    kernel), against the oracle's makeNoiseMap."""
    m, weights = _noise_case(shape, 10 + numBins)
    ref = _oracle_noise(m, weights, 0.5 / 60, gridArcmin, numBins if numBins > 1 else None)
    got, wm = _run_cell_rms(ops64, m, weights, gridArcmin, numBins)
    assert np.array_equal(wm, np.median(np.array(weights), axis=0))
    err = _relerr(got, ref)
    report("cell_rms_f64", numBins=numBins, shape=shape, grid=gridArcmin, relerr=err, zeros_equal=int(np.array_equal(got == 0, ref == 0)))
    assert np.array_equal(got == 0, ref == 0)
    assert err < 1e-12


@pytest.mark.parametrize("case", ["offset_mean", "two_noise_levels", "heavy_tails", "bright_sources", "mostly_masked"])
def test_cell_rms_one_bin_hard_cases(ops64, report, case):
    """One weight bin per cell (zone moments + list-based clipping passes, noise_onebin.cu) on maps that leave the easy regime: a
    mean far from zero (pivoted sums), cells straddling a 20x jump of the noise level (the list of values above t_lo overflows:
    sweeping passes), heavy tails and bright sources (thresholds that fall below t_lo: the cell is swept again), and a map that
    is mostly zero.  Same answer as the oracle's makeNoiseMap in every case."""
    rng = np.random.default_rng(500 + ["offset_mean", "two_noise_levels", "heavy_tails", "bright_sources", "mostly_masked"].index(case))
    H, W = 420, 520
    m = rng.standard_normal((H, W))
    if case == "offset_mean":
        m = 50.0 + m
    elif case == "two_noise_levels":
        m[:, W // 2 + 17:] *= 20.0
        m[H // 2 + 5:, :] *= 3.0
    elif case == "heavy_tails":
        m = rng.standard_t(1.5, size=(H, W))
    elif case == "bright_sources":
        yy, xx = np.mgrid[0:H, 0:W]
        for _ in range(12):
            y0, x0 = rng.integers(0, H), rng.integers(0, W)
            m += rng.uniform(50, 3000) * np.exp(-((yy - y0) ** 2 + (xx - x0) ** 2) / (2 * 3.0 ** 2))
    elif case == "mostly_masked":
        m[rng.random((H, W)) < 0.97] = 0
        m[:, 100:300] = 0
    m[:7, :] = 0
    m[:, -5:] = 0
    w = 1.0 + rng.random((H, W))
    for gridArcmin in (20.0, 40.0):
        ref = _oracle_noise(m, [w], 0.5 / 60, gridArcmin, None)
        got, _ = _run_cell_rms(ops64, m, [w], gridArcmin, 1)
        err = _relerr(got, ref)
        report("cell_rms_one_bin_" + case, grid=gridArcmin, relerr=err)
        assert np.array_equal(got == 0, ref == 0)
        assert err < 1e-11


@pytest.mark.parametrize("numBins", [1, 8])
def test_cell_rms_constant_weights(ops64, report, numBins):
    """All weights equal: every percentile edge collapses onto the same value and only the last bin holds pixels."""
    m, weights = _noise_case((256, 320), 3, constantWeights=True)
    ref = _oracle_noise(m, weights, 0.5 / 60, 20.0, numBins if numBins > 1 else None)
    got, _ = _run_cell_rms(ops64, m, weights, 20.0, numBins)
    assert np.array_equal(got == 0, ref == 0)
    assert _relerr(got, ref) < 1e-12


@pytest.mark.parametrize("gridArcmin", [20.0, 40.0])
def test_cell_rms_clustered_weights(ops64, report, gridArcmin):
    """Weights packed into a narrow range plus a few far outliers: the percentile buckets of the first histogram
    level hold thousands of distinct keys, which takes the radix-select fallback inside the on-chip kernel."""
    m, _ = _noise_case((512, 640), 5)
    rng = np.random.default_rng(77)
    w = 1.0 + 1e-6 * rng.random(m.shape)
    w[rng.random(m.shape) < 0.01] = 1e6
    ref = _oracle_noise(m, [w], 0.5 / 60, gridArcmin, 8)
    got, _ = _run_cell_rms(ops64, m, [w], gridArcmin, 8)
    err = _relerr(got, ref)
    report("cell_rms_clustered", grid=gridArcmin, relerr=err)
    assert np.array_equal(got == 0, ref == 0)
    assert err < 1e-12


@pytest.mark.parametrize("numBins,shape,gridArcmin", [(1, (256, 256), 20.0), (4, (300, 420), 20.0), (8, (512, 640), 40.0)])
def test_cell_rms_percentile_estimator(ops64, report, numBins, shape, gridArcmin):
    """RMSEstimator: 'percentile' (filters.py:464-465): np.percentile(abs(binValues), 68.3) per cell and weight bin, exact
    order statistics by radix select, against the oracle."""
    m, weights = _noise_case(shape, 60 + numBins)
    ref = _oracle_noise(m, weights, 0.5 / 60, gridArcmin, numBins if numBins > 1 else None, estimator='percentile')
    H, W = m.shape
    gridSize = int(round((gridArcmin / 60.) / (0.5 / 60)))
    yc, xc, ov = _chunks(H, gridSize), _chunks(W, gridSize), gridSize // 2
    wm = ops64.median_weights(ops64.to_device(np.array(weights)[None]))
    rms, table, edges = ops64.cell_rms(ops64.to_device(m[None]), wm, 1, yc, xc, ov, numBins, estimator='percentile')
    got = _np(rms)[0]
    err = _relerr(got, ref)
    report("cell_rms_percentile", numBins=numBins, shape=shape, relerr=err)
    assert np.array_equal(got == 0, ref == 0)
    assert err < 1e-14


@pytest.mark.parametrize("numBins,shape,gridArcmin", [(1, (256, 256), 20.0), (4, (300, 420), 20.0), (8, (512, 640), 40.0), (64, (128, 128), 20.0)])
def test_cell_rms_biweight_estimator(ops64, report, numBins, shape, gridArcmin):
    """RMSEstimator: 'biweight' (filters.py:461-465): astropy's biweight_scale(binValues, c = 9, modify_sample_size = True) per
    cell and weight bin for bins of at least 10 pixels (the last case has bins below that: RMS 0, nothing painted), median and
    median absolute deviation by exact radix select, against the oracle's restatement of the published definition."""
    m, weights = _noise_case(shape, 80 + numBins)
    ref = _oracle_noise(m, weights, 0.5 / 60, gridArcmin, numBins if numBins > 1 else None, estimator='biweight')
    H, W = m.shape
    gridSize = int(round((gridArcmin / 60.) / (0.5 / 60)))
    yc, xc, ov = _chunks(H, gridSize), _chunks(W, gridSize), gridSize // 2
    wm = ops64.median_weights(ops64.to_device(np.array(weights)[None]))
    rms, table, edges = ops64.cell_rms(ops64.to_device(m[None]), wm, 1, yc, xc, ov, numBins, estimator='biweight')
    got = _np(rms)[0]
    err = _relerr(got, ref)
    report("cell_rms_biweight", numBins=numBins, shape=shape, relerr=err, painted=float((ref > 0).mean()))
    assert np.array_equal(got == 0, ref == 0)
    assert err < 1e-12


@pytest.mark.parametrize("numBins", [2, 6, 100])
@pytest.mark.parametrize("shape", [(256, 256), (300, 420)])
def test_smart_noise_bins_f64(ops64, report, numBins, shape):
    """noiseGridArcmin = 'smart' (filters.py:366-407): whole-map weight bins with np.linspace edges and strict inequalities,
    3-sigma clipped, against the oracle.  The case has zeroed regions, outliers and pixels whose weight sits exactly on the
    first and last edge (the extremes), which belong to no bin."""
    m, weights = _noise_case(shape, 40 + numBins)
    ref = _oracle_noise(m, weights, 0.5 / 60, "smart", numBins)
    wm = ops64.median_weights(ops64.to_device(np.array(weights)[None]))
    rms, table, edges = ops64.smart_rms(ops64.to_device(m[None]), wm, 1, numBins)
    got = _np(rms)[0]
    medW = np.median(np.array(weights), axis=0)
    assert np.array_equal(_np(edges)[0], np.linspace(medW.min(), medW.max(), numBins))
    err = _relerr(got, ref)
    report("smart_rms_f64", numBins=numBins, shape=shape, relerr=err, painted=float((ref > 0).mean()))
    assert np.array_equal(got == 0, ref == 0)
    assert (ref > 0).mean() > 0.5
    assert err < 1e-11
    # the table holds the values painted
    assert set(np.unique(got[got > 0])) <= set(_np(table)[0].tolist())


def test_smart_noise_constant_weights(ops64):
    """All weights equal: every edge coincides, no pixel lies strictly between two edges, the noise map is zero."""
    m, weights = _noise_case((256, 320), 3, constantWeights=True)
    weights = [np.full(m.shape, 0.25)]
    ref = _oracle_noise(m, weights, 0.5 / 60, "smart", 8)
    wm = ops64.median_weights(ops64.to_device(np.array(weights)[None]))
    rms, _, _ = ops64.smart_rms(ops64.to_device(m[None]), wm, 1, 8)
    assert not ref.any() and not _np(rms).any()


@pytest.mark.parametrize("numBins,gridArcmin", [(1, 20.0), (8, 40.0)])
def test_cell_rms_f32(ops32, report, numBins, gridArcmin):
    """fp32 storage: the oracle is fed the same fp32-rounded maps, so only the final rounding of the RMS map differs."""
    m, weights = _noise_case((512, 512), 21)
    m = m.astype(np.float32).astype(np.float64)
    weights = [w.astype(np.float32).astype(np.float64) for w in weights]
    # fp32 median of two values = their mean rounded to fp32; give the oracle that median as a single weight map
    wmed32 = ((weights[0] + weights[1]) * 0.5).astype(np.float32).astype(np.float64)
    ref = _oracle_noise(m, [wmed32], 0.5 / 60, gridArcmin, numBins if numBins > 1 else None)
    got, wm = _run_cell_rms(ops32, m, weights, gridArcmin, numBins)
    assert np.array_equal(wm.astype(np.float64), wmed32)
    err = _relerr(got, ref)
    report("cell_rms_f32", numBins=numBins, grid=gridArcmin, relerr=err)
    assert np.array_equal(got == 0, ref == 0)
    assert err < 1e-6


@pytest.mark.parametrize("trim,realSpace", [(0, False), (24, False), (31, True), (120, False), (300, False), (700, False)])
def test_snr_mask_f64(ops64, report, trim, realSpace):
    rng = np.random.default_rng(12)
    H, W = 256, 320
    m = rng.standard_normal((2, H, W))
    m[:, :15, :] = 0
    m[:, :, -7:] = 0
    m[0, 200, 100] = 0
    ps = np.ones((1, H, W), dtype=np.uint8)
    ps[0, 60:70, 60:75] = 0
    survey = (rng.random((1, H, W)) > 0.05).astype(np.uint8)
    rmsIn = np.abs(rng.standard_normal((2, H, W))) + 0.1
    rmsIn[:, 100:110, :] = 0
    m = m * ps
    apodPix = 20
    refs = []
    for b in range(2):
        fm = m[b]
        SN = np.zeros(fm.shape) + fm
        valid = rmsIn[b] > 0
        SN[valid] = SN[valid] / rmsIn[b][valid]
        if trim > 0:
            edge = np.array(ndimage.rank_filter(abs(fm + (1 - ps[0])), 0, size=(trim, trim)) > 0, dtype=float)
        else:
            edge = np.ones(fm.shape)
        d = fm * edge
        sm = edge * survey[0] * ps[0]
        if not realSpace:
            d = d * sm
        apodMask = np.equal(enmap.apod(np.ones(fm.shape), apodPix), 1)
        sm = sm * apodMask
        SN = SN * sm
        SN[np.isnan(SN)] = 0
        refs.append((d, SN, sm, rmsIn[b] * sm))
    t = ops64.torch
    data, sn, maskOut, rms = ops64.snr_mask(ops64.to_device(m), ops64.to_device(rmsIn), ops64.to_device(ps, t.uint8),
                                            ops64.to_device(survey, t.uint8), 2, trim, apodPix, realSpace)
    for b in range(2):
        assert np.array_equal(_np(maskOut)[b], refs[b][2].astype(np.uint8)), "mask differs (trim=%d)" % trim
        assert np.array_equal(_np(data)[b], refs[b][0])
        assert np.array_equal(_np(sn)[b], refs[b][1])
        assert np.array_equal(_np(rms)[b], refs[b][3])
    report("snr_mask_f64", trim=trim, realSpace=realSpace, ok=1)


@pytest.mark.parametrize("shape", [(300, 420), (128, 256)])
def test_zero_flags_of_apply_filter_feed_the_edge_filter(ops64, report, shape):
    """nb_apply_filter reports, per image, whether map + (1 - psMask) has an exactly zero pixel (what the rank filter of K7 looks
    for, filters.py:730); nb_snr_mask* skips the filter for images without one.  Three images: random (no zero: flag 0 on the
    compile-time-plan path), all-zero spectra (every pixel zero: flag 1), random with a point-source hole (the hole is filled back
    in, flag 0).  The outputs with the flags must equal those without, bit for bit; sides without a compile-time plan answer 1
    (unknown) for every image."""
    H, W = shape
    rng = np.random.default_rng(21)
    t = ops64.torch
    maps = rng.standard_normal((3, 1, H, W))
    maps[1] = 0
    ps = np.ones((3, H, W), dtype=np.uint8)
    ps[2, 40:50, 60:80] = 0
    survey = np.ones((3, H, W), dtype=np.uint8)
    fm = ops64.rfft2(ops64.to_device(maps.reshape(3, H, W))).reshape(3, 1, H, W // 2 + 1)
    filt = ops64.fold_filter(ops64.to_device(rng.random((3, 1, H, W)) + 0.5))
    psT, svT = ops64.to_device(ps, t.uint8), ops64.to_device(survey, t.uint8)
    out, flags = ops64.apply_filter(fm, filt, 1, mask=psT, W=W, zeroFlags=True)
    ref = ops64.apply_filter(fm, filt, 1, mask=psT, W=W)
    assert np.array_equal(_np(out), _np(ref))
    want = [int(np.any(_np(ref)[b] + (1 - ps[b].astype(float)) == 0)) for b in range(3)]
    assert want == [0, 1, 0]
    planned = W in (300, 320, 420, 840, 1440)      # the row pass decides
    assert [int(v != 0) for v in _np(flags)] == (want if planned else [1, 1, 1])
    rms = ops64.to_device(np.abs(rng.standard_normal((3, H, W))) + 0.1)
    a = ops64.snr_mask(out.clone(), rms.clone(), psT, svT, 1, 48, 20, False, zeroFlags=flags)
    b = ops64.snr_mask(ref.clone(), rms.clone(), psT, svT, 1, 48, 20, False)
    for x, y in zip(a, b):
        assert np.array_equal(_np(x), _np(y))
    assert not _np(a[2])[1].any() and _np(a[2])[0].any()      # the all-zero image is trimmed away entirely, the random one is not
    report("zero_flags", shape=shape, flags=[int(v) for v in _np(flags)])


# ----------------------------------------------------------------------------------------------------------
NONPOW2 = [(96, 100), (300, 420), (420, 300), (257, 1024), (1024, 330), (840, 1440), (1440, 840), (256, 320), (320, 300), (1031, 1032),
           (1031, 1033), (301, 421), (128, 65), (840, 1441)]      # odd widths: Wc = (W + 1) / 2, no Nyquist column


@pytest.mark.parametrize("shape", NONPOW2)
def test_rfft2_irfft2_any_size_f64(ops64, report, shape):
    """Sides that are not powers of two go through the Bluestein passes (exact DFT of the tile, no padding)."""
    rng = np.random.default_rng(31)
    H, W = shape
    x = rng.standard_normal((3, H, W))
    got = _np(ops64.rfft2(ops64.to_device(x), apodPix=20))
    from oracle import pixell_shim as enmap
    ref = (np.fft.rfft2(enmap.apod(x, 20)) / np.sqrt(H * W))
    err = _relerr(got, ref)
    spec = ops64.to_device(np.fft.rfft2(x), ops64.cdtype)
    assert got.shape[-1] == W // 2 + 1 == spec.shape[-1]
    back = _np(ops64.irfft2(spec, W, scale=1.0 / (H * W)))
    err2 = _relerr(back, x)
    report("rfft2_any_size", shape=shape, relerr=err, relerr_inverse=err2)
    assert err < 1e-11 and err2 < 1e-11


@pytest.mark.parametrize("shape", [(300, 420), (257, 1024), (1024, 330), (840, 1440), (320, 256), (301, 421), (1031, 1033)])
def test_apply_filter_and_window_any_size_f64(ops64, report, shape):
    rng = np.random.default_rng(32)
    H, W = shape
    nT, F, S = 2, 2, 2
    Wc = W // 2 + 1
    maps = rng.standard_normal((nT, F, H, W))
    fm = np.fft.fft2(maps) / np.sqrt(H * W)
    filt = rng.standard_normal((nT * S, F, H, W))
    filt = 0.5 * (filt + np.roll(np.flip(filt, axis=(-2, -1)), (1, 1), axis=(-2, -1)))      # real transform needs an even filter
    norm = rng.uniform(0.5, 2.0, nT * S)
    mask = (rng.random((nT, H, W)) > 0.1).astype(np.uint8)
    fmapsT = ops64.rfft2(ops64.to_device(maps.reshape(nT * F, H, W))).reshape(nT, F, H, Wc)
    got = _np(ops64.apply_filter(fmapsT, ops64.to_device(np.ascontiguousarray(filt[..., :Wc])), S, norm=ops64.to_device(norm),
                                 mask=ops64.to_device(mask, ops64.torch.uint8)))
    ref = np.empty((nT * S, H, W))
    for o in range(nT * S):
        ref[o] = np.real(np.fft.ifft2(fm[o // S] * filt[o]) * H * W).sum(axis=0) * norm[o] * mask[o // S]
    err = _relerr(got, ref)
    # pixel window undo with zero restore
    from oracle import pixell_shim as enmap
    d = rng.standard_normal((2, H, W))
    d[:, 10:20, 30:50] = 0
    gotW = _np(ops64.apply_window(ops64.to_device(d.copy()), -1.0, restoreZeros=True))
    refW = np.array([enmap.apply_window(x, pow=-1.0) for x in d])
    refW[d == 0] = 0
    errW = _relerr(gotW, refW)
    report("apply_window_any_size", shape=shape, relerr_apply=err, relerr_window=errW)
    assert err < 1e-10 and errW < 1e-10


@pytest.mark.parametrize("shape,smooth,thr", [((64, 64), 0.0, 1.0), ((256, 320), 1.5, 0.3), ((1024, 1024), 2.0, 0.5),
                                               ((512, 2048), 3.0, 0.0), ((300, 301), 1.0, -0.2), ((1024, 1024), 0.0, 4.0)])
def test_label4_and_moments(ops64, report, shape, smooth, thr):
    rng = np.random.default_rng(13)
    B = 2
    m = rng.standard_normal((B,) + shape)
    if smooth > 0:
        m = np.array([ndimage.gaussian_filter(x, smooth) for x in m])
        m /= m.std()
    labels, nLabels = ops64.label4(ops64.to_device(m), thr)
    labels, nLabels = _np(labels), _np(nLabels)
    nmax = int(nLabels.max())
    table = _np(ops64.label_moments(ops64.to_device(m), ops64.to_device(labels, ops64.torch.int32), nmax))
    for b in range(B):
        ref, n = ndimage.label(m[b] > thr)
        assert nLabels[b] == n
        assert np.array_equal(labels[b], ref), "labels differ"
        ids = np.unique(ref)
        com = np.array(ndimage.center_of_mass(m[b], labels=ref, index=ids))
        npix = ndimage.sum(m[b] > thr, labels=ref, index=ids)
        mpos = np.array(ndimage.maximum_position(m[b], labels=ref, index=ids))
        t = table[b]
        gotCom = np.stack([t[ids, 1] / t[ids, 0], t[ids, 2] / t[ids, 0]], axis=1)
        if ids[0] == 0:
            # background moments can cancel catastrophically; compare the sums instead
            gotCom, com = gotCom[1:], com[1:]
        assert np.abs(gotCom - com).max() < 1e-9 if len(com) else True
        cnt = t[ids, 3].copy()
        if ids[0] == 0:
            cnt[0] = 0
        assert np.array_equal(cnt, npix)
        gotPos = np.stack([t[ids, 5] // shape[1], t[ids, 5] % shape[1]], axis=1)
        assert np.array_equal(gotPos, mpos)
    report("label4", shape=shape, smooth=smooth, thr=thr, nlabels=int(nLabels[0]))


@pytest.mark.parametrize("shape,smooth,thr,sign", [((64, 64), 0.0, 2.5, 1.0), ((256, 320), 1.5, 2.0, 1.0), ((1024, 1024), 2.0, 3.0, 1.0),
                                                    ((512, 2048), 3.0, 2.5, -1.0), ((300, 301), 1.0, 4.0, 1.0), ((1024, 1024), 0.0, 4.0, 1.0)])
def test_detect_sparse_matches_scipy(ops64, report, shape, smooth, thr, sign):
    """nb_detect (one dense pass + compact foreground list) against ndimage.label / center_of_mass / sum /
    maximum_position, labels bit-exact."""
    rng = np.random.default_rng(17)
    B = 3
    m = rng.standard_normal((B,) + shape)
    if smooth > 0:
        m = np.array([ndimage.gaussian_filter(x, smooth) for x in m])
        m /= m.std()
    m[2] = -5.0                                   # an image with no foreground at all
    if sign < 0:
        m[2] = 5.0
    labels, nLabels, status, table = ops64.detect(ops64.to_device(m), thr, sign=sign, wantArgmax=True)
    labels, nLabels, status, table = _np(labels), _np(nLabels), _np(status), _np(table)
    assert not status.any()
    for b in range(B):
        v = sign * m[b]
        ref, n = ndimage.label(v > thr)
        assert nLabels[b] == n
        assert np.array_equal(labels[b], ref), "labels differ"
        if n == 0:
            assert table[b, 0, 3] == v.size
            continue
        ids = np.arange(1, n + 1)
        com = np.array(ndimage.center_of_mass(v, labels=ref, index=ids))
        t = table[b]
        gotCom = np.stack([t[ids, 1] / t[ids, 0], t[ids, 2] / t[ids, 0]], axis=1)
        assert np.abs(gotCom - com).max() < 1e-9
        assert np.array_equal(t[ids, 3], ndimage.sum(v > thr, labels=ref, index=ids))
        mpos = np.array(ndimage.maximum_position(v, labels=ref, index=np.arange(0, n + 1)))
        gotPos = np.stack([t[:n + 1, 5] // shape[1], t[:n + 1, 5] % shape[1]], axis=1)
        assert np.array_equal(gotPos, mpos)
        assert np.isclose(t[0, 0], v[ref == 0].sum(), rtol=1e-9, atol=1e-6) and t[0, 3] == (ref == 0).sum()
        first = np.array([np.flatnonzero(ref.ravel() == i)[0] for i in ids])
        assert np.array_equal(t[ids, 6], first)
    report("detect_sparse", shape=shape, thr=thr, nlabels=int(nLabels[0]))


def test_detect_overflow_falls_back_to_dense_path(ops64):
    """A threshold low enough that more than H*W/8 pixels are foreground sets the status flag; photometry._segment then
    uses the dense kernels and still agrees with scipy."""
    from nemo_b200 import photometry
    rng = np.random.default_rng(18)
    m = rng.standard_normal((2, 128, 160))
    _, _, status, _ = ops64.detect(ops64.to_device(m), 0.0)
    assert _np(status).all()
    labels, n, table = photometry._segment(ops64, ops64.to_device(m), 0.0)
    for b in range(2):
        ref, nref = ndimage.label(m[b] > 0.0)
        assert n[b] == nref and np.array_equal(_np(labels)[b], ref)
    # many small components (> 8192 roots) in a large image also overflow
    big = rng.standard_normal((1, 1024, 1024))
    _, nl, status, _ = ops64.detect(ops64.to_device(big), 1.6)
    assert _np(status)[0] == 1 or _np(nl)[0] <= 8192
    labels, n, table = photometry._segment(ops64, ops64.to_device(big), 1.6)
    ref, nref = ndimage.label(big[0] > 1.6)
    assert n[0] == nref and np.array_equal(_np(labels)[0], ref)


def test_label4_edge_cases(ops64):
    t = ops64.torch
    for m in [np.zeros((1, 64, 64)), np.ones((1, 64, 64)), np.eye(64)[None], np.indices((64, 64)).sum(axis=0)[None] % 2.0]:
        labels, n = ops64.label4(ops64.to_device(m), 0.5)
        ref, nref = ndimage.label(m[0] > 0.5)
        assert int(_np(n)[0]) == nref
        assert np.array_equal(_np(labels)[0], ref)
    # spiral: one long snake, stresses the union-find
    s = np.zeros((1, 128, 128))
    for k in range(0, 64, 2):
        s[0, k, k:128 - k] = 1
        s[0, k:128 - k, 127 - k] = 1
        s[0, 127 - k, k:128 - k] = 1
        s[0, k + 2:128 - k, k] = 1
    labels, n = ops64.label4(ops64.to_device(s), 0.5)
    ref, nref = ndimage.label(s[0] > 0.5)
    assert int(_np(n)[0]) == nref and np.array_equal(_np(labels)[0], ref)
    # inverted sign (invertMap)
    rng = np.random.default_rng(3)
    m = rng.standard_normal((1, 100, 130))
    labels, n = ops64.label4(ops64.to_device(m), 1.0, sign=-1.0)
    ref, nref = ndimage.label(-m[0] > 1.0)
    assert np.array_equal(_np(labels)[0], ref)


def test_spline_eval_f64(ops64, report):
    rng = np.random.default_rng(14)
    B, H, W = 2, 200, 260
    img = rng.standard_normal((B, H, W))
    K = 300
    yx = np.stack([rng.uniform(26, H - 27, K), rng.uniform(26, W - 27, K)], axis=1)
    yx[:5] = np.round(yx[:5])               # exactly on pixels
    idx = rng.integers(0, B, K).astype(np.int32)
    got = _np(ops64.spline_eval(ops64.to_device(img), ops64.to_device(idx, ops64.torch.int32), ops64.to_device(yx)))
    sp = [interpolate.RectBivariateSpline(np.arange(H), np.arange(W), img[b], kx=3, ky=3) for b in range(B)]
    ref = np.array([sp[idx[k]](yx[k, 0], yx[k, 1])[0][0] for k in range(K)])
    err = float(np.abs(got - ref).max())
    report("spline_eval", abserr=err)
    assert err < 1e-11


@pytest.mark.parametrize("sig", [(3.0, 3.0), (7.3, 11.9), (1.0, 25.0), (48.0, 82.0), (130.0, 5.0)])
def test_gauss_filter_f64(ops64, report, sig):
    """Radii up to 127 pass scipy's weights as a kernel argument; wider ones (background-subtraction scales of tens of
    pixels, radius 4 sigma) tabulate them in shared memory -- the reflect boundary then folds more than once."""
    rng = np.random.default_rng(15)
    x = rng.standard_normal((2, 200, 330))
    ref = np.array([ndimage.gaussian_filter(a, sig) for a in x])
    got = _np(ops64.gauss_filter(ops64.to_device(x), sig[0], sig[1]))
    got2 = _np(ops64.gauss_filter(ops64.to_device(x), sig[0], sig[1], subtract=True))
    err = _relerr(got, ref)
    err2 = _relerr(got2, x - ref)
    report("gauss_filter", sig=sig, relerr=err, relerr_sub=err2)
    assert err < 1e-12 and err2 < 1e-12


@pytest.mark.parametrize("k", [(29, 29), (5, 7), (1, 1), (63, 3)])
def test_conv2d_f64(ops64, report, k):
    rng = np.random.default_rng(16)
    B, F, H, W = 2, 2, 150, 170
    x = rng.standard_normal((B, F, H, W))
    kern = rng.standard_normal((F,) + k)
    norm = rng.random(B) + 0.5
    mask = (rng.random((B, H, W)) > 0.2).astype(np.uint8)
    ref = np.array([sum(ndimage.convolve(x[b, f], kern[f]) for f in range(F)) * norm[b] * 0.7 * mask[b] for b in range(B)])
    got, planes = ops64.conv2d(ops64.to_device(x), ops64.to_device(kern, ops64.torch.float64), scale=0.7, norm=ops64.to_device(norm),
                               mask=ops64.to_device(mask, ops64.torch.uint8), wantPlanes=True)
    err = _relerr(_np(got), ref)
    errP = _relerr(_np(planes)[1, 0], ndimage.convolve(x[1, 0], kern[0]))
    report("conv2d", k=k, relerr=err, relerr_plane=errP)
    assert err < 1e-12 and errP < 1e-12


@pytest.mark.parametrize("k,shape", [((29, 29), (200, 264)), ((5, 7), (97, 131)), ((1, 1), (64, 64)), ((63, 3), (128, 70))])
def test_conv2d_batch_f64(ops64, report, k, shape):
    """nb_conv2d_batch (register-blocked K13 with a kernel per image) against scipy.ndimage.convolve: templates shared by the
    tiles (inMod) and the filters of a tile (inDiv), normalisation and mask fused."""
    from scipy import ndimage
    rng = np.random.default_rng(41)
    H, W = shape
    nT, S, F = 3, 2, 2
    nOut = nT * S
    kern = rng.standard_normal((nOut, F) + k)
    norm = rng.uniform(0.5, 2.0, nOut)
    mask = (rng.random((nT, H, W)) > 0.1).astype(np.uint8)
    tiles = rng.standard_normal((nT, F, H, W))
    templ = rng.standard_normal((S, F, H, W))
    kT = ops64.to_device(kern, ops64.torch.float64)
    got = _np(ops64.conv2d_batch(ops64.to_device(tiles), kT, nOut, norm=ops64.to_device(norm, ops64.torch.float64),
                                 mask=ops64.to_device(mask, ops64.torch.uint8), inDiv=S, maskDiv=S))
    gotT, planes = ops64.conv2d_batch(ops64.to_device(templ), kT, nOut, inMod=S, wantPlanes=True)
    gotT, planes = _np(gotT), _np(planes)
    worst = 0.0
    for o in range(nOut):
        ref = sum(ndimage.convolve(tiles[o // S, f], kern[o, f]) for f in range(F)) * norm[o] * mask[o // S]
        refT = [ndimage.convolve(templ[o % S, f], kern[o, f]) for f in range(F)]
        worst = max(worst, _relerr(got[o], ref), _relerr(gotT[o], sum(refT)), _relerr(planes[o], np.array(refT)))
    # one common kernel for every image
    got1 = _np(ops64.conv2d_batch(ops64.to_device(tiles), ops64.to_device(kern[0], ops64.torch.float64), nT))
    for t in range(nT):
        worst = max(worst, _relerr(got1[t], sum(ndimage.convolve(tiles[t, f], kern[0, f]) for f in range(F))))
    report("conv2d_batch_f64", k=k, shape=shape, relerr=worst)
    assert worst < 1e-12


def test_mean_over_rows_f64(ops64, report):
    rng = np.random.default_rng(42)
    x = rng.standard_normal((5, 301, 211))
    got = _np(ops64.mean_over_rows(ops64.to_device(x)))
    err = _relerr(got, x.mean(axis=1))
    report("mean_over_rows", relerr=err)
    assert err < 1e-13


@pytest.mark.parametrize("shape", [(840, 1440), (300, 420)])
def test_rfft2_irfft2_smooth_size_f32(ops32, report, shape):
    """The compile-time mixed-radix passes in fp32 storage mode (sides 2-3-5-7-smooth with a plan in sfft_host.cuh)."""
    rng = np.random.default_rng(33)
    H, W = shape
    x = rng.standard_normal((2, H, W)).astype(np.float32)
    got = _np(ops32.rfft2(ops32.to_device(x), apodPix=20))
    from oracle import pixell_shim as enmap
    ref = (np.fft.rfft2(enmap.apod(x.astype(np.float64), 20)) / np.sqrt(H * W))
    err = _relerr(got, ref)
    spec = ops32.to_device(np.fft.rfft2(x.astype(np.float64)).astype(np.complex64), ops32.cdtype)
    back = _np(ops32.irfft2(spec, W, scale=1.0 / (H * W)))
    err2 = _relerr(back, x)
    report("rfft2_smooth_f32", shape=shape, relerr=err, relerr_inverse=err2)
    assert err < 2e-5 and err2 < 2e-5


# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,spread", [(1, 1.0), (2, 0.001), (300, 0.3), (5000, 1.5), (20000, 6.0)])
def test_xmatch_keep_best_matches_host_rule(ops64, report, n, spread):
    """nb_xmatch_keep_best (every centre against every row on the device) against the host rule of catalogs._crossMatchKeepBest,
    itself tested against the reference's quadratic loop (test_host.py) and pinned to a reference run (ref_misc.npz opt_*):
    crowded fields, equal S/N values (lowest row index wins), rows across RA = 0, and the centres split into three ranges whose
    masks are OR-ed, as the ranks of a multi-GPU run do."""
    from nemo_b200 import catalogs
    rng = np.random.default_rng(n)
    nObj = max(1, n // 9)
    ra0 = (359.8 + rng.random(nObj) * spread) % 360.0
    dec0 = -42.0 + rng.random(nObj) * spread
    pick = rng.integers(0, nObj, n)
    ra = (ra0[pick] + rng.normal(0, 0.006, n)) % 360.0
    dec = dec0[pick] + rng.normal(0, 0.006, n)
    snr = np.round(4 + rng.random(n) * 6, 1)          # one decimal: plenty of ties
    want = catalogs._crossMatchKeepBest(ra, dec, snr, catalogs.XMATCH_RADIUS_DEG)
    got = _np(ops64.xmatch_keep_best(ra, dec, snr, catalogs.XMATCH_RADIUS_DEG)) != 0
    assert np.array_equal(got, want)
    parts = np.zeros(n, dtype=bool)
    per = (n + 2) // 3
    for r in range(3):
        first = min(r * per, n)
        parts |= _np(ops64.xmatch_keep_best(ra, dec, snr, catalogs.XMATCH_RADIUS_DEG, first, min(per, n - first))) != 0
    assert np.array_equal(parts, want)
    assert np.array_equal(catalogs.crossMatchKeepBest(ra, dec, snr, catalogs.XMATCH_RADIUS_DEG, ops=ops64), want)
    report("xmatch_keep_best", n=n, removed=int(want.sum()))
