"""
Batched device pipeline for the filter -> noise -> S/N hot path.  Everything here works on torch CUDA tensors
(handles) through :class:`nemo_b200.ops.Ops`; it is shared by the class API in ``filters.py`` (one tile, one
scale) and by the tile x scale batches of ``pipelines.py``.  Image index convention for a batch of nT tiles
and S filter scales: o = tile * S + scale.

Reference arithmetic restated (nemo/filters.py): :526-529 input FFTs, :568-587 noise covariance, :590-630 filter,
:635-688 normalisation, :699/:824-859 application, :716-758 noise / S-N / masks, :101-104 pixel-window undo.
"""

import os

import numpy as np

from . import signals
from . import maps as nmaps


FUSED_FILTER_BUILD = os.environ.get('NEMO_B200_FUSED_FILTER_BUILD', '0') == '1'


# ------------------------------------------------------------------------------------------------------
# host-side parameter logic
def noiseGrid(shape, noiseParams, pixDeg):
    """(yChunks, xChunks, overlapPix, numBins, estimator) of MapFilter.makeNoiseMap's grid branch (filters.py:412-429), or
    ('smart', numBins) for its 'smart' branch (:366-372)."""
    estimator = 'clip'
    if 'RMSEstimator' in noiseParams and noiseParams['RMSEstimator'] == 'biweight':
        estimator = 'biweight'
    if 'RMSEstimator' in noiseParams and noiseParams['RMSEstimator'] == 'percentile':
        estimator = 'percentile'
    if noiseParams.get('noiseGridArcmin') == "smart":
        if estimator != 'clip':
            raise Exception("RMSEstimator '%s' with noiseGridArcmin 'smart' is not implemented in nemo_b200" % estimator)
        # whole-map weight bins instead of grid cells (filters.py:366-372): ('smart', numBins)
        try:
            numBins = noiseParams['numNoiseBins']
        except Exception:
            raise Exception("Need to give numNoiseBins in noiseParams when using noiseGridArcmin = 'smart'")
        return ('smart', int(numBins))
    if 'noiseGridArcmin' not in noiseParams.keys() or noiseParams['noiseGridArcmin'] is None:
        overlapPix = 0
        numXChunks = 1
        numYChunks = 1
    else:
        gridSize = int(round((noiseParams['noiseGridArcmin'] / 60.) / pixDeg))
        overlapPix = int(gridSize / 2)
        numXChunks = shape[1] / gridSize
        numYChunks = shape[0] / gridSize
    yChunks = np.linspace(0, shape[0], int(numYChunks + 1), dtype=int)
    xChunks = np.linspace(0, shape[1], int(numXChunks + 1), dtype=int)
    numBins = noiseParams['numNoiseBins'] if 'numNoiseBins' in noiseParams.keys() else 1
    # the kernels take the percentile list np.arange(0, 100, 100 / numBins) (filters.py:450) to have numBins entries; floating-point
    # steps give one more for a few values (29, 31, 58, ...), where the reference ends up with numBins + 1 bins
    if len(np.arange(0, 100, 100. / numBins)) != int(numBins):
        raise Exception("numNoiseBins = %d: np.arange(0, 100, 100 / numBins) has %d entries, so the reference makes %d weight bins; "
                        "use a neighbouring value" % (numBins, len(np.arange(0, 100, 100. / numBins)), len(np.arange(0, 100, 100. / numBins))))
    return yChunks, xChunks, overlapPix, int(numBins), estimator


def trimSizePix(params, pixDeg, realSpace):
    """Edge-trim window of the rank filter (filters.py:728-735 Fourier class, :1137-1141 real-space class)."""
    if realSpace:
        if 'edgeTrimArcmin' in params.keys():
            return int(round((params['edgeTrimArcmin'] / 60.) / pixDeg))
        gridSize = int(round((params['noiseParams']['noiseGridArcmin'] / 60.) / pixDeg))
        return int(round(gridSize * 3.0))
    if 'edgeTrimArcmin' in params.keys() and params['edgeTrimArcmin'] > 0:
        return int(round((params['edgeTrimArcmin'] / 60.) / pixDeg))
    elif 'noiseGridArcmin' in params['noiseParams'] and params['noiseParams']['noiseGridArcmin'] != "smart" \
            and params['noiseParams']['noiseGridArcmin'] is not None:
        gridSize = int(round((params['noiseParams']['noiseGridArcmin'] / 60.) / pixDeg))
        return int(round(gridSize * 3.0))
    return 0


def spectralWeights(mapDicts, params):
    """Signal frequency weighting w (filters.py:589-611)."""
    w = []
    for mapDict in mapDicts:
        if mapDict['units'] != 'yc':
            if 'specWeight' in mapDict.keys():
                w.append(mapDict['specWeight'])
            else:
                if params['outputUnits'] == 'yc':
                    w.append(signals.fSZ(mapDict['obsFreqGHz']))
                elif params['outputUnits'] == 'uK':
                    if 'alpha' in params.keys() and params['alpha'] is not None:
                        w.append(np.power(mapDict['obsFreqGHz'] / mapDicts[0]['obsFreqGHz'], params['alpha']))
                    else:
                        w.append(1.0)
                else:
                    raise Exception('need to specify "outputUnits" ("yc" or "uK") in filter params')
        else:
            w.append(1.0)
    return np.array(w, dtype=np.float64)


def normAmplitudes(mapDicts, params, y0=2e-4):
    """Amplitudes handed to makeSignalTemplateMap for the normalisation templates (filters.py:639-646); None for 'uK'."""
    if params['outputUnits'] != 'yc':
        return None
    amps = []
    for mapDict in mapDicts:
        if mapDict['units'] == 'yc':
            amps.append(y0)
        else:
            amps.append(nmaps.convertToDeltaT(y0, mapDict['obsFreqGHz']))
    return amps


# ------------------------------------------------------------------------------------------------------
# device pipeline
def templateSpectra(ops, templUnit, templNorm, outputUnits):
    """(fsig, fsigN): spectra of the unit templates (filters.py:590-611 takes their moduli) and of the normalisation templates
    with the pixel window applied (:639-652; the unit spectra again for 'uK', :678-686).  A function of the templates only:
    FilterSetTemplates keeps them per tile geometry, so that the tiles of a survey do not repeat these transforms."""
    nS, F, H, W = templUnit.shape
    Wc = W // 2 + 1
    fsig = ops.rfft2(templUnit.reshape(nS * F, H, W), apodPix=0).reshape(nS, F, H, Wc)
    if outputUnits == 'yc':
        tn = templNorm.reshape(nS * F, H, W).clone()
        ops.apply_window(tn, 1.0, restoreZeros=False)   # enmap.apply_window(signalMap, pow=1.0), filters.py:647
        fsigN = ops.rfft2(tn, apodPix=0).reshape(nS, F, H, Wc)
    elif outputUnits == 'uK':
        fsigN = fsig
    else:
        raise Exception('need to specify "outputUnits" ("yc" or "uK") in filter params')
    return fsig, fsigN


def buildFourierFilters(ops, fmapsNoise, templUnit, templNorm, w, outputUnits, cy, cx, y0=2e-4, sigPerTile=False,
                        wantFull=False, bckSubSigmas=None, spectra=None):
    """Builds matched filters for nT tiles x S scales.

    fmapsNoise [nT, F, H, Wc]    spectra the noise covariance is measured from
    templUnit  [nS*, F, H, W]    unit-normalised signal templates (nS* = S, or nT*S when sigPerTile)
    templNorm  same shape        templates with the normalisation amplitude, pixel window NOT yet applied ('yc');
                                 ignored for 'uK' (the unit templates are used, filters.py:678-686)
    spectra    (fsig, fsigN) of templateSpectra() when they were made before (kept per tile geometry by the batched drivers)
    Returns dict: filtH [nT*S, F, H, Wc], filtFull (or None), signalNorm [nT*S] fp64, fRelWeights [nT*S, F] (or None),
    all on the device; nothing is synchronised.
    """
    torch = ops.torch
    nT, F, H, Wc = fmapsNoise.shape
    W = int(templUnit.shape[-1])          # (odd widths: Wc = (W + 1) / 2, so the width is not implied by Wc)
    nS = templUnit.shape[0]
    S = nS // nT if sigPerTile else nS
    nOut = nT * S
    # spectra: (fsig, fsigN) made before (templateSpectra), e.g. once per tile geometry
    fsig, fsigN = spectra if spectra is not None else templateSpectra(ops, templUnit, templNorm, outputUnits)
    if FUSED_FILTER_BUILD:
        # K2 + K3 in one kernel, the covariance staying in shared memory: saves the two fp64 [nT, P, H, W] arrays, but the 12-pixel
        # halos of its 32 x 32 tiles move more bytes through L2 than the covariance costs in HBM (0.51 ms against 0.39 ms for
        # 16 tiles of 1024^2, F = 1), so it is off unless memory is the constraint
        filtH, filtFull = ops.build_filter_fused(fmapsNoise, fsig, w, S, sigPerTile, sigma=3.0, wantFull=wantFull, W=W)
    else:
        # K2: smoothed noise covariance (filters.py:568-587); K3: per-mode solve (:624-630)
        cov = ops.xpower_smooth(fmapsNoise, 3.0, W=W)
        filtH, filtFull = ops.build_filter(cov, fsig, w, S, sigPerTile, wantFull=wantFull)
        del cov
    if outputUnits == 'yc':
        mode, numer = 1, y0
    elif outputUnits == 'uK':
        mode, numer = 0, 1.0
    else:
        raise Exception('need to specify "outputUnits" ("yc" or "uK") in filter params')
    # filteredSignal = applyFilter(fSignalMaps) with signalNorm still 1 (filters.py:653, 686)
    if sigPerTile:
        filteredSignal = ops.apply_filter(fsigN, filtH, 1, W=W)
    else:
        filteredSignal = ops.apply_filter(fsigN, filtH, S, perScale=True, W=W)
    if bckSubSigmas is not None:
        filteredSignal = ops.gauss_filter(filteredSignal, bckSubSigmas[0], bckSubSigmas[1], subtract=True)
    signalNorm, argmaxIdx, maxVal, centreVal = ops.template_norm(filteredSignal, mode, cy, cx, numer)
    fRel = None
    if outputUnits == 'yc':
        # (the planes add up to filteredSignal -- unless a background was subtracted from it -- so the last one is the remainder)
        planes = ops.plane_values_at(fsigN, filtH, argmaxIdx, S=0 if sigPerTile else S, W=W, total=maxVal if bckSubSigmas is None else None)
        # fRelWeight = plane value at the argmax pixel / total signal there (filters.py:664-670)
        fRel = (planes, maxVal)
    return {'filtH': filtH, 'filtFull': filtFull, 'signalNorm': signalNorm, 'fRel': fRel, 'S': S}


def applyFourierFilters(ops, fmaps, filtH, S, signalNorm, psMask, bckSubSigmas=None, W=None, zeroFlags=False):
    """MatchedFilter.applyFilter (filters.py:824-859) for every (tile, scale), times psMask (:717).  W: the map width when it
    is odd (or always); by default that of psMask, else the even width that goes with the half spectra.  zeroFlags=True: returns
    (maps, flags): per image, 0 when the map has no pixel for the edge filter of noiseSNAndMasks to find (None when not known)."""
    if bckSubSigmas is None:
        if zeroFlags and psMask is not None:
            return ops.apply_filter(fmaps, filtH, S, norm=signalNorm, mask=psMask, W=W, zeroFlags=True)
        m = ops.apply_filter(fmaps, filtH, S, norm=signalNorm, mask=psMask, W=W)
        return (m, None) if zeroFlags else m
    if zeroFlags:
        return applyFourierFilters(ops, fmaps, filtH, S, signalNorm, psMask, bckSubSigmas=bckSubSigmas, W=W), None
    m = ops.apply_filter(fmaps, filtH, S, W=W if W is not None else (int(psMask.shape[-1]) if psMask is not None else None))
    m = ops.gauss_filter(m, bckSubSigmas[0], bckSubSigmas[1], subtract=True)
    maskFull = psMask if S == 1 or psMask is None else psMask.repeat_interleave(S, dim=0)
    return ops.scale_mask(m, norm=signalNorm, mask=maskFull)


def noiseSNAndMasks(ops, filtered, wmed, psMask, surveyMask, S, grid, trimPix, apodPix, realSpace, wantRMS=True, zeroFlags=None):
    """makeNoiseMap + S/N + edge trim + masks (filters.py:719-758).  `filtered` is overwritten with `data`.
    Returns (data, SNMap, surveyMaskOut u8, RMSMap, cellTable, cellEdges); RMSMap is None with wantRMS=False on the grid path
    (the map is then never materialised: the S/N map takes its RMS straight from the cell table)."""
    if isinstance(grid[0], str) and grid[0] == 'smart':
        rms, table, edges = ops.smart_rms(filtered, wmed, S, grid[1])
        data, sn, maskOut, rms = ops.snr_mask(filtered, rms, psMask, surveyMask, S, trimPix, apodPix, realSpace, zeroFlags=zeroFlags)
        return data, sn, maskOut, rms, table, edges
    yChunks, xChunks, overlapPix, numBins, estimator = grid
    _, table, edges = ops.cell_rms(filtered, wmed, S, yChunks, xChunks, overlapPix, numBins, estimator=estimator, paint=False)
    data, sn, maskOut, rms = ops.snr_mask_cells(filtered, wmed, table, edges, yChunks, xChunks, overlapPix, numBins, psMask, surveyMask,
                                                S, trimPix, apodPix, realSpace, wantRMS=wantRMS, zeroFlags=zeroFlags)
    return data, sn, maskOut, rms, table, edges


def undoPixelWindow(ops, data, tables=None):
    """filterMaps' final step (filters.py:101-104): deconvolve the pixel window, keep exact zeros."""
    return ops.apply_window(data, -1.0, restoreZeros=True, tables=tables)
