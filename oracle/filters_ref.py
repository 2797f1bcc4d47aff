"""
numpy/scipy restatement of the reference's filter classes (/root/reference/nemo/filters.py), fp64.
TEST INFRASTRUCTURE -- see oracle/__init__.py.

  MapFilter.__init__ / makeRadiansMap        filters.py:145-239
  MapFilter.makeNoiseMap                     filters.py:345-483
  MapFilter.makeRealSpaceFilterProfile       filters.py:282-304
  MatchedFilter.buildAndApply / applyFilter  filters.py:524-781, 824-859
  RealSpaceMatchedFilter.buildKernel / buildAndApply / applyFilter   filters.py:892-1218
  filterMaps                                 filters.py:54-109

Differences from the reference that are forced by the environment, not choices:
  * map dictionaries arrive already "preprocessed" (MapDict.preprocess is FITS I/O, out of scope): plain
    dicts with data, weights, surveyMask, pointSourceMask, flagMask, wcs, obsFreqGHz, units, beam, label;
  * filter caches are .npz files with the same content as the reference's FITS (filt + SIGNORM + RW*);
  * the 'model' and 'max(dataMap,CMB)' noise methods need pixell.curvedsky / modlmap and raise here.
"""

import os
import numpy as np
from numpy import fft as npfft
from scipy import ndimage, interpolate

from . import pixell_shim as enmap
from . import signals_ref as signals
from .wcs_stub import calcAngSepDeg


def biweight_scale(data, c=9.0, modify_sample_size=False):
    """astropy.stats.biweight_scale (astropy is absent from this image; third-party dependency of the reference, called at
    nemo/filters.py:385 and :462 with c = 9.0, modify_sample_size = True).  Restated from astropy's published definition
    (astropy.stats.biweight.biweight_midvariance, astropy >= 3.1): with M the median, MAD the median absolute deviation,
    u = (x - M) / (c MAD) and the sums over |u| < 1,

        zeta^2 = n sum (x - M)^2 (1 - u^2)^4 / (sum (1 - u^2)(1 - 5 u^2))^2,

    n = the number of points with |u| < 1 when modify_sample_size, else the sample size; 0 where MAD = 0; scale = sqrt.
    PARITY UNPINNED for this function: there is no astropy here to run it against; tests/test_oracle.py checks the properties
    the definition implies (location / scale behaviour, the Gaussian limit, a hand-evaluated sample)."""
    data = np.asanyarray(data).astype(np.float64)
    M = np.median(data)
    d = data - M
    mad = np.median(np.abs(d))
    if mad == 0.:
        return 0.0
    u = d / (c * mad)
    mask = np.abs(u) >= 1
    u = u ** 2
    n = (~mask).sum() if modify_sample_size else data.size
    f1 = d * d * (1. - u) ** 4
    f1[mask] = 0.
    f1 = f1.sum()
    f2 = (1. - u) * (1. - 5. * u)
    f2[mask] = 0.
    f2 = np.abs(np.sum(f2)) ** 2
    return float(np.sqrt(n * f1 / f2))


class MapFilter(object):

    def __init__(self, label, unfilteredMapsDictList, paramsDict, tileName='PRIMARY', writeFilter=False,
                 forceRebuild=False, diagnosticsDir=None, selFnDir=None):
        self.label = label
        self.params = paramsDict
        self.diagnosticsDir = diagnosticsDir
        self.selFnDir = selFnDir
        self.tileName = tileName
        if diagnosticsDir is not None:
            self.filterFileName = diagnosticsDir + os.path.sep + tileName + os.path.sep + \
                "filter_%s#%s.npz" % (label, tileName)
        else:
            self.filterFileName = None
        self.unfilteredMapsDictList = []
        for mapDict in unfilteredMapsDictList:
            if 'mapToUse' in self.params.keys() and mapDict['label'] != self.params['mapToUse']:
                continue
            self.unfilteredMapsDictList.append(dict(mapDict))
        self.wcs = self.unfilteredMapsDictList[0]['wcs']
        self.shape = self.unfilteredMapsDictList[0]['data'].shape
        self.flagMask = np.zeros(self.shape, dtype=int)
        for i, mapDict in enumerate(self.unfilteredMapsDictList):
            self.flagMask = self.flagMask + (mapDict['flagMask'] * (i + 1))
        self.beamSolidAnglesDict = {}
        for mapDict in self.unfilteredMapsDictList:
            self.beamSolidAnglesDict[mapDict['obsFreqGHz']] = mapDict.get('solidAngle_nsr', 0.0)
        self.apodPix = 20
        for mapDict in self.unfilteredMapsDictList:
            if mapDict['data'].shape != self.shape:
                raise Exception("Maps at different frequencies have different dimensions!")
        self.makeRadiansMap()
        self.signalNorm = 1.0
        self.fRelWeights = {}

    def makeRadiansMap(self):
        shape = self.shape
        x0 = int(shape[1] / 2)
        y0 = int(shape[0] / 2)
        ra0, dec0 = self.wcs.pix2wcs(x0, y0)
        ra1, dec1 = self.wcs.pix2wcs(x0 + 1, y0 + 1)
        self.degPerPixX = calcAngSepDeg(ra0, dec0, ra1, dec0)
        self.degPerPixY = calcAngSepDeg(ra0, dec0, ra0, dec1)
        xRadRange = np.array([np.arange(int(np.floor(-shape[1] / 2)), int(shape[1] / 2), dtype=np.float64)
                              * np.radians(self.degPerPixX)] * shape[0])
        yRadRange = np.array([np.arange(int(np.floor(-shape[0] / 2)), int(shape[0] / 2), dtype=np.float64)
                              * np.radians(self.degPerPixY)] * shape[1]).transpose()
        self.radiansMap = np.sqrt(xRadRange ** 2 + yRadRange ** 2)

    def makeRealSpaceFilterProfile(self):
        realSpace = npfft.ifft2(self.filt).real
        realSpace = npfft.fftshift(realSpace)
        x0 = int(realSpace.shape[2] / 2)
        y0 = int(realSpace.shape[1] / 2)
        normFactor = abs(realSpace[:, y0, x0:]).max()
        prof = realSpace[:, y0, x0:] / normFactor
        arcminRange = np.arange(0, prof.shape[1]) * self.degPerPixX * 60.0
        return prof, arcminRange

    def makeNoiseMap(self, mapData):
        medWeights = np.median(np.array([m['weights'] for m in self.unfilteredMapsDictList]), axis=0)
        noiseParams = self.params['noiseParams']
        biweightEstimator = 'RMSEstimator' in noiseParams and noiseParams['RMSEstimator'] == 'biweight'
        percentileEstimator = 'RMSEstimator' in noiseParams and noiseParams['RMSEstimator'] == 'percentile'
        if (percentileEstimator or biweightEstimator) and noiseParams.get('noiseGridArcmin') == "smart":
            raise Exception("oracle: percentile / biweight estimators restated for the grid branch only")
        if noiseParams.get('noiseGridArcmin') == "smart":
            # 'smart' branch (filters.py:366-407): numBins - 1 whole-map bins between the smallest and the largest median weight,
            # strict inequalities on both sides (a pixel whose weight equals an edge is in no bin)
            try:
                numBins = noiseParams['numNoiseBins']
            except Exception:
                raise Exception("Need to give numNoiseBins in noiseParams when using noiseGridArcmin = 'smart'")
            binEdges = np.linspace(medWeights.min(), medWeights.max(), numBins)
            RMSMap = np.zeros(medWeights.shape)
            apodMask = np.not_equal(mapData, 0)
            for i in range(len(binEdges) - 1):
                binMin = binEdges[i]
                binMax = binEdges[i + 1]
                weightMask = np.logical_and(medWeights > binMin, medWeights < binMax)
                chunkValues = mapData[weightMask]
                goodAreaMask = np.greater_equal(apodMask[weightMask], 1.0)
                if np.not_equal(chunkValues, 0).sum() != 0:
                    chunkMean = np.mean(chunkValues[goodAreaMask])
                    chunkRMS = np.std(chunkValues[goodAreaMask])
                    sigmaClip = 3.0
                    for c in range(10):
                        mask = np.less(abs(chunkValues), abs(chunkMean + sigmaClip * chunkRMS))
                        mask = np.logical_and(goodAreaMask, mask)
                        if mask.sum() > 0:
                            chunkMean = np.mean(chunkValues[mask])
                            chunkRMS = np.std(chunkValues[mask])
                else:
                    chunkRMS = 0.
                if chunkRMS > 0:
                    RMSMap[weightMask] = chunkRMS
            return RMSMap
        if 'noiseGridArcmin' not in noiseParams.keys() or noiseParams['noiseGridArcmin'] is None:
            overlapPix = 0
            numXChunks = 1
            numYChunks = 1
        else:
            gridSize = int(round((noiseParams['noiseGridArcmin'] / 60.) / self.wcs.getPixelSizeDeg()))
            overlapPix = int(gridSize / 2)
            numXChunks = mapData.shape[1] / gridSize
            numYChunks = mapData.shape[0] / gridSize
        yChunks = np.linspace(0, mapData.shape[0], int(numYChunks + 1), dtype=int)
        xChunks = np.linspace(0, mapData.shape[1], int(numXChunks + 1), dtype=int)
        apodMask = np.not_equal(mapData, 0)
        RMSMap = np.zeros(mapData.shape)
        numBins = noiseParams['numNoiseBins'] if 'numNoiseBins' in noiseParams.keys() else 1
        for i in range(len(yChunks) - 1):
            for k in range(len(xChunks) - 1):
                y0 = max(yChunks[i] - overlapPix, 0)
                y1 = min(yChunks[i + 1] + overlapPix, mapData.shape[0])
                x0 = max(xChunks[k] - overlapPix, 0)
                x1 = min(xChunks[k + 1] + overlapPix, mapData.shape[1])
                chunkValues = mapData[y0:y1, x0:x1]
                goodAreaMask = np.greater_equal(apodMask[y0:y1, x0:x1], 1.0)
                if goodAreaMask.sum() == 0:
                    continue
                weightValues = medWeights[y0:y1, x0:x1]
                percentiles = np.arange(0, 100, 100 / numBins)
                binEdges = []
                for p in percentiles:
                    binEdges.append(np.percentile(weightValues[goodAreaMask], p))
                binEdges.append(weightValues[goodAreaMask].max() + 1e-6)
                for b in range(len(binEdges) - 1):
                    binMin = binEdges[b]
                    binMax = binEdges[b + 1]
                    binMask = np.logical_and(weightValues >= binMin, weightValues < binMax)
                    binValues = chunkValues[binMask * goodAreaMask]
                    if biweightEstimator:
                        # filters.py:461-465
                        if (binMask * goodAreaMask).sum() >= 10:
                            chunkRMS = biweight_scale(binValues, c=9.0, modify_sample_size=True)
                        else:
                            chunkRMS = 0.
                    elif percentileEstimator:
                        # filters.py:464-465; an empty bin gives nan (with a numpy warning), which paints nothing
                        chunkRMS = np.percentile(abs(binValues), 68.3) if len(binValues) > 0 else np.nan
                    elif np.not_equal(binValues, 0).sum() != 0:
                        chunkMean = np.mean(binValues)
                        chunkRMS = np.std(binValues)
                        sigmaClip = 3.0
                        for c in range(10):
                            mask = np.less(abs(binValues), abs(chunkMean + sigmaClip * chunkRMS))
                            if mask.sum() > 0:
                                chunkMean = np.mean(binValues[mask])
                                chunkRMS = np.std(binValues[mask])
                    else:
                        chunkRMS = 0.
                    if chunkRMS > 0:
                        RMSMap[y0:y1, x0:x1][binMask] = chunkRMS
        return RMSMap

    # -- the S/N + mask block shared by both classes (filters.py:716-758, 1112-1161)
    def _trimSizePix(self, realSpace):
        if realSpace:
            if 'edgeTrimArcmin' in self.params.keys():
                return int(round((self.params['edgeTrimArcmin'] / 60.) / self.wcs.getPixelSizeDeg()))
            gridSize = int(round((self.params['noiseParams']['noiseGridArcmin'] / 60.) / self.wcs.getPixelSizeDeg()))
            return int(round(gridSize * 3.0))
        if 'edgeTrimArcmin' in self.params.keys() and self.params['edgeTrimArcmin'] > 0:
            return int(round((self.params['edgeTrimArcmin'] / 60.) / self.wcs.getPixelSizeDeg()))
        elif 'noiseGridArcmin' in self.params['noiseParams'] and self.params['noiseParams']['noiseGridArcmin'] != "smart" \
                and self.params['noiseParams']['noiseGridArcmin'] is not None:
            gridSize = int(round((self.params['noiseParams']['noiseGridArcmin'] / 60.) / self.wcs.getPixelSizeDeg()))
            return int(round(gridSize * 3.0))
        return 0

    def _unitsEtc(self, realSpace):
        if self.params['outputUnits'] == 'yc':
            return 'yc', 'yc', 0.0
        elif self.params['outputUnits'] == 'uK':
            if realSpace and len(self.unfilteredMapsDictList) > 1:
                raise Exception("multi-frequency filtering not currently supported for outputUnits 'uK' (point source finding)")
            combinedObsFreqGHz = float(list(self.beamSolidAnglesDict.keys())[0])
            return 'uK', combinedObsFreqGHz, self.beamSolidAnglesDict[combinedObsFreqGHz]
        raise Exception('need to specify "outputUnits" ("yc" or "uK") in filter params')

    def _noiseSNAndMasks(self, filteredMap, surveyMask, psMask, realSpace):
        filteredMap = filteredMap * psMask
        RMSMap = self.makeNoiseMap(filteredMap)
        validMask = np.greater(RMSMap, 0)
        SNMap = np.zeros(filteredMap.shape) + filteredMap
        SNMap[validMask] = SNMap[validMask] / RMSMap[validMask]
        trimSizePix = self._trimSizePix(realSpace)
        if trimSizePix > 0:
            edgeCheck = ndimage.rank_filter(abs(filteredMap + (1 - psMask)), 0, size=(trimSizePix, trimSizePix))
            edgeCheck = np.array(np.greater(edgeCheck, 0), dtype=float)
        else:
            edgeCheck = np.ones(filteredMap.shape)
        filteredMap = filteredMap * edgeCheck
        surveyMask = edgeCheck * surveyMask * psMask
        if not realSpace:
            filteredMap = filteredMap * surveyMask      # filters.py:743 (Fourier class only)
        apodMask = np.equal(enmap.apod(np.ones(filteredMap.shape), self.apodPix), 1)
        surveyMask = surveyMask * apodMask
        SNMap = SNMap * surveyMask
        SNMap[np.isnan(SNMap)] = 0.
        RMSMap = RMSMap * surveyMask
        self.RMSMap = RMSMap
        return filteredMap, SNMap, surveyMask, RMSMap


class MatchedFilter(MapFilter):

    def buildAndApply(self, useCachedFilter=False, vectorised=True):
        fMapsToFilter = []
        for mapDict in self.unfilteredMapsDictList:
            fMapsToFilter.append(enmap.fft(enmap.apod(mapDict['data'], self.apodPix)))
        fMapsToFilter = np.array(fMapsToFilter)
        surveyMask = self.unfilteredMapsDictList[0]['surveyMask']
        psMask = self.unfilteredMapsDictList[0]['pointSourceMask']

        haveCache = self.filterFileName is not None and os.path.exists(self.filterFileName)
        if haveCache == False and useCachedFilter == False:
            fMapsForNoise = []
            for i in range(len(self.unfilteredMapsDictList)):
                mapDict = self.unfilteredMapsDictList[i]
                d = mapDict['data']
                if self.params['noiseParams']['method'] == 'dataMap':
                    if 'noiseModelCatalog' in self.params.keys() and self.params['noiseModelCatalog'] is not None:
                        raise Exception("oracle: noiseModelCatalog needs maps.makeModelImage (out of scope)")
                    fMapsForNoise.append(enmap.fft(enmap.apod(d, self.apodPix)))
                elif self.params['noiseParams']['method'] == 'model':
                    raise Exception("oracle: noise method 'model' needs pixell.curvedsky (absent)")
                else:
                    raise Exception("'%s' is not a valid filter noise method name - fix the .yml config file" % (self.params['noiseParams']['method']))
            fMapsForNoise = np.array(fMapsForNoise)

            kernelSize = (3, 3)
            F = len(self.unfilteredMapsDictList)
            noiseCov = []
            for i in range(F):
                row = []
                for j in range(F):
                    NP = np.real(fMapsForNoise[i] * fMapsForNoise[j].conj())
                    NP = ndimage.gaussian_filter(NP, kernelSize)
                    row.append(NP)
                noiseCov.append(row)
            del fMapsForNoise
            noiseCov = np.array(noiseCov)

            w = []
            for mapDict in self.unfilteredMapsDictList:
                if mapDict['units'] != 'yc':
                    if 'specWeight' in mapDict.keys():
                        w.append(mapDict['specWeight'])
                    else:
                        if self.params['outputUnits'] == 'yc':
                            w.append(signals.fSZ(mapDict['obsFreqGHz']))
                        elif self.params['outputUnits'] == 'uK':
                            if 'alpha' in self.params.keys() and self.params['alpha'] is not None:
                                w.append(np.power(mapDict['obsFreqGHz'] / self.unfilteredMapsDictList[0]['obsFreqGHz'],
                                                  self.params['alpha']))
                            else:
                                w.append(1.0)
                        else:
                            raise Exception('need to specify "outputUnits" ("yc" or "uK") in filter params')
                else:
                    w.append(1.0)
            w = np.array(w)

            fSignalsArr = []
            for mapDict in self.unfilteredMapsDictList:
                signalMap = self.makeSignalTemplateMap(mapDict['beam'])
                fSignalsArr.append(enmap.fft(signalMap))
            fSignalsArr = np.array(fSignalsArr)

            self.filt = np.zeros([F, self.shape[0], self.shape[1]])
            if vectorised:
                # same arithmetic as the loop below, np.linalg.inv broadcast over [H, W, F, F]
                cov = np.ascontiguousarray(np.moveaxis(noiseCov, (0, 1), (2, 3)))
                rhs = np.moveaxis(w[:, None, None] * abs(fSignalsArr), 0, 2)[..., None]
                try:
                    self.filt = np.moveaxis((np.linalg.inv(cov) @ rhs)[..., 0], 2, 0).copy()
                except np.linalg.LinAlgError:
                    vectorised = False
            if not vectorised:
                for y in range(0, self.shape[0]):
                    for x in range(0, self.shape[1]):
                        try:
                            self.filt[:, y, x] = np.dot(np.linalg.inv(noiseCov[:, :, y, x]), w * abs(fSignalsArr[:, y, x]))
                        except Exception:
                            continue
            del fSignalsArr
            del noiseCov

            if self.params['outputUnits'] == 'yc':
                fSignalMaps = []
                y0 = 2e-4
                for mapDict in self.unfilteredMapsDictList:
                    if mapDict['units'] == 'yc':
                        signalMap = self.makeSignalTemplateMap(mapDict['beam'], amplitude=y0)
                    else:
                        deltaT0 = signals.convertToDeltaT(y0, mapDict['obsFreqGHz'])
                        signalMap = self.makeSignalTemplateMap(mapDict['beam'], amplitude=deltaT0)
                    signalMap = enmap.apply_window(signalMap, pow=1.0)
                    fSignalMaps.append(enmap.fft(signalMap))
                fSignalMaps = np.array(fSignalMaps)
                filteredSignal = self.applyFilter(fSignalMaps)
                cRADeg, cDecDeg = self.wcs.getCentreWCSCoords()
                cx, cy = self.wcs.wcs2pix(cRADeg, cDecDeg)
                mapInterpolator = interpolate.RectBivariateSpline(np.arange(filteredSignal.shape[0]),
                                                                  np.arange(filteredSignal.shape[1]),
                                                                  filteredSignal, kx=3, ky=3)
                peakFilteredSignal = mapInterpolator(cy, cx)[0][0]
                self.signalNorm = y0 / peakFilteredSignal
                totalSignal = filteredSignal.flatten()[np.argmax(filteredSignal)]
                filteredSignalCube = np.real(enmap.ifft(fSignalMaps * self.filt, normalize=False))
                self.fRelWeights = {}
                for filteredSignalPlane, mapDict in zip(filteredSignalCube, self.unfilteredMapsDictList):
                    freqGHz = mapDict['obsFreqGHz']
                    self.fRelWeights[freqGHz] = filteredSignalPlane.flatten()[np.argmax(filteredSignal)] / totalSignal
                del fSignalMaps
            elif self.params['outputUnits'] == 'uK':
                fSignalMaps = []
                for mapDict in self.unfilteredMapsDictList:
                    signalMap = self.makeSignalTemplateMap(mapDict['beam'])
                    fSignalMaps.append(enmap.fft(signalMap))
                fSignalMaps = np.array(fSignalMaps)
                filteredSignal = self.applyFilter(fSignalMaps)
                self.signalNorm = 1.0 / filteredSignal.max()
                del fSignalMaps
            else:
                raise Exception('need to specify "outputUnits" ("yc" or "uK") in filter params')
        else:
            self.loadFilter()
            self.params['saveRMSMap'] = False
            self.params['saveFilter'] = False
            self.params['savePlots'] = False

        filteredMap = self.applyFilter(fMapsToFilter)
        del fMapsToFilter
        mapUnits, combinedObsFreqGHz, beamSolidAngle_nsr = self._unitsEtc(False)
        filteredMap, SNMap, surveyMask, RMSMap = self._noiseSNAndMasks(filteredMap, surveyMask, psMask, False)

        if 'saveFilter' in self.params and self.params['saveFilter'] == True and self.filterFileName is not None:
            os.makedirs(os.path.dirname(self.filterFileName), exist_ok=True)
            np.savez(self.filterFileName, filt=self.filt, SIGNORM=self.signalNorm,
                     RW_GHZ=np.array(list(self.fRelWeights.keys()), dtype=float),
                     RW=np.array(list(self.fRelWeights.values()), dtype=float))

        return {'data': filteredMap, 'wcs': self.wcs, 'obsFreqGHz': combinedObsFreqGHz, 'SNMap': SNMap,
                'surveyMask': surveyMask, 'flagMask': self.flagMask, 'mapUnits': mapUnits,
                'beamSolidAngle_nsr': beamSolidAngle_nsr, 'label': self.label, 'tileName': self.tileName}

    def loadFilter(self):
        with np.load(self.filterFileName) as img:
            self.filt = img['filt']
            self.signalNorm = float(img['SIGNORM'])
            self.fRelWeights = {float(k): float(v) for k, v in zip(img['RW_GHZ'], img['RW'])}

    def applyFilter(self, mapDataToFilter):
        if mapDataToFilter.shape == self.filt.shape:
            filt = self.filt
        else:
            raise Exception("oracle: reshapeFilter (pixell.enmap.laxes) not restated")
        if 'complex' in mapDataToFilter.dtype.name:
            fMapsToFilter = mapDataToFilter
        else:
            fMapsToFilter = enmap.fft(enmap.apod(mapDataToFilter, self.apodPix))
        filteredMap = np.real(enmap.ifft(fMapsToFilter * filt, normalize=False)).sum(axis=0)
        if 'bckSub' in self.params.keys() and 'bckSubScaleArcmin' in self.params.keys() and self.params['bckSub'] == True:
            filteredMap = signals.subtractBackground(filteredMap, self.wcs, smoothScaleDeg=self.params['bckSubScaleArcmin'] / 60.)
        return filteredMap * self.signalNorm


class RealSpaceMatchedFilter(MapFilter):

    def buildKernel(self, kernelMapsDictList):
        """filters.py:892-1041.  ``kernelMapsDictList`` is the already-cut noise region (the reference
        re-reads it from FITS with ``RADecSection``; the cut is I/O and done by the caller)."""
        matchedFilterClass = self.params['noiseParams']['matchedFilterClass']
        if isinstance(matchedFilterClass, str):
            matchedFilterClass = globals()[matchedFilterClass]
        matchedFilter = matchedFilterClass("realSpaceKernel_%s" % (self.label), kernelMapsDictList, self.params,
                                           tileName=self.tileName)
        matchedFilter.buildAndApply()
        self.matchedFilter = matchedFilter

        kernelMaxArcmin = self.params['noiseParams']['kernelMaxArcmin']
        prof, arcminRange = matchedFilter.makeRealSpaceFilterProfile()
        rIndex = np.where(arcminRange > kernelMaxArcmin)[0][0]
        mask = np.less(arcminRange, kernelMaxArcmin)
        if 'symmetrize' in self.params['noiseParams'].keys() and self.params['noiseParams']['symmetrize'] == True:
            rRadians = np.radians(arcminRange / 60.)
            profile2d = []
            for i in range(prof.shape[0]):
                r2p = interpolate.interp1d(rRadians[mask], prof[i, mask], bounds_error=False, fill_value=0.0)
                profile2d.append(r2p(matchedFilter.radiansMap))
            profile2d = np.array(profile2d)
        else:
            profile2d = npfft.ifft2(matchedFilter.filt).real
            profile2d = npfft.fftshift(profile2d)
        z, y, x = np.where(abs(profile2d) == abs(profile2d).max())
        y = y[0]
        x = x[0]
        yMin = y - rIndex
        yMax = y + rIndex
        xMin = x - rIndex
        xMax = x + rIndex
        if (yMax - yMin) % 2 == 0:
            yMin = yMin + 1
        if (xMax - xMin) % 2 == 0:
            xMin = xMin + 1
        self.kern2d = profile2d[:, yMin:yMax, xMin:xMax]
        if 'bckSubScaleArcmin' in self.params.keys():
            self.bckSubScaleArcmin = self.params['bckSubScaleArcmin']
        else:
            func = np.min if np.greater(prof[0, 0], 0) == True else np.max
            self.bckSubScaleArcmin = arcminRange[prof[0] == func(prof[0])][0]

        signalMaps = []
        for mapDict in self.unfilteredMapsDictList:
            if self.params['outputUnits'] == 'yc':
                y0 = 2e-4
                if mapDict['obsFreqGHz'] is not None:
                    deltaT0 = signals.convertToDeltaT(y0, mapDict['obsFreqGHz'])
                    signalMap = self.makeSignalTemplateMap(mapDict['beam'], amplitude=deltaT0)
                else:
                    signalMap = self.makeSignalTemplateMap(mapDict['beam'], amplitude=y0)
            elif self.params['outputUnits'] == 'uK':
                signalMap = self.makeSignalTemplateMap(mapDict['beam'])
            else:
                raise Exception('need to specify "outputUnits" ("yc" or "uK") in filter params')
            signalMaps.append(signalMap)
        signalMaps = np.array(signalMaps)
        filteredSignal = self.applyFilter(signalMaps, calcFRelWeights=True)
        if self.params['outputUnits'] == 'yc':
            self.signalNorm = y0 / filteredSignal.max()
        else:
            self.signalNorm = 1.0 / filteredSignal.max()

    def buildAndApply(self, kernelMapsDictList=None, rawMapsToFilter=None):
        surveyMask = self.unfilteredMapsDictList[0]['surveyMask']
        psMask = self.unfilteredMapsDictList[0]['pointSourceMask']
        RAMin, RAMax, decMin, decMax = self.wcs.getImageMinMaxWCSCoords()
        self.applyDecCentre = (decMax + decMin) / 2.
        self.applyRACentre = (RAMax + RAMin) / 2.
        if kernelMapsDictList is None:
            kernelMapsDictList = self.unfilteredMapsDictList
        self.buildKernel(kernelMapsDictList)
        # filters.py:1106-1110: the *raw* tile (loadTile), not the preprocessed map
        if rawMapsToFilter is None:
            rawMapsToFilter = [m['data'] for m in self.unfilteredMapsDictList]
        mapDataToFilter = np.array(rawMapsToFilter, dtype=np.float64)
        filteredMap = self.applyFilter(mapDataToFilter)
        mapUnits, combinedObsFreqGHz, beamSolidAngle_nsr = self._unitsEtc(True)
        filteredMap, SNMap, surveyMask, RMSMap = self._noiseSNAndMasks(filteredMap, surveyMask, psMask, True)
        return {'data': filteredMap, 'wcs': self.wcs, 'obsFreqGHz': combinedObsFreqGHz, 'SNMap': SNMap,
                'surveyMask': surveyMask, 'flagMask': self.flagMask, 'mapUnits': mapUnits,
                'beamSolidAngle_nsr': beamSolidAngle_nsr, 'label': self.label, 'tileName': self.tileName}

    def applyFilter(self, mapDataToFilter, calcFRelWeights=False):
        filteredMap = np.zeros(mapDataToFilter.shape)
        if self.params['bckSub'] == True and self.bckSubScaleArcmin > 0:
            for i in range(mapDataToFilter.shape[0]):
                filteredMap[i] = signals.subtractBackground(mapDataToFilter[i], self.wcs,
                                                            RADeg=self.applyRACentre, decDeg=self.applyDecCentre,
                                                            smoothScaleDeg=self.bckSubScaleArcmin / 60.)
        else:
            filteredMap = filteredMap + mapDataToFilter
        for i in range(filteredMap.shape[0]):
            filteredMap[i] = ndimage.convolve(filteredMap[i], self.kern2d[i])
        if calcFRelWeights == True:
            self.fRelWeights = {}
            maxIndex = np.argmax(filteredMap.sum(axis=0))
            totalSignal = filteredMap.sum(axis=0).flatten()[maxIndex]
            for filteredSignalPlane, mapDict in zip(filteredMap, self.unfilteredMapsDictList):
                self.fRelWeights[mapDict['obsFreqGHz']] = filteredSignalPlane.flatten()[maxIndex] / totalSignal
        filteredMap = filteredMap.sum(axis=0)
        return filteredMap * self.signalNorm


class BeamFilter(MapFilter):
    def makeSignalTemplateMap(self, beam, amplitude=None):
        return signals.makeBeamModelSignalMap(np.degrees(self.radiansMap), self.wcs, beam, amplitude=amplitude)


_arnaudProfileCache = {}


class ArnaudModelFilter(MapFilter):
    def makeSignalTemplateMap(self, beam, amplitude=None):
        key = (self.params['z'], self.params['M500MSun'])
        if key not in _arnaudProfileCache:
            _arnaudProfileCache[key] = signals.makeArnaudModelProfile(self.params['z'], self.params['M500MSun'],
                                                                      GNFWParams=self.params.get('GNFWParams', 'default'))
        tckP = _arnaudProfileCache[key]['tckP']
        degreesMap = np.degrees(self.radiansMap)
        rDeg = np.linspace(0.0, 15.0, 5000)
        profile1d = interpolate.splev(rDeg, tckP, ext=1)
        if amplitude is not None:
            profile1d = profile1d * amplitude
        r2p = interpolate.interp1d(rDeg, profile1d, bounds_error=False, fill_value=0.0)
        signalMap = r2p(degreesMap)
        # beam convolution on the periodic tile grid (stands in for maps.convolveMapWithBeam, maps.py:1371-1430)
        beamMap = signals.makeBeamModelSignalMap(degreesMap, self.wcs, beam)
        beamMap = beamMap / beamMap.sum()
        return np.fft.ifft2(np.fft.fft2(signalMap) * np.fft.fft2(np.fft.ifftshift(beamMap))).real


class ArnaudModelMatchedFilter(MatchedFilter, ArnaudModelFilter):
    pass


class BeamMatchedFilter(MatchedFilter, BeamFilter):
    pass


class ArnaudModelRealSpaceMatchedFilter(RealSpaceMatchedFilter, ArnaudModelFilter):
    pass


class BeamRealSpaceMatchedFilter(RealSpaceMatchedFilter, BeamFilter):
    pass


def filterMaps(unfilteredMapsDictList, filterParams, tileName, diagnosticsDir=None, selFnDir=None,
               verbose=True, undoPixelWindow=True, useCachedFilter=False, returnFilter=False, **kwargs):
    f = filterParams
    filterClass = f['class'] if not isinstance(f['class'], str) else globals()[f['class']]
    filterObj = filterClass(f['label'], unfilteredMapsDictList, f['params'], tileName=tileName,
                            diagnosticsDir=diagnosticsDir, selFnDir=selFnDir)
    if isinstance(filterObj, RealSpaceMatchedFilter):
        filteredMapDict = filterObj.buildAndApply(**kwargs)
    else:
        filteredMapDict = filterObj.buildAndApply(useCachedFilter=useCachedFilter, **kwargs)
    filteredMapDict['wcs'].header['BUNIT'] = filteredMapDict['mapUnits']
    if 'beamSolidAngle_nsr' in filteredMapDict.keys() and filteredMapDict['beamSolidAngle_nsr'] > 0:
        filteredMapDict['wcs'].header['BEAMNSR'] = filteredMapDict['beamSolidAngle_nsr']
        filteredMapDict['wcs'].header['FREQGHZ'] = filteredMapDict['obsFreqGHz']
    filteredMapDict['wcs'].updateFromHeader()
    if undoPixelWindow == True:
        mask = np.equal(filteredMapDict['data'], 0)
        filteredMapDict['data'] = enmap.apply_window(filteredMapDict['data'], pow=-1.0)
        filteredMapDict['data'][mask] = 0
    if returnFilter == True:
        return filteredMapDict, filterObj
    return filteredMapDict
