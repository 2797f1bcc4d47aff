"""
Drop-in for the reference's ``nemo.filters`` on the hot path: same function / class / method names, arguments and
returned dictionary (reference nemo/filters.py:54-109 ``filterMaps``; :112-516 ``MapFilter``; :519-859
``MatchedFilter``; :862-1218 ``RealSpaceMatchedFilter``; :1222-1331 template mix-ins and concrete classes), with
the arithmetic done by hand-written sm_100a CUDA kernels behind ``libnemo_b200.so``.  There is no CPU path.

Differences a caller can see:
  * the maps in the returned dictionary (``data``, ``SNMap``, ``surveyMask``) are :class:`nemo_b200.device.DeviceArray`
    objects -- numpy-compatible (``np.asarray`` works) but kept on the GPU so ``photometry.findObjects`` /
    ``measureFluxes`` do not copy them back and forth;
  * map dictionaries may be plain dicts that already hold ``data``, ``weights``, ``surveyMask``,
    ``pointSourceMask``, ``flagMask``, ``wcs``, ``obsFreqGHz``, ``units``, ``beamFileName`` (a path or a
    ``signals.BeamProfile``); objects with a ``preprocess`` method (the reference's ``maps.MapDict``) are
    preprocessed exactly as the reference does;
  * filter caches, RMS maps and area masks are FITS files in the reference's layout, written and read by ``nemo_b200.fitsio``
    (astropy is not needed; tile compression is not implemented, the images are plain HDUs of the same pixel type);
  * options whose dependencies are absent raise instead of silently differing: noise methods ``model`` and
    ``max(dataMap,CMB)``, ``noiseModelCatalog``, ``RMSEstimator`` percentile / biweight together with the ``smart`` noise grid;
    tile sides must lie in 64..4096 (sides that are not powers of two take mixed-radix passes when
    their prime factors are 2, 3, 5, 7, the slower Bluestein passes otherwise -- odd sides included).
"""

import os
import numpy as np
from numpy import fft
from scipy import interpolate

from . import device
from . import engine
from . import maps
from . import signals
from . import wcs as _wcs
from .device import DeviceArray


# -------------------------------------------------------------------------------------------------------------
def filterMaps(unfilteredMapsDictList, filterParams, tileName, diagnosticsDir='.', selFnDir='.', verbose=True,
               undoPixelWindow=True, useCachedFilter=False, returnFilter=False):
    """Builds and applies filters to the unfiltered map(s); same contract as the reference's ``filterMaps``."""
    f = filterParams
    label = f['label'] + "#" + tileName
    if verbose:
        print("... making filtered map %s" % (label))
    filterClass = f['class'] if not isinstance(f['class'], str) else globals()[f['class']]
    filterObj = filterClass(f['label'], unfilteredMapsDictList, f['params'], tileName=tileName,
                            diagnosticsDir=diagnosticsDir, selFnDir=selFnDir)
    filteredMapDict = filterObj.buildAndApply(useCachedFilter=useCachedFilter)

    # Keywords needed for photometry later
    filteredMapDict['wcs'].header['BUNIT'] = filteredMapDict['mapUnits']
    if 'beamSolidAngle_nsr' in filteredMapDict.keys() and filteredMapDict['beamSolidAngle_nsr'] > 0:
        filteredMapDict['wcs'].header['BEAMNSR'] = filteredMapDict['beamSolidAngle_nsr']
        filteredMapDict['wcs'].header['FREQGHZ'] = filteredMapDict['obsFreqGHz']
    filteredMapDict['wcs'].updateFromHeader()

    if undoPixelWindow == True:
        ops = filterObj.ops
        data = filteredMapDict['data'].tensor.unsqueeze(0)
        engine.undoPixelWindow(ops, data)
        filteredMapDict['data'] = DeviceArray(data[0])

    if returnFilter == True:
        return filteredMapDict, filterObj
    return filteredMapDict


# -------------------------------------------------------------------------------------------------------------
_SOLID_ANGLES = {}


def beamSolidAngle_nsr(mapDict):
    """Beam solid angle (nsr) of a map dictionary as MapFilter.__init__ finds it (filters.py:180-196): the 'solidAngle_nsr' key,
    else the 'solid angle = ... nsr' line of the beam profile file, else 0 (no Jy fluxes).  Used by the class API and by the
    batched driver, so that both report the same fluxJy."""
    if 'solidAngle_nsr' in mapDict.keys():
        return mapDict['solidAngle_nsr']
    beamFileName = mapDict.get('beamFileName')
    if not isinstance(beamFileName, str):
        return 0.0
    key = (beamFileName, os.path.getmtime(beamFileName))
    if key not in _SOLID_ANGLES:
        solidAngle_nsr = 0.0
        with open(beamFileName, "r") as inFile:
            for line in inFile.readlines():
                if line.find("solid angle") != -1:
                    solidAngle_nsr = float(line.split("=")[1].split()[0])
                    break
        _SOLID_ANGLES[key] = solidAngle_nsr
    return _SOLID_ANGLES[key]


def _maskToU8(m, shape):
    if m is None:
        return np.ones(shape, dtype=np.uint8)
    m = np.asarray(m)
    return np.ascontiguousarray(np.not_equal(m, 0), dtype=np.uint8)


_RADIANS_MAPS = {}


class MapFilter(object):
    """Base class: prepares the maps of one tile on the device and provides the noise-map machinery."""

    def __init__(self, label, unfilteredMapsDictList, paramsDict, tileName='PRIMARY', writeFilter=False,
                 forceRebuild=False, diagnosticsDir=None, selFnDir=None):
        self.label = label
        self.params = paramsDict
        self.diagnosticsDir = diagnosticsDir
        self.selFnDir = selFnDir
        self.tileName = tileName
        if diagnosticsDir is not None:
            self.filterFileName = diagnosticsDir + os.path.sep + tileName + os.path.sep + "filter_%s#%s.fits" % (label, tileName)
        else:
            self.filterFileName = None
        self.ops = device.getOps(self.params.get('precision') if hasattr(self.params, 'get') else None)

        self.unfilteredMapsDictList = []
        for mapDict in unfilteredMapsDictList:
            if 'mapToUse' in self.params.keys() and mapDict['label'] != self.params['mapToUse']:
                continue
            newMapDict = mapDict.copy()
            if hasattr(newMapDict, 'preprocess'):
                newMapDict.preprocess(tileName=tileName, diagnosticsDir=diagnosticsDir)
            self.unfilteredMapsDictList.append(newMapDict)
        self.wcs = self.unfilteredMapsDictList[0]['wcs']
        self.shape = tuple(self.unfilteredMapsDictList[0]['data'].shape)

        self.flagMask = np.zeros(self.shape, dtype=int)
        for i, mapDict in enumerate(self.unfilteredMapsDictList):
            if mapDict.get('flagMask') is not None:
                self.flagMask = self.flagMask + (np.asarray(mapDict['flagMask']) * (i + 1))

        self.beamSolidAnglesDict = {}
        for mapDict in self.unfilteredMapsDictList:
            self.beamSolidAnglesDict[mapDict['obsFreqGHz']] = beamSolidAngle_nsr(mapDict)

        self.apodPix = 20
        for mapDict in self.unfilteredMapsDictList:
            if tuple(mapDict['data'].shape) != self.shape:
                raise Exception("Maps at different frequencies have different dimensions!")
        self.makeRadiansMap()
        self.signalNorm = 1.0
        self.fRelWeights = {}
        self._filtFullT = None      # [F, H, W] fp64 on the device
        self._filtHT = None         # [1, F, H, Wc] folded filter on the device
        self._filtHost = None
        self._dev = None

    # -- device staging ---------------------------------------------------------------------------------
    def _stage(self):
        """Uploads this tile's maps once: data / weights [1, F, H, W], masks u8 [1, H, W], median weights."""
        if self._dev is not None:
            return self._dev
        ops = self.ops
        t = ops.torch
        F = len(self.unfilteredMapsDictList)
        H, W = self.shape
        data = np.empty((1, F, H, W), dtype=ops.npdtype)
        weights = np.empty((1, F, H, W), dtype=ops.npdtype)
        for i, m in enumerate(self.unfilteredMapsDictList):
            data[0, i] = np.asarray(m['data'])
            weights[0, i] = np.asarray(m['weights']) if m.get('weights') is not None else 1.0
        m0 = self.unfilteredMapsDictList[0]
        dev = {'data': ops.to_device(data), 'weights': ops.to_device(weights),
               'surveyMask': ops.to_device(_maskToU8(m0.get('surveyMask'), self.shape)[None], t.uint8),
               'psMask': ops.to_device(_maskToU8(m0.get('pointSourceMask'), self.shape)[None], t.uint8)}
        dev['wmed'] = ops.median_weights(dev['weights'])
        self._dev = dev
        return dev

    @property
    def filt(self):
        """The filter in the reference's layout: float64 [F, H, W] on the full Fourier plane."""
        if self._filtHost is None and self._filtFullT is not None:
            self._filtHost = self._filtFullT.cpu().numpy()
        return self._filtHost

    @filt.setter
    def filt(self, value):
        self._filtHost = np.array(value, dtype=np.float64)
        self._filtFullT = self.ops.to_device(self._filtHost, self.ops.torch.float64)
        self._filtHT = self.ops.fold_filter(self._filtFullT).unsqueeze(0)

    def makeRadiansMap(self):
        """Map of distance in radians from the centre of the tile (filters.py:214-239)."""
        shape = self.shape
        x0 = int(shape[1] / 2)
        y0 = int(shape[0] / 2)
        ra0, dec0 = self.wcs.pix2wcs(x0, y0)
        ra1, dec1 = self.wcs.pix2wcs(x0 + 1, y0 + 1)
        self.degPerPixX = _wcs.calcAngSepDeg(ra0, dec0, ra1, dec0)
        self.degPerPixY = _wcs.calcAngSepDeg(ra0, dec0, ra0, dec1)
        # a function of the tile shape and the two central pixel scales only: tiles of one geometry (and the nested filter of a
        # real-space kernel build) share one read-only array
        key = (tuple(shape), float(self.degPerPixX), float(self.degPerPixY))
        cached = _RADIANS_MAPS.get(key)
        if cached is None:
            xRadRange = np.array([np.arange(int(np.floor(-shape[1] / 2)), int(shape[1] / 2), dtype=np.float64)
                                  * np.radians(self.degPerPixX)] * shape[0])
            yRadRange = np.array([np.arange(int(np.floor(-shape[0] / 2)), int(shape[0] / 2), dtype=np.float64)
                                  * np.radians(self.degPerPixY)] * shape[1]).transpose()
            cached = np.sqrt(xRadRange ** 2 + yRadRange ** 2)
            cached.setflags(write=False)
            if len(_RADIANS_MAPS) >= 8:
                _RADIANS_MAPS.pop(next(iter(_RADIANS_MAPS)))
            _RADIANS_MAPS[key] = cached
        self.radiansMap = cached

    def buildAndApply(self):
        raise Exception("Called a base filter class without a buildAndApply() function implemented.")

    def makeSignalTemplateMap(self, beam, amplitude=None):
        raise Exception("Called a base filter class without a makeSignalTemplateMap() function implemented.")

    def makeRealSpaceFilterProfile(self):
        """1-d real-space profile of the filter and its angular range in arcmin (filters.py:282-304)."""
        # The reference takes row H/2, columns W/2.. of fftshift(ifft2(filt).real) -- fftshift over ALL axes, the frequency
        # axis included.  That row is row 0 of the unshifted inverse transform, i.e. the 1-d inverse transform along x of the
        # mean over ky, so the full 2-d inverse transform is not needed.
        filt = np.asarray(self.filt)
        W = filt.shape[2]
        rowCut = fft.ifft(filt.mean(axis=1), axis=-1).real
        return realSpaceProfileFromRow(rowCut, W, self.degPerPixX)

    def makeNoiseMap(self, mapData):
        """Noise map from local 3-sigma-clipped measurements in grid cells (filters.py:345-483), on the GPU."""
        ops = self.ops
        dev = self._stage()
        t = device.asTensor(mapData, ops)
        if t.dim() == 2:
            t = t.unsqueeze(0)
        grid = engine.noiseGrid(self.shape, self.params['noiseParams'], self.wcs.getPixelSizeDeg())
        if isinstance(grid[0], str) and grid[0] == 'smart':
            rms, table, edges = ops.smart_rms(t.contiguous(), dev['wmed'], 1, grid[1])
        else:
            rms, table, edges = ops.cell_rms(t.contiguous(), dev['wmed'], 1, *grid)
        self.noiseTable = (table, edges)
        return rms[0].cpu().numpy()

    def loadFRelWeights(self):
        self._loadFilterFile()

    # -- filter cache -------------------------------------------------------------------------------------
    def _cachePath(self):
        if self.filterFileName is None:
            return None
        if os.path.exists(self.filterFileName):
            return self.filterFileName
        alt = self.filterFileName[:-5] + ".npz"
        if os.path.exists(alt):
            return alt
        return None

    def _saveFilterFile(self, arr, header):
        """The reference's filter cache (filters.py:767-777): primary HDU = the filter, SIGNORM / RW<n>_GHZ / RW<n> keywords."""
        from . import fitsio
        os.makedirs(os.path.dirname(self.filterFileName) or '.', exist_ok=True)
        fitsio.writeImage(self.filterFileName, np.asarray(arr), header)

    def _readFilterFile(self):
        from . import fitsio
        path = self._cachePath()
        if path.endswith(".npz"):           # caches written by earlier versions of this package
            with np.load(path) as z:
                return z['data'], {str(k): float(v) for k, v in zip(z['headerKeys'], z['headerVals'])}
        return fitsio.readImage(path)


# -------------------------------------------------------------------------------------------------------------
class MatchedFilter(MapFilter):
    """Multi-frequency matched filter in Fourier space (filters.py:519-859)."""

    def _bckSubSigmas(self):
        if 'bckSub' in self.params.keys() and 'bckSubScaleArcmin' in self.params.keys() and self.params['bckSub'] == True:
            return maps.smoothScalesPix(self.wcs, smoothScaleDeg=self.params['bckSubScaleArcmin'] / 60.)
        return None

    def buildAndApply(self, useCachedFilter=False):
        ops = self.ops
        dev = self._stage()
        F = len(self.unfilteredMapsDictList)
        H, W = self.shape
        Wc = W // 2 + 1
        fMapsToFilter = ops.rfft2(dev['data'].reshape(F, H, W), apodPix=self.apodPix).reshape(1, F, H, Wc)

        if self._cachePath() is None and useCachedFilter == False:
            method = self.params['noiseParams']['method']
            if method == 'dataMap':
                if 'noiseModelCatalog' in self.params.keys() and self.params['noiseModelCatalog'] is not None:
                    raise Exception("noiseModelCatalog needs maps.makeModelImage, which is outside the nemo_b200 hot path")
                fMapsForNoise = fMapsToFilter       # same maps, same apodisation (filters.py:528 vs :551)
            elif method in ['model', 'max(dataMap,CMB)']:
                raise Exception("noise method '%s' needs pixell (simCMBMap / modlmap) and is not implemented in nemo_b200" % (method))
            else:
                raise Exception("'%s' is not a valid filter noise method name - fix the .yml config file" % (method))
            w = engine.spectralWeights(self.unfilteredMapsDictList, self.params)
            templUnit = np.array([self.makeSignalTemplateMap(m['beamFileName']) for m in self.unfilteredMapsDictList])
            amps = engine.normAmplitudes(self.unfilteredMapsDictList, self.params)
            if amps is not None:
                templNorm = np.array([self.makeSignalTemplateMap(m['beamFileName'], amplitude=a)
                                      for m, a in zip(self.unfilteredMapsDictList, amps)])
                templNormT = ops.to_device(templNorm[None])
            else:
                templNormT = None
            cRADeg, cDecDeg = self.wcs.getCentreWCSCoords()
            cx, cy = self.wcs.wcs2pix(cRADeg, cDecDeg)
            res = engine.buildFourierFilters(ops, fMapsForNoise, ops.to_device(templUnit[None]), templNormT, w,
                                             self.params['outputUnits'], cy, cx, wantFull=True,
                                             bckSubSigmas=self._bckSubSigmas())
            self._filtHT = res['filtH']
            self._filtFullT = res['filtFull'][0]
            self._filtHost = None
            self._signalNormT = res['signalNorm']
            self.signalNorm = float(res['signalNorm'].cpu().numpy()[0])
            self.fRelWeights = {}
            if res['fRel'] is not None:
                planes, total = res['fRel']
                planes, total = planes.cpu().numpy()[0], float(total.cpu().numpy()[0])
                for v, mapDict in zip(planes, self.unfilteredMapsDictList):
                    self.fRelWeights[mapDict['obsFreqGHz']] = v / total
        else:
            print("... loading cached filter")
            self.loadFilter()
            self.params['saveRMSMap'] = False
            self.params['saveFilter'] = False
            self.params['savePlots'] = False

        filteredMap, zeroFlags = engine.applyFourierFilters(ops, fMapsToFilter, self._filtHT, 1, self._signalNormT, dev['psMask'],
                                                            bckSubSigmas=self._bckSubSigmas(), W=self.shape[1], zeroFlags=True)
        del fMapsToFilter

        if self.params['outputUnits'] == 'yc':
            mapUnits = 'yc'
            combinedObsFreqGHz = 'yc'
            beamSolidAngle_nsr = 0.0
        elif self.params['outputUnits'] == 'uK':
            combinedObsFreqGHz = float(list(self.beamSolidAnglesDict.keys())[0])
            mapUnits = 'uK'
            beamSolidAngle_nsr = self.beamSolidAnglesDict[combinedObsFreqGHz]
        else:
            raise Exception('need to specify "outputUnits" ("yc" or "uK") in filter params')

        pixDeg = self.wcs.getPixelSizeDeg()
        grid = engine.noiseGrid(self.shape, self.params['noiseParams'], pixDeg)
        trimPix = engine.trimSizePix(self.params, pixDeg, realSpace=False)
        data, sn, maskOut, rms, table, edges = engine.noiseSNAndMasks(ops, filteredMap, dev['wmed'], dev['psMask'],
                                                                      dev['surveyMask'], 1, grid, trimPix, self.apodPix, False,
                                                                      zeroFlags=zeroFlags)
        self.noiseTable = (table, edges)
        self.RMSMap = DeviceArray(rms[0])

        if 'saveRMSMap' in self.params and self.params['saveRMSMap'] == True and self.selFnDir is not None:
            self._saveRMSMap(rms[0])
        if 'saveFilter' in self.params and self.params['saveFilter'] == True and self.filterFileName is not None:
            header = {'SIGNORM': self.signalNorm}
            for count, key in enumerate(self.fRelWeights.keys(), start=1):
                header['RW%d_GHZ' % (count)] = key
                header['RW%d' % (count)] = self.fRelWeights[key]
            self._saveFilterFile(self.filt, header)

        return {'data': DeviceArray(data[0]), 'wcs': self.wcs, 'obsFreqGHz': combinedObsFreqGHz, 'SNMap': DeviceArray(sn[0]),
                'surveyMask': DeviceArray(maskOut[0]), 'flagMask': self.flagMask, 'mapUnits': mapUnits,
                'beamSolidAngle_nsr': beamSolidAngle_nsr, 'label': self.label, 'tileName': self.tileName}

    def _saveRMSMap(self, rmsT):
        """filters.py:760-762: selFnDir/<tile>/RMSMap_<label>#<tile>.fits with the tile's WCS (float32, as the reference's
        RICE_1 file holds; written uncompressed, see nemo_b200/fitsio.py)."""
        from . import fitsio
        outDir = self.selFnDir + os.path.sep + self.tileName
        os.makedirs(outDir, exist_ok=True)
        fitsio.saveFITS(outDir + os.path.sep + "RMSMap_%s#%s.fits" % (self.label, self.tileName), rmsT.cpu().numpy(), self.wcs,
                        compressionType="RICE_1")

    def loadFilter(self):
        """Loads a previously saved filter (filters.py:784-794)."""
        arr, header = self._readFilterFile()
        self.filt = arr
        self.signalNorm = float(header['SIGNORM'])
        self._signalNormT = self.ops.to_device(np.array([self.signalNorm]), self.ops.torch.float64)
        self.fRelWeights = {}
        for i in range(1, 10):
            if 'RW%d_GHZ' % (i) in header.keys():
                self.fRelWeights[header['RW%d_GHZ' % (i)]] = header['RW%d' % (i)]

    def _loadFilterFile(self):
        self.loadFilter()

    def reshapeFilter(self, shape):
        """The filter on the Fourier grid of a map of another shape with the same pixel scales (filters.py:797-821): linear
        interpolation in (l_y, l_x), l = 2 pi fftfreq(n, d = pixel size in radians) along each axis (pixell's enmap.laxes).

        The reference does this with scipy's interp1d (multipole -> fractional index of the old grid) and interp2d, which SciPy
        1.14 removed; its index map also interpolates across the jump between the last negative multipole and zero.  Here the
        multipole axes are put in ascending order first, so every output mode is interpolated between its two neighbours in l
        along each axis; beyond the old grid's largest |l| the edge value is used."""
        if len(shape) == 2:
            shape = [np.asarray(self.filt).shape[0], shape[0], shape[1]]
        assert (len(shape) == 3)
        filt = np.asarray(self.filt, dtype=np.float64)
        F, H, W = filt.shape

        def axisWeights(nOld, nNew, degPerPix):
            d = np.radians(degPerPix)
            lOld, lNew = 2 * np.pi * fft.fftfreq(nOld, d=d), 2 * np.pi * fft.fftfreq(nNew, d=d)
            order = np.argsort(lOld, kind='stable')
            pos = np.interp(lNew, lOld[order], np.arange(nOld, dtype=np.float64))       # clamps beyond the ends
            lo = np.minimum(np.floor(pos).astype(int), nOld - 2) if nOld > 1 else np.zeros(nNew, dtype=int)
            frac = pos - lo
            return order[lo], order[np.minimum(lo + 1, nOld - 1)], frac
        y0, y1, fy = axisWeights(H, shape[1], self.degPerPixY)
        x0, x1, fx = axisWeights(W, shape[2], self.degPerPixX)
        reshapedFilt = np.zeros(shape)
        for i in range(F):
            top = filt[i][y0][:, x0] * (1 - fx)[None, :] + filt[i][y0][:, x1] * fx[None, :]
            bottom = filt[i][y1][:, x0] * (1 - fx)[None, :] + filt[i][y1][:, x1] * fx[None, :]
            reshapedFilt[i] = top * (1 - fy)[:, None] + bottom * fy[:, None]
        return reshapedFilt

    def applyFilter(self, mapDataToFilter):
        """Applies the filter to a cube [F, H, W]: real maps are apodised and transformed first; complex input is
        taken as the full-plane FFT of real maps (filters.py:824-859).  Maps of another shape than the filter's get the filter
        interpolated onto their Fourier grid (reshapeFilter).  Returns a numpy array."""
        ops = self.ops
        F = len(self.unfilteredMapsDictList)
        H, W = self.shape
        if isinstance(mapDataToFilter, DeviceArray):
            mapDataToFilter = mapDataToFilter.tensor
        filtHT = self._filtHT
        if tuple(mapDataToFilter.shape) != (F, H, W):
            reshaped = self.reshapeFilter(tuple(mapDataToFilter.shape))
            H, W = reshaped.shape[1:]
            filtHT = ops.fold_filter(ops.to_device(reshaped[None], ops.torch.float64))
        isTensor = hasattr(mapDataToFilter, 'data_ptr')
        if not isTensor and 'complex' in np.asarray(mapDataToFilter).dtype.name:
            half = np.ascontiguousarray(np.asarray(mapDataToFilter)[:, :, :W // 2 + 1])
            fMaps = ops.to_device(half, ops.cdtype).reshape(1, F, H, W // 2 + 1)
        else:
            fMaps = ops.rfft2(device.asTensor(mapDataToFilter, ops).reshape(F, H, W).contiguous(),
                              apodPix=self.apodPix).reshape(1, F, H, W // 2 + 1)
        normT = ops.to_device(np.array([self.signalNorm], dtype=np.float64), ops.torch.float64)
        out = engine.applyFourierFilters(ops, fMaps, filtHT, 1, normT, None, bckSubSigmas=self._bckSubSigmas(), W=W)
        return out[0].cpu().numpy()


# -------------------------------------------------------------------------------------------------------------
class RealSpaceMatchedFilter(MapFilter):
    """Matched filter built in Fourier space on a noise region, truncated to a small real-space kernel and applied
    by direct convolution (filters.py:862-1218)."""

    def loadFilter(self):
        arr, header = self._readFilterFile()
        self.kern2d = np.array(arr, dtype=np.float64)
        self.signalNorm = float(header['SIGNORM'])
        self.bckSubScaleArcmin = header['BCKSCALE'] if 'BCKSCALE' in header.keys() else 0
        self.applyRACentre = header['APP_RA']
        self.applyDecCentre = header['APP_DEC']
        self.fRelWeights = {}
        for i in range(1, 10):
            if 'RW%d_GHZ' % (i) in header.keys():
                self.fRelWeights[header['RW%d_GHZ' % (i)]] = header['RW%d' % (i)]

    def _kernelMapsDictList(self, RADecSection):
        """The noise region the kernel is built from.  The reference re-reads it from disk with
        mapDict['RADecSection'] (filters.py:920-929); in-memory tiles are cut with the WCS instead."""
        out = []
        for mapDict in self.unfilteredMapsDictList:
            if RADecSection is None:
                out.append(mapDict.copy())
                continue
            RAMin, RAMax, decMin, decMax = RADecSection
            wcs = mapDict['wcs']
            xa, ya = wcs.wcs2pix(RAMax, decMin)
            xb, yb = wcs.wcs2pix(RAMin, decMax)
            x0, x1 = int(round(min(xa, xb))), int(round(max(xa, xb)))
            y0, y1 = int(round(min(ya, yb))), int(round(max(ya, yb)))
            H, W = self.shape
            x0, x1, y0, y1 = max(x0, 0), min(x1, W), max(y0, 0), min(y1, H)
            if not hasattr(wcs, 'crpix'):
                raise Exception("cutting a RADecSection needs a wcs with crpix/pixDeg (nemo_b200.wcs.CARWCS); "
                                "pass noiseParams['RADecSection'] = None to use the whole tile")
            newDict = mapDict.copy()
            for key in ['data', 'weights', 'surveyMask', 'pointSourceMask', 'flagMask']:
                if newDict.get(key) is not None:
                    newDict[key] = np.ascontiguousarray(np.asarray(newDict[key])[y0:y1, x0:x1])
            newDict['wcs'] = _wcs.CARWCS((y1 - y0, x1 - x0), wcs.pixDeg, wcs.crval, (wcs.crpix[0] - x0, wcs.crpix[1] - y0))
            out.append(newDict)
        return out

    def buildKernel(self, RADecSection, RADeg='centre', decDeg='centre'):
        """Builds the real-space kernel (filters.py:892-1041)."""
        if self._cachePath() is not None:
            return self.loadFilter()
        kernelUnfilteredMapsDictList = self._kernelMapsDictList(RADecSection)
        kernelLabel = "realSpaceKernel_%s" % (self.label)
        matchedFilterClass = self.params['noiseParams']['matchedFilterClass']
        if isinstance(matchedFilterClass, str):
            matchedFilterClass = globals()[matchedFilterClass]
        matchedFilter = matchedFilterClass(kernelLabel, kernelUnfilteredMapsDictList, self.params, tileName=self.tileName,
                                           diagnosticsDir=None, selFnDir=None)
        if RADecSection is None and len(matchedFilter.unfilteredMapsDictList) == len(self.unfilteredMapsDictList):
            matchedFilter._dev = self._stage()      # the noise region is the whole tile: the same arrays, already on the device
        matchedFilter.buildAndApply()
        self.matchedFilter = matchedFilter

        prof, arcminRange = matchedFilter.makeRealSpaceFilterProfile()
        self.kern2d, self.bckSubScaleArcmin = realSpaceKernelFromProfile(prof, arcminRange, matchedFilter.radiansMap, self.params,
                                                                         filt=matchedFilter.filt)

        signalMaps = []
        for mapDict in self.unfilteredMapsDictList:
            if self.params['outputUnits'] == 'yc':
                y0 = 2e-4
                if mapDict['obsFreqGHz'] is not None:
                    signalMap = self.makeSignalTemplateMap(mapDict['beamFileName'],
                                                           amplitude=maps.convertToDeltaT(y0, mapDict['obsFreqGHz']))
                else:
                    signalMap = self.makeSignalTemplateMap(mapDict['beamFileName'], amplitude=y0)
            elif self.params['outputUnits'] == 'uK':
                signalMap = self.makeSignalTemplateMap(mapDict['beamFileName'])
            else:
                raise Exception('need to specify "outputUnits" ("yc" or "uK") in filter params')
            signalMaps.append(signalMap)
        signalMaps = np.array(signalMaps)
        # applyFilter(signalMaps, calcFRelWeights=True) followed by .max() (filters.py:1029-1035), with the maximum, its position
        # and the per-frequency values there found on the device: a few numbers come back instead of F + 1 maps
        peak = self._filteredSignalPeak(signalMaps)
        if self.params['outputUnits'] == 'yc':
            self.signalNorm = y0 / peak
        else:
            self.signalNorm = 1.0 / peak

        if self.filterFileName is not None:
            header = {'SIGNORM': self.signalNorm, 'APP_RA': self.applyRACentre, 'APP_DEC': self.applyDecCentre}
            if self.params['bckSub'] == True:
                header['BCKSCALE'] = self.bckSubScaleArcmin
            for count, key in enumerate(self.fRelWeights.keys(), start=1):
                header['RW%d_GHZ' % (count)] = key
                header['RW%d' % (count)] = self.fRelWeights[key]
            self._saveFilterFile(self.kern2d, header)

    def buildAndApply(self, useCachedFilter=False):
        ops = self.ops
        dev = self._stage()
        if self._cachePath() is None:
            RAMin, RAMax, decMin, decMax = self.wcs.getImageMinMaxWCSCoords()
            section = self.params['noiseParams'].get('RADecSection')
            if section == 'tileNoiseRegions':
                RADecSection = [self.wcs.header['NRAMIN'], self.wcs.header['NRAMAX'],
                                self.wcs.header['NDEMIN'], self.wcs.header['NDEMAX']]
            elif section == 'auto':
                cRA, cDec = self.wcs.getCentreWCSCoords()
                halfSizeDeg = 2.0
                RADecSection = [cRA - halfSizeDeg / np.cos(np.radians(cDec)), cRA + halfSizeDeg / np.cos(np.radians(cDec)),
                                cDec - halfSizeDeg, cDec + halfSizeDeg]
            else:
                RADecSection = section
            self.applyDecCentre = (decMax + decMin) / 2.
            self.applyRACentre = (RAMax + RAMin) / 2.
            self.buildKernel(RADecSection, RADeg=self.applyRACentre, decDeg=self.applyDecCentre)
        else:
            self.loadFilter()

        # Apply the kernel to the *raw* tile (filters.py:1106-1110)
        raw, staged = [], True
        for mapDict in self.unfilteredMapsDictList:
            if hasattr(mapDict, 'loadTile'):
                raw.append(mapDict.loadTile('mapFileName', tileName=self.tileName))
                staged = False
            else:
                staged = staged and mapDict.get('rawData') is None
                raw.append(mapDict['rawData'] if mapDict.get('rawData') is not None else mapDict['data'])
        # in-memory tiles without a separate raw map: what was staged IS the raw tile
        rawT = dev['data'] if staged else ops.to_device(np.array(raw)[None])
        normT = ops.to_device(np.array([self.signalNorm], dtype=np.float64), ops.torch.float64)
        filteredMap = self._applyOnDevice(rawT, normT, dev['psMask'])

        if self.params['outputUnits'] == 'yc':
            mapUnits = 'yc'
            combinedObsFreqGHz = 'yc'
            beamSolidAngle_nsr = 0.0
        elif self.params['outputUnits'] == 'uK':
            if len(self.unfilteredMapsDictList) > 1:
                raise Exception("multi-frequency filtering not currently supported for outputUnits 'uK' (point source finding)")
            combinedObsFreqGHz = float(list(self.beamSolidAnglesDict.keys())[0])
            mapUnits = 'uK'
            beamSolidAngle_nsr = self.beamSolidAnglesDict[combinedObsFreqGHz]
        else:
            raise Exception('need to specify "outputUnits" ("yc" or "uK") in filter params')

        pixDeg = self.wcs.getPixelSizeDeg()
        grid = engine.noiseGrid(self.shape, self.params['noiseParams'], pixDeg)
        trimPix = engine.trimSizePix(self.params, pixDeg, realSpace=True)
        data, sn, maskOut, rms, table, edges = engine.noiseSNAndMasks(ops, filteredMap, dev['wmed'], dev['psMask'],
                                                                      dev['surveyMask'], 1, grid, trimPix, self.apodPix, True)
        self.noiseTable = (table, edges)
        self.RMSMap = DeviceArray(rms[0])
        return {'data': DeviceArray(data[0]), 'wcs': self.wcs, 'obsFreqGHz': combinedObsFreqGHz, 'SNMap': DeviceArray(sn[0]),
                'surveyMask': DeviceArray(maskOut[0]), 'flagMask': self.flagMask, 'mapUnits': mapUnits,
                'beamSolidAngle_nsr': beamSolidAngle_nsr, 'label': self.label, 'tileName': self.tileName}

    def _applyOnDevice(self, cubeT, normT, psMask, wantPlanes=False):
        """cubeT [B, F, H, W] -> [B, H, W]: optional Gaussian high-pass per frequency, kernel convolution, sum over
        frequency, normalisation, point-source mask (filters.py:1188-1216, 1113)."""
        ops = self.ops
        if self.params['bckSub'] == True and self.bckSubScaleArcmin > 0:
            sy, sx = maps.smoothScalesPix(self.wcs, RADeg=self.applyRACentre, decDeg=self.applyDecCentre,
                                          smoothScaleDeg=self.bckSubScaleArcmin / 60.)
            cubeT = ops.gauss_filter(cubeT.contiguous(), sy, sx, subtract=True)
        kernT = ops.to_device(self.kern2d, ops.torch.float64)
        return ops.conv2d(cubeT.contiguous(), kernT, norm=normT, mask=psMask, wantPlanes=wantPlanes)

    def _filteredSignalPeak(self, signalMaps):
        """max of the filtered signal templates (signalNorm still 1) and, as applyFilter(..., calcFRelWeights=True) does, the
        relative weight of every frequency plane at that pixel (filters.py:1204-1212)."""
        ops = self.ops
        cubeT = device.asTensor(signalMaps, ops)
        if cubeT.dim() == 3:
            cubeT = cubeT.unsqueeze(0)
        normT = ops.to_device(np.array([self.signalNorm], dtype=np.float64), ops.torch.float64)
        out, planes = self._applyOnDevice(cubeT, normT, None, wantPlanes=True)
        _, argmaxIdx, maxVal, _ = ops.template_norm(out.contiguous(), 0, 0.0, 0.0, 1.0)
        idx = int(device.toHost(argmaxIdx)[0])
        vals = device.toHost(planes[0].reshape(planes.shape[1], -1)[:, idx].contiguous())
        peak = float(device.toHost(maxVal)[0])
        self.fRelWeights = {}
        totalSignal = vals.sum()
        for v, mapDict in zip(vals, self.unfilteredMapsDictList):
            self.fRelWeights[mapDict['obsFreqGHz']] = v / totalSignal
        return peak

    def applyFilter(self, mapDataToFilter, calcFRelWeights=False):
        """Applies the kernel to a cube [F, H, W] (filters.py:1172-1218).  Returns a numpy array."""
        ops = self.ops
        cubeT = device.asTensor(mapDataToFilter, ops)
        if cubeT.dim() == 3:
            cubeT = cubeT.unsqueeze(0)
        normT = ops.to_device(np.array([self.signalNorm], dtype=np.float64), ops.torch.float64)
        if calcFRelWeights == True:
            out, planes = self._applyOnDevice(cubeT, normT, None, wantPlanes=True)
            planes = planes[0].cpu().numpy()
            self.fRelWeights = {}
            maxIndex = np.argmax(planes.sum(axis=0))
            totalSignal = planes.sum(axis=0).flatten()[maxIndex]
            for plane, mapDict in zip(planes, self.unfilteredMapsDictList):
                self.fRelWeights[mapDict['obsFreqGHz']] = plane.flatten()[maxIndex] / totalSignal
        else:
            out = self._applyOnDevice(cubeT, normT, None)
        return out[0].cpu().numpy()


# -------------------------------------------------------------------------------------------------------------
# Pieces of the real-space kernel construction shared by the class API (one tile at a time) and the batched driver
# (nemo_b200.realspace): everything after the Fourier filter is small host work.
def realSpaceProfileFromRow(rowCut, W, degPerPixX):
    """(prof, arcminRange) of MapFilter.makeRealSpaceFilterProfile (filters.py:282-304) from rowCut [F, W] = row 0 of
    ifft2(filt).real, i.e. the 1-d inverse transform along x of the ky-mean of the filter.  The reference takes row H/2,
    columns W/2.. of fftshift(...) over ALL axes: the frequency axis is shifted too."""
    rowCut = fft.fftshift(rowCut, axes=0)
    cut = rowCut[:, :W - W // 2]
    normFactor = abs(cut).max()
    prof = cut / normFactor
    arcminRange = np.arange(0, prof.shape[1]) * degPerPixX * 60.0
    return prof, arcminRange


_KERNEL_WINDOWS = {}


def _interpProfile(r, xp, fp):
    """interpolate.interp1d(xp, fp, bounds_error=False, fill_value=0.0)(r) for the linear default: scipy evaluates it with
    np.interp and fills outside [xp[0], xp[-1]] (filters.py:957-958)."""
    out = np.interp(r, xp, fp)
    out[(r < xp[0]) | (r > xp[-1])] = 0.0
    return out


def _kernelWindow(rmap, rRadians, mask, rIndex, kernelMaxArcmin, arcminRange):
    """The window of the radians map around the tile centre that can hold a symmetrised kernel (a function of the geometry
    only: kept per radians map)."""
    key = (id(rmap), rmap.shape, float(kernelMaxArcmin), len(arcminRange), float(arcminRange[1] if len(arcminRange) > 1 else 0.0))
    win = _KERNEL_WINDOWS.get(key)
    if win is None or win[0] is not rmap:
        yc, xc = int(np.argmin(rmap[:, 0])), int(np.argmin(rmap[0, :]))
        rMax = rRadians[mask].max() if mask.any() else 0.0
        rows, cols = np.flatnonzero(rmap[:, xc] <= rMax), np.flatnonzero(rmap[yc, :] <= rMax)
        pad = rIndex + 2
        wy0, wy1 = max(rows.min() - pad, 0), min(rows.max() + pad + 1, rmap.shape[0])
        wx0, wx1 = max(cols.min() - pad, 0), min(cols.max() + pad + 1, rmap.shape[1])
        if len(_KERNEL_WINDOWS) >= 16:
            _KERNEL_WINDOWS.pop(next(iter(_KERNEL_WINDOWS)))
        win = _KERNEL_WINDOWS[key] = (rmap, np.ascontiguousarray(rmap[wy0:wy1, wx0:wx1]))
    return win[1]


def realSpaceKernelsFromRows(rowCuts, W, degPerPixX, rmap, params):
    """realSpaceProfileFromRow + realSpaceKernelFromProfile for n filters at once (rowCuts [n, F, W]; symmetrised kernels with
    a common kernelMaxArcmin and a given bckSubScaleArcmin, the batched driver's case): the same arithmetic written over
    arrays -- the linear interpolation as np.interp evaluates it, slope * (x - xp[j]) + fp[j].  Returns kern2d [n, F, kh, kw];
    kernels whose peak does not sit where the first one's does are made one by one."""
    n, F, _ = rowCuts.shape
    rc = np.roll(rowCuts, F // 2, axis=1)                       # fftshift over ALL axes shifts the frequency axis too
    cut = rc[:, :, :W - W // 2]
    prof = cut / np.abs(cut).reshape(n, -1).max(axis=1)[:, None, None]
    arcminRange = np.arange(0, cut.shape[2]) * degPerPixX * 60.0
    kernelMaxArcmin = params['noiseParams']['kernelMaxArcmin']
    rIndex = np.where(arcminRange > kernelMaxArcmin)[0][0]
    mask = np.less(arcminRange, kernelMaxArcmin)
    rRadians = np.radians(arcminRange / 60.)
    rwin = _kernelWindow(rmap, rRadians, mask, rIndex, kernelMaxArcmin, arcminRange)
    xp = rRadians[mask]
    one = lambda o: realSpaceKernelFromProfile(prof[o], arcminRange, rmap, params)[0]
    if len(xp) < 2:
        return np.array([one(o) for o in range(n)])
    fp = np.ascontiguousarray(prof[:, :, mask])
    slope = np.diff(fp, axis=2) / np.diff(xp)

    def interp(sel, r):
        """the interpolated profiles of filters `sel` at the radii r [h, w] -> [len(sel), F, h, w]"""
        j = np.clip(np.searchsorted(xp, r, side='right') - 1, 0, len(xp) - 2)
        v = slope[sel][:, :, j] * (r - xp[j]) + fp[sel][:, :, j]
        v[:, :, r == xp[-1]] = fp[sel][:, :, -1][:, :, None]
        v[:, :, (r < xp[0]) | (r > xp[-1])] = 0.0
        return v

    # The interpolated profile is piecewise linear in r, so its largest absolute value over the window is the largest |knot
    # value|.  When that is the value at r = 0, strictly, the peak is the centre pixel (the only one at r = 0) and the kernel is
    # the block of +- rIndex pixels around it: only that block is evaluated -- a fifth of the window.
    out = [None] * n
    cyx = np.argwhere(rwin == 0.0)
    if len(cyx) == 1 and xp[0] == 0.0:
        cy, cx = int(cyx[0][0]), int(cyx[0][1])
        yMin, yMax, xMin, xMax = cy - rIndex, cy + rIndex, cx - rIndex, cx + rIndex
        if yMin >= 0 and xMin >= 0 and yMax <= rwin.shape[0] and xMax <= rwin.shape[1]:
            if (yMax - yMin) % 2 == 0:
                yMin = yMin + 1
            if (xMax - xMin) % 2 == 0:
                xMin = xMin + 1
            afp = np.abs(fp)
            central = np.flatnonzero(afp[:, :, 0].max(axis=1) > afp[:, :, 1:].reshape(n, -1).max(axis=1))
            if len(central) > 0:
                block = interp(central, rwin[yMin:yMax, xMin:xMax])
                for i, o in enumerate(central):
                    out[o] = block[i]
    rest = np.array([o for o in range(n) if out[o] is None], dtype=int)
    if len(rest) == 0:
        return np.ascontiguousarray(np.array(out))
    profile2d = interp(rest, rwin)                                              # [len(rest), F, hw, ww]
    n = len(rest)
    ap = np.abs(profile2d).reshape(n, -1)
    amax = ap.max(axis=1)
    first = (ap == amax[:, None]).argmax(axis=1)
    z, y, x = np.unravel_index(first, profile2d.shape[1:])
    yMin, yMax, xMin, xMax = y - rIndex, y + rIndex, x - rIndex, x + rIndex
    ok = (amax > 0) & (yMin >= 0) & (xMin >= 0) & (yMax <= profile2d.shape[2]) & (xMax <= profile2d.shape[3])
    yMin = np.where((yMax - yMin) % 2 == 0, yMin + 1, yMin)
    xMin = np.where((xMax - xMin) % 2 == 0, xMin + 1, xMin)
    for i, o in enumerate(rest):
        out[o] = profile2d[i, :, yMin[i]:yMax[i], xMin[i]:xMax[i]] if ok[i] else one(o)
    return np.ascontiguousarray(np.array(out))


def realSpaceKernelFromProfile(prof, arcminRange, rmap, params, filt=None):
    """(kern2d [F, kh, kw], bckSubScaleArcmin) of RealSpaceMatchedFilter.buildKernel (filters.py:944-985) from the 1-d profile.
    rmap: the radians map of the region the filter was built on; filt: the full Fourier filter, needed only without
    'symmetrize' (the 2-d kernel is then cut out of its inverse transform)."""
    kernelMaxArcmin = params['noiseParams']['kernelMaxArcmin']
    rIndex = np.where(arcminRange > kernelMaxArcmin)[0][0]
    mask = np.less(arcminRange, kernelMaxArcmin)
    if 'symmetrize' in params['noiseParams'].keys() and params['noiseParams']['symmetrize'] == True:
        rRadians = np.radians(arcminRange / 60.)
        # The interpolated profile is zero beyond the last radius below kernelMaxArcmin, so only a window around the tile
        # centre can hold the peak and the kernel: evaluating it there gives the same kern2d as the reference's full-map
        # evaluation (filters.py:955-960) for a hundredth of the work.  The window depends on the geometry only: kept.
        rwin = _kernelWindow(rmap, rRadians, mask, rIndex, kernelMaxArcmin, arcminRange)
        xp = rRadians[mask]
        profile2d = np.array([_interpProfile(rwin, xp, prof[i, mask]) for i in range(prof.shape[0])])
        amax = abs(profile2d).max()
        z, y, x = np.unravel_index(int(np.argmax(abs(profile2d) == amax)), profile2d.shape)
        if not (amax > 0 and y - rIndex >= 0 and x - rIndex >= 0 and y + rIndex <= profile2d.shape[1]
                and x + rIndex <= profile2d.shape[2]):
            # degenerate profile or a kernel that leaves the window: the full map, as the reference
            profile2d = np.array([_interpProfile(rmap, xp, prof[i, mask]) for i in range(prof.shape[0])])
    else:
        if filt is None:
            raise Exception("a real-space kernel without 'symmetrize' is cut out of the filter's 2-d inverse transform: filt needed")
        profile2d = fft.ifft2(np.asarray(filt)).real
        profile2d = fft.fftshift(profile2d)
    # first position (in z, y, x order) of the largest absolute value: np.where(abs(profile2d) == abs(profile2d).max())[..][0]
    ap = abs(profile2d)
    z, y, x = np.unravel_index(int(np.argmax(ap == ap.max())), profile2d.shape)
    yMin, yMax, xMin, xMax = y - rIndex, y + rIndex, x - rIndex, x + rIndex
    if (yMax - yMin) % 2 == 0:
        yMin = yMin + 1
    if (xMax - xMin) % 2 == 0:
        xMin = xMin + 1
    kern2d = np.ascontiguousarray(profile2d[:, yMin:yMax, xMin:xMax])
    if 'bckSubScaleArcmin' in params.keys():
        bckSubScaleArcmin = params['bckSubScaleArcmin']
    else:
        func = np.min if np.greater(prof[0, 0], 0) == True else np.max
        bckSubScaleArcmin = arcminRange[prof[0] == func(prof[0])][0]
    return kern2d, bckSubScaleArcmin

# -------------------------------------------------------------------------------------------------------------
# Signal templates are host-made inputs of the hot path and depend only on the tile geometry (shape and central pixel
# scales), the beam and the model parameters -- not on the map.  The reference paints them afresh for every tile and
# filter (~0.1 s each with the beam convolution); tiles that share a geometry reuse them here.  Entries are read-only.
_TEMPLATE_CACHE = {}
_TEMPLATE_CACHE_MAX = 64


def _beamKey(beam):
    if isinstance(beam, str):
        return ('file', beam, os.path.getmtime(beam) if os.path.exists(beam) else None)
    import hashlib
    return ('profile', hashlib.sha1(np.ascontiguousarray(beam.rDeg).tobytes() + np.ascontiguousarray(beam.profile1d).tobytes()).hexdigest())


def _cachedTemplate(key, make):
    hit = _TEMPLATE_CACHE.get(key)
    if hit is None:
        hit = np.asarray(make())
        hit.setflags(write=False)
        if len(_TEMPLATE_CACHE) >= _TEMPLATE_CACHE_MAX:
            _TEMPLATE_CACHE.pop(next(iter(_TEMPLATE_CACHE)))
        _TEMPLATE_CACHE[key] = hit
    return hit


class BeamFilter(MapFilter):
    """Templates from ACT-style beam profiles (filters.py:1222-1233)."""

    def makeSignalTemplateMap(self, beamFileName, amplitude=None):
        key = ('beam', tuple(self.shape), self.degPerPixX, self.degPerPixY, _beamKey(beamFileName), amplitude)
        return _cachedTemplate(key, lambda: signals.makeBeamModelSignalMap(np.degrees(self.radiansMap), self.wcs, beamFileName,
                                                                           amplitude=amplitude))


class ArnaudModelFilter(MapFilter):
    """Templates from the Arnaud et al. (2010) GNFW profile (filters.py:1236-1252)."""

    def makeSignalTemplateMap(self, beamFileName, amplitude=None):
        gnfw = self.params.get('GNFWParams', 'default')
        # the reference paints the cluster at wcs.getCentreWCSCoords() (filters.py:1244-1250), which is where the normalisation
        # spline is evaluated (:655-662); that is pixel (W/2, H/2) of radiansMap for a WCS whose centre is an integer pixel and
        # half a pixel away from it for e.g. astWCS's (NAXIS/2 in FITS counting) convention
        cRADeg, cDecDeg = self.wcs.getCentreWCSCoords()
        cx, cy = self.wcs.wcs2pix(cRADeg, cDecDeg)
        key = ('arnaud', self.params['z'], self.params['M500MSun'], repr(gnfw), tuple(self.shape), self.degPerPixX, self.degPerPixY,
               float(cx), float(cy), _beamKey(beamFileName), amplitude)

        def make():
            beamDegreesMap = np.degrees(self.radiansMap)
            if float(cx) == int(self.shape[1] / 2) and float(cy) == int(self.shape[0] / 2):
                degreesMap = beamDegreesMap
            else:
                xDeg = (np.arange(self.shape[1], dtype=np.float64) - cx) * self.degPerPixX
                yDeg = (np.arange(self.shape[0], dtype=np.float64) - cy) * self.degPerPixY
                degreesMap = np.sqrt(xDeg[None, :] ** 2 + yDeg[:, None] ** 2)
            return signals.makeArnaudModelSignalMap(self.params['z'], self.params['M500MSun'], degreesMap, self.wcs, beam=beamFileName,
                                                    GNFWParams=gnfw, amplitude=amplitude, convolveWithBeam=True,
                                                    beamDegreesMap=beamDegreesMap)
        return _cachedTemplate(key, make)


class BattagliaModelFilter(MapFilter):
    """Battaglia et al. (2012) templates need pyccl for the redshift evolution (filters.py:1255-1277)."""

    def makeSignalTemplateMap(self, beamFileName, amplitude=None):
        raise Exception("BattagliaModelFilter templates need pyccl, which nemo_b200 does not depend on; "
                        "use ArnaudModelFilter or pass template maps directly to nemo_b200.pipelines")


class ArnaudModelMatchedFilter(MatchedFilter, ArnaudModelFilter):
    pass


class BattagliaModelMatchedFilter(MatchedFilter, BattagliaModelFilter):
    pass


class BeamMatchedFilter(MatchedFilter, BeamFilter):
    pass


class ArnaudModelRealSpaceMatchedFilter(RealSpaceMatchedFilter, ArnaudModelFilter):
    pass


class BattagliaModelRealSpaceMatchedFilter(RealSpaceMatchedFilter, BattagliaModelFilter):
    pass


class BeamRealSpaceMatchedFilter(RealSpaceMatchedFilter, BeamFilter):
    pass
