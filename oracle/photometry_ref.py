"""
numpy/scipy restatement of /root/reference/nemo/photometry.py:25-351.  TEST INFRASTRUCTURE -- see
oracle/__init__.py.  Catalogs are lists of dicts (what the reference hands to
``catalogs.catalogListToTab``, catalogs.py:443-468); astropy is absent so no Table is built.

  getObjectPositions   photometry.py:193-222
  findObjects          photometry.py:25-190   (incl. the measureShapes ellipses, :126-178)
  getSNRValues         photometry.py:225-255
  measureFluxes        photometry.py:258-351
"""

import numpy as np
from scipy import ndimage, interpolate

from . import signals_ref as signals
from .wcs_stub import j2000ToGalactic


def _makeRA(myRADeg):
    """catalogs.py:295-313."""
    hours = (myRADeg / 360) * 24
    strHours = ("%.10f" % (hours))
    if hours < 10:
        sHours = "0" + strHours[0]
    else:
        sHours = strHours[:2]
    mins = float(strHours[strHours.index("."):]) * 60
    strMins = ("%.10f" % (mins))
    if mins < 10:
        sMins = "0" + strMins[:3]
    else:
        sMins = strMins[:4]
    return (sHours + sMins)


def _makeDec(myDecDeg):
    """catalogs.py:316-346."""
    text = str(float(myDecDeg))
    mins = float(text[text.index("."):]) * 60
    if mins < 10:
        sMins = "0" + str(mins)[:1]
    else:
        sMins = str(mins)[:2]
    if myDecDeg > 0:
        if myDecDeg < 10:
            sDeg = "0" + text[0]
        else:
            sDeg = text[:2]
        return "+" + sDeg + sMins
    if myDecDeg > -10:
        sDeg = "-0" + text[1]
    else:
        sDeg = text[:3]
    return str(sDeg + sMins)


def makeName(RADeg, decDeg, prefix='ACT-CL'):
    """catalogs.py:259-274: prefix JHHMM.m+/-DDMM, digits cut from the printed values."""
    return prefix + " J" + _makeRA(float(RADeg)) + _makeDec(float(decDeg))


def getObjectPositions(mapData, threshold, findCenterOfMass=True):
    if threshold < 0:
        raise Exception("Detection threshold (thresholdSigma in the config file) cannot be negative unless in forced photometry mode.")
    sigPix = np.array(np.greater(mapData, threshold), dtype=int)
    sigPixMask = np.equal(sigPix, 1)
    segmentationMap, numObjects = ndimage.label(sigPix)
    objIDs = np.unique(segmentationMap)
    if findCenterOfMass == True:
        objPositions = ndimage.center_of_mass(mapData, labels=segmentationMap, index=objIDs)
    else:
        objPositions = ndimage.maximum_position(mapData, labels=segmentationMap, index=objIDs)
    objNumPix = ndimage.sum(sigPixMask, labels=segmentationMap, index=objIDs)
    return objIDs, objPositions, objNumPix, segmentationMap


def findObjects(filteredMapDict, threshold=3.0, minObjPix=3, rejectBorder=10, findCenterOfMass=True,
                removeRings=True, ringThresholdSigma=0, invertMap=False, objIdent='ACT-CL', longNames=False,
                verbose=True, useInterpolator=True, measureShapes=False, DS9RegionsPath=None):
    if rejectBorder is None:
        rejectBorder = 0
    data = filteredMapDict['SNMap']
    wcs = filteredMapDict['wcs']
    flagMask = filteredMapDict['flagMask']
    if invertMap == True:
        data = data * -1
    objIDs, objPositions, objNumPix, segMap = getObjectPositions(data, threshold, findCenterOfMass=findCenterOfMass)
    if removeRings == True:
        minRingPix = 30
        ringIDs, ringPositions, ringNumPix, ringSegMap = getObjectPositions(data, ringThresholdSigma,
                                                                           findCenterOfMass=True)
        for i in range(len(ringIDs)):
            if type(ringNumPix) != float and ringNumPix[i] > minRingPix:
                y, x = ringPositions[i]
                if ringSegMap[int(y), int(x)] != ringIDs[i]:
                    ringSegMap[np.equal(ringSegMap, ringIDs[i])] = -1 * ringSegMap[np.equal(ringSegMap, ringIDs[i])]
        ringMask = np.array(np.less(ringSegMap, 0), dtype=int)
    else:
        ringMask = None
    if useInterpolator == True:
        mapInterpolator = interpolate.RectBivariateSpline(np.arange(data.shape[0]), np.arange(data.shape[1]),
                                                          data, kx=3, ky=3)
    catalog = []
    idNumCount = 1
    for i in range(len(objIDs)):
        if type(objNumPix) != float and objNumPix[i] > minObjPix:
            objDict = {}
            objDict['id'] = idNumCount
            objDict['x'] = objPositions[i][1]
            objDict['y'] = objPositions[i][0]
            if ringMask is not None:
                if ringMask[int(objDict['y']), int(objDict['x'])] > 0:
                    continue
            objDict['RADeg'], objDict['decDeg'] = wcs.pix2wcs(objDict['x'], objDict['y'])
            if objDict['RADeg'] < 0:
                objDict['RADeg'] = 360.0 + objDict['RADeg']
            galLong, galLat = j2000ToGalactic(objDict['RADeg'], objDict['decDeg'])
            objDict['galacticLatDeg'] = galLat
            objDict['name'] = makeName(objDict['RADeg'], objDict['decDeg'], prefix=objIdent)
            objDict['numSigPix'] = objNumPix[i]
            objDict['template'] = filteredMapDict['label']
            objDict['tileName'] = filteredMapDict['tileName']
            if useInterpolator == True:
                objDict['SNR'] = mapInterpolator(objDict['y'], objDict['x'])[0][0]
            else:
                objDict['SNR'] = data[int(round(objDict['y'])), int(round(objDict['x']))]
            objDict['flags'] = flagMask[int(round(objDict['y'])), int(round(objDict['x']))]
            if measureShapes == True:
                # SExtractor-style shape measurements (photometry.py:126-178)
                doubleCheck = False
                if objDict['numSigPix'] > 9:
                    mask = np.equal(segMap, objIDs[i])
                    ys, xs = np.where(segMap == objIDs[i])
                    yMin, xMin = ys.min(), xs.min()
                    xs = xs - xMin
                    ys = ys - yMin
                    cx2 = (xs * data[mask]).sum() / data[mask].sum()
                    cy2 = (ys * data[mask]).sum() / data[mask].sum()
                    x2 = (((xs ** 2) * data[mask]).sum() / data[mask].sum()) - cx2 ** 2
                    y2 = (((ys ** 2) * data[mask]).sum() / data[mask].sum()) - cy2 ** 2
                    xy = (((xs * ys) * data[mask]).sum() / data[mask].sum()) - cx2 * cy2
                    with np.errstate(divide='ignore', invalid='ignore'):
                        theta = np.degrees(np.arctan(2 * (xy / (x2 - y2))) / 2.0)
                    if xy > 0 and theta < 0:
                        theta = theta + 90
                    elif xy < 0 and theta > 0:
                        theta = theta - 90
                    if theta > 0 and xy > 0:
                        doubleCheck = True
                    if theta < 0 and xy < 0:
                        doubleCheck = True
                    if doubleCheck == True:
                        A = np.sqrt((x2 + y2) / 2.0 + np.sqrt(((x2 - y2) / 2) ** 2 + xy ** 2))
                        B = np.sqrt((x2 + y2) / 2.0 - np.sqrt(((x2 - y2) / 2) ** 2 + xy ** 2))
                        segArea = float(np.count_nonzero(np.equal(segMap, objIDs[i])))
                        curArea = A * B * np.pi
                        scaleFactor = np.sqrt(segArea / curArea)
                        A = A * scaleFactor
                        B = B * scaleFactor
                        ecc = np.sqrt(1 - B ** 2 / A ** 2)
                        objDict['ellipse_PA'] = theta
                        objDict['ellipse_A'] = A
                        objDict['ellipse_B'] = B
                        objDict['ellipse_x0'] = cx2 + xMin
                        objDict['ellipse_y0'] = cy2 + yMin
                        objDict['ellipse_e'] = ecc
                if objDict['numSigPix'] <= 9 or doubleCheck == False:
                    for key in ('ellipse_PA', 'ellipse_A', 'ellipse_B', 'ellipse_x0', 'ellipse_y0', 'ellipse_e'):
                        objDict[key] = -99
            if objDict['SNR'] > threshold:
                catalog.append(objDict)
            idNumCount = idNumCount + 1
    return catalog


def getSNRValues(catalog, SNMap, wcs, useInterpolator=True, invertMap=False, prefix=''):
    if invertMap == True:
        SNMap = SNMap * -1
    if useInterpolator == True:
        mapInterpolator = interpolate.RectBivariateSpline(np.arange(SNMap.shape[0]), np.arange(SNMap.shape[1]),
                                                          SNMap, kx=3, ky=3)
    for obj in catalog:
        obj[prefix + 'SNR'] = 0.0
    for obj in catalog:
        x, y = wcs.wcs2pix(obj['RADeg'], obj['decDeg'])
        if int(x) > 0 and int(x) < SNMap.shape[1] and int(y) > 0 and int(y) < SNMap.shape[0]:
            if useInterpolator == True:
                obj[prefix + 'SNR'] = mapInterpolator(y, x)[0][0]
            else:
                obj[prefix + 'SNR'] = SNMap[int(round(y)), int(round(x))]


def measureFluxes(catalog, filteredMapDict, diagnosticsDir=None, photFilteredMapDict=None,
                  useInterpolator=True, ycObsFreqGHz=148.0):
    mapData = filteredMapDict['data']
    wcs = filteredMapDict['wcs']
    mapUnits = wcs.header['BUNIT']
    if photFilteredMapDict is not None:
        getSNRValues(catalog, photFilteredMapDict['SNMap'], wcs, prefix='fixed_', useInterpolator=useInterpolator)
    if 'BEAMNSR' in wcs.header.keys():
        beamSolidAngle_nsr = wcs.header['BEAMNSR']
        obsFreqGHz = wcs.header['FREQGHZ']
        reportJyFluxes = True
    else:
        reportJyFluxes = False

    def _interp(m):
        if useInterpolator == True:
            return interpolate.RectBivariateSpline(np.arange(m.shape[0]), np.arange(m.shape[1]), m, kx=3, ky=3)
        return None

    mapDataList = [mapData]
    interpolatorList = [_interp(mapData)]
    prefixList = ['']
    if photFilteredMapDict is not None:
        mapDataList.append(photFilteredMapDict['data'])
        prefixList.append('fixed_')
        interpolatorList.append(_interp(photFilteredMapDict['data']))
    keysToAdd = ['deltaT_c', 'err_deltaT_c']
    if mapUnits == 'yc':
        keysToAdd = keysToAdd + ['y_c', 'err_y_c']
    elif mapUnits == 'uK':
        keysToAdd = keysToAdd + ['fluxJy', 'err_fluxJy']
    for prefix in prefixList:
        for k in keysToAdd:
            for obj in catalog:
                obj[prefix + k] = 0.0
    for obj in catalog:
        x, y = wcs.wcs2pix(obj['RADeg'], obj['decDeg'])
        for data, prefix, interpolator in zip(mapDataList, prefixList, interpolatorList):
            if useInterpolator == True:
                mapValue = interpolator(y, x)[0][0]
            else:
                mapValue = data[int(round(y)), int(round(x))]
            if mapUnits == 'yc':
                yc = mapValue
                obj[prefix + 'y_c'] = yc / 1e-4
                obj[prefix + 'err_y_c'] = obj[prefix + 'y_c'] / obj[prefix + 'SNR']
                deltaTc = signals.convertToDeltaT(yc, obsFrequencyGHz=ycObsFreqGHz)
                obj[prefix + 'deltaT_c'] = deltaTc
                obj[prefix + 'err_deltaT_c'] = abs(deltaTc / obj[prefix + 'SNR'])
            elif mapUnits == 'uK':
                deltaTc = mapValue
                obj[prefix + 'deltaT_c'] = deltaTc
                obj[prefix + 'err_deltaT_c'] = deltaTc / obj[prefix + 'SNR']
                if reportJyFluxes == True:
                    obj[prefix + "fluxJy"] = signals.deltaTToJyPerSr(obj[prefix + 'deltaT_c'], obsFreqGHz) * beamSolidAngle_nsr * 1.e-9
                    obj[prefix + "err_fluxJy"] = signals.deltaTToJyPerSr(obj[prefix + 'err_deltaT_c'], obsFreqGHz) * beamSolidAngle_nsr * 1.e-9
