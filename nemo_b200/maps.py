"""
Map helpers that sit on the hot path (reference nemo/maps.py): unit conversions (:660-688) are host scalars;
``smoothMap`` / ``subtractBackground`` (:1355-1368, :1433-1458) run on the GPU through nb_gauss_filter.
"""

import numpy as np

from . import signals
from . import wcs as _wcs


def convertToY(mapData, obsFrequencyGHz=148):
    fx = signals.fSZ(obsFrequencyGHz)
    return (mapData / (signals.TCMB * 1e6)) / fx


def convertToDeltaT(mapData, obsFrequencyGHz=148, TCMBAlpha=0.0, z=None):
    fx = signals.fSZ(obsFrequencyGHz, TCMBAlpha=TCMBAlpha, z=z)
    return mapData * fx * (signals.TCMB * 1e6)


def smoothScalesPix(wcs, RADeg='centre', decDeg='centre', smoothScaleDeg=5.0 / 60.0):
    """(sigmaY, sigmaX) in pixels of the Gaussian maps.smoothMap applies."""
    xPixScale, yPixScale = _wcs.pixelScalesDeg(wcs, RADeg, decDeg)
    return smoothScaleDeg / yPixScale, smoothScaleDeg / xPixScale


def smoothMap(data, wcs, RADeg='centre', decDeg='centre', smoothScaleDeg=5.0 / 60.0, ops=None):
    """Gaussian smoothing on the GPU.  `data` may be a numpy array (returned as numpy) or a device tensor."""
    from . import device
    return device.gaussFilter(data, *smoothScalesPix(wcs, RADeg, decDeg, smoothScaleDeg), subtract=False, ops=ops)


def subtractBackground(data, wcs, RADeg='centre', decDeg='centre', smoothScaleDeg=30.0 / 60.0, ops=None):
    from . import device
    return device.gaussFilter(data, *smoothScalesPix(wcs, RADeg, decDeg, smoothScaleDeg), subtract=True, ops=ops)
