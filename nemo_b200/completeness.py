"""
Noise level versus area tables -- the part of the reference's ``nemo.completeness`` that follows the filtering of a tile
directly (``getRMSTab``, nemo/completeness.py:949-1005, called from pipelines.py:395-401 and gathered over ranks at :405):
the RMS map of the reference filter compressed to (y0RMS, areaDeg2) rows.  The reference does it with ``np.unique`` over the
map and one boolean sum per level; here every pixel looks its own level up among the values the noise step painted
(``nb_rms_area_table``), so the map is read once, on the device, while it is still there.
"""

import numpy as np

from . import catalogs
from . import device
from . import wcs as _wcs


def rowAreasDeg2(shape, wcs):
    """Pixel area of every row in deg^2: what maps.getPixelAreaArcmin2Map computes (nemo/maps.py:1461-1482), / 3600."""
    H = int(shape[0])
    RACentre, decCentre = wcs.getCentreWCSCoords()
    x0, _ = wcs.wcs2pix(RACentre, decCentre)
    ys = np.arange(H, dtype=np.float64)
    try:
        ra0, dec0 = wcs.pix2wcs(np.full(H, x0), ys)
        ra1, dec1 = wcs.pix2wcs(np.full(H, x0 + 1), ys + 1)
        ra0, dec0, ra1, dec1 = (np.asarray(v, dtype=np.float64) for v in (ra0, dec0, ra1, dec1))
        if ra0.shape != (H,):
            raise ValueError
    except Exception:
        pts0 = [wcs.pix2wcs(x0, y) for y in range(H)]
        pts1 = [wcs.pix2wcs(x0 + 1, y + 1) for y in range(H)]
        ra0, dec0 = np.array([p[0] for p in pts0]), np.array([p[1] for p in pts0])
        ra1, dec1 = np.array([p[0] for p in pts1]), np.array([p[1] for p in pts1])
    xPixScale = _wcs.calcAngSepDeg(ra0, dec0, ra1, dec0)
    yPixScale = _wcs.calcAngSepDeg(ra0, dec0, ra0, dec1)
    return xPixScale * yPixScale


def getPixelAreaArcmin2Map(shape, wcs):
    """Map of pixel area in arcmin^2 (nemo/maps.py:1461-1482)."""
    return np.array([rowAreasDeg2(shape, wcs) * (60 ** 2)] * int(shape[1])).transpose()


def startRMSTables(ops, rms, mask, rowArea, S, noiseTable):
    """Device part for a batch: rms [nImg, H, W], mask u8 [nImg, H, W], rowArea [nImg / S, H] f64 (device), noiseTable
    [nImg, ...] f64 (the cell / bin table of the noise step).  Nothing is synchronised; pass the result to finishRMSTables."""
    torch = ops.torch
    nImg = rms.shape[0]
    cand = torch.sort(noiseTable.reshape(nImg, -1).to(torch.float64), dim=1).values.contiguous()   # index table, not map data
    area, total, miss = ops.rms_area_table(rms, mask, rowArea, S, cand)
    return cand, area, total, miss


def finishRMSTables(job, tol=0.003):
    """[(y0RMS ascending, areaDeg2)] per image; raises like the reference when the areas do not add up to the masked area."""
    cand, area, total, miss = (device.toHost(t) for t in job)
    out = []
    for i in range(cand.shape[0]):
        if miss[i] != 0:
            raise Exception("RMS map holds %d values that are not in its noise table" % int(miss[i]))
        keep = area[i] > 0
        values, areas = cand[i][keep], area[i][keep]
        if abs(areas.sum() - total[i]) > tol:
            raise Exception("Mismatch between area map and area in RMSTab (image %d of the batch)" % i)
        out.append((values, areas))
    return out


def getRMSTab(RMSMap, areaMask, wcs, noiseTable, precision=None):
    """Table of map area (``areaDeg2``) against noise level (``y0RMS``) for one tile -- completeness.py:981-999 on arrays.
    RMSMap / areaMask: what filterMaps returned for the reference filter (its RMS map and ``surveyMask``); noiseTable: the
    filter object's ``noiseTable[0]`` (the values the noise step painted)."""
    ops = device.getOps(precision)
    torch = ops.torch
    rms = device.asTensor(RMSMap, ops)
    rms = rms.unsqueeze(0) if rms.dim() == 2 else rms
    mask = device.asTensor(areaMask, ops, dtype=torch.uint8)
    mask = mask.unsqueeze(0) if mask.dim() == 2 else mask
    table = device.asTensor(noiseTable, ops, dtype=torch.float64)
    rowArea = ops.to_device(rowAreasDeg2(rms.shape[1:], wcs)[None], torch.float64)
    values, areas = finishRMSTables(startRMSTables(ops, rms.contiguous(), mask.contiguous(), rowArea, 1, table.reshape(1, -1)))[0]
    return catalogs.Table.fromColumns([('areaDeg2', areas), ('y0RMS', values)])
