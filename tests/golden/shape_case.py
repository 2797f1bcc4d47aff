"""Seeded S/N map with elongated, rotated blobs of many sizes for the ellipse measurements of findObjects
(measureShapes = True, nemo/photometry.py:127-178): a 'filtered map' dictionary with the keys findObjects reads."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPE = (200, 260)
DETECT = dict(threshold=4.0, minObjPix=1, findCenterOfMass=True, removeRings=True, ringThresholdSigma=3, useInterpolator=True,
              rejectBorder=0, measureShapes=True)
SHAPE_KEYS = ['ellipse_PA', 'ellipse_A', 'ellipse_B', 'ellipse_x0', 'ellipse_y0', 'ellipse_e']


def shapeMap(seed=5, shape=SHAPE, numBlobs=26):
    rng = np.random.default_rng(seed)
    H, W = shape
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    sn = rng.normal(0.0, 0.6, shape)
    for k in range(numBlobs):
        cy, cx = rng.uniform(12, H - 12), rng.uniform(12, W - 12)
        a, b = rng.uniform(1.0, 6.5), rng.uniform(0.8, 3.5)
        th = rng.uniform(-np.pi / 2, np.pi / 2)
        amp = rng.uniform(6.0, 40.0)
        if k % 7 == 0:          # a round one and an axis-aligned one now and then: xy = 0 / x2 = y2 corners of the formulas
            b = a
        if k % 9 == 0:
            th = 0.0
        dx, dy = xx - cx, yy - cy
        u = dx * np.cos(th) + dy * np.sin(th)
        v = -dx * np.sin(th) + dy * np.cos(th)
        sn += amp * np.exp(-0.5 * ((u / a) ** 2 + (v / b) ** 2))
    # exactly symmetric blobs on the pixel grid (xy == 0, x2 == y2: the arctan(0 / 0) corner), without noise under them
    for cy, cx, r in ((30, 40, 3.0), (150, 200, 2.2)):
        box = (np.abs(yy - cy) < 12) & (np.abs(xx - cx) < 12)
        sn[box] = 0.0
        sn += np.where(box, 20.0 * np.exp(-0.5 * ((xx - cx) ** 2 + (yy - cy) ** 2) / r ** 2), 0.0)
    return sn


def shapeTile(seed=5, shape=SHAPE):
    from oracle import wcs_stub
    sn = shapeMap(seed, shape)
    wcs = wcs_stub.SimpleCARWCS(shape, pixDeg=0.5 / 60.0, crval=(40.0, -5.0))
    return {'SNMap': sn, 'data': sn * 1e-5, 'surveyMask': np.ones(shape), 'flagMask': np.zeros(shape, dtype=int), 'wcs': wcs,
            'label': 'shapes', 'tileName': 'T0', 'mapUnits': 'yc'}
