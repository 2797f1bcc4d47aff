"""
Catalog plumbing for the hot path's outputs (reference nemo/catalogs.py: catalogListToTab :443-468, makeName
:259-274, makeOptimalCatalog :149-189).  The reference builds astropy Tables; astropy is optional here: a small
column-oriented :class:`Table` with the handful of operations the hot path uses (add_column, row iteration
with item assignment, len, column access, vstack, sort) stands in when it is absent.
"""

import numpy as np

from . import wcs as _wcs

XMATCH_RADIUS_DEG = 1.4 / 60.0

COLUMN_NAMES = ['name', 'RADeg', 'decDeg', 'SNR', 'numSigPix', 'template', 'tileName', 'flags', 'galacticLatDeg',
                'x', 'y', 'id', 'ellipse_PA', 'ellipse_A', 'ellipse_B', 'ellipse_x0', 'ellipse_y0', 'ellipse_e']


class Column(object):
    def __init__(self, data, name):
        self.data = np.asarray(data)
        self.name = name


class _Row(object):
    def __init__(self, table, index):
        self._t = table
        self._i = index

    def __getitem__(self, key):
        return self._t._cols[key][self._i]

    def __setitem__(self, key, value):
        self._t._cols[key][self._i] = value

    def keys(self):
        return self._t.keys()

    def asDict(self):
        return {k: self._t._cols[k][self._i] for k in self._t.keys()}


class _LazyDict(dict):
    """dict whose values may be zero-argument callables evaluated (once) on first access."""

    def defer(self, key, maker):
        dict.__setitem__(self, key, _Deferred(maker))

    def put(self, key, value):
        dict.__setitem__(self, key, value)

    def isDeferred(self, key):
        return isinstance(dict.__getitem__(self, key), _Deferred)

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if isinstance(v, _Deferred):
            v = v.value()
            dict.__setitem__(self, key, v)
        return v

    def raw(self, key):
        return dict.__getitem__(self, key)

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


class _Deferred(object):
    def __init__(self, maker):
        self.maker = maker
        self._value = None

    def value(self):
        if self._value is None:
            self._value = np.asarray(self.maker())
            self.maker = None
        return self._value

    def then(self, fn):
        """A deferred column derived from this one (row subset, reordering) without evaluating it now."""
        return _Deferred(lambda: fn(self.value()))


def deferred(maker):
    """A column value for Table.fromColumns that is computed by ``maker()`` when first read."""
    return _Deferred(maker)


class Table(object):
    """Minimal stand-in for astropy.table.Table."""

    def __init__(self, names=None):
        self._cols = _LazyDict()
        if names is not None:
            for n in names:
                self._cols[n] = np.zeros(0)

    @classmethod
    def fromColumns(cls, pairs):
        """Table from (name, array or deferred(...)) pairs of equal length, without per-column checks."""
        tab = cls()
        for name, data in pairs:
            tab._cols.put(name, data)
        return tab

    def add_deferred_column(self, name, maker):
        """Column whose values are produced by ``maker()`` the first time it is read (object names are only
        formatted when somebody looks at them)."""
        self._cols.defer(name, maker)

    def add_column(self, col, name=None):
        if isinstance(col, Column):
            name, data = col.name, col.data
        else:
            data = np.asarray(col)
        if any(not self._cols.isDeferred(k) for k in self._cols.keys()) and len(data) != len(self):
            raise ValueError("column length mismatch")
        self._cols[name] = np.array(data)

    def keys(self):
        return list(self._cols.keys())

    @property
    def colnames(self):
        return self.keys()

    def __len__(self):
        for k in self._cols.keys():
            if not self._cols.isDeferred(k):
                return len(self._cols[k])
        for k in self._cols.keys():
            return len(self._cols[k])
        return 0

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._cols[key]
        if isinstance(key, (int, np.integer)):
            return _Row(self, int(key))
        out = Table()
        for k in self._cols.keys():
            out._cols.put(k, self._derived(k, lambda v: v[key]))
        return out

    def _derived(self, k, fn):
        raw = self._cols.raw(k)
        if isinstance(raw, _Deferred):
            return raw.then(fn)
        return fn(raw)

    def __setitem__(self, key, value):
        self._cols[key] = np.asarray(value)

    def __iter__(self):
        for i in range(len(self)):
            yield _Row(self, i)

    def copy(self):
        out = Table()
        for k in self._cols.keys():
            out._cols.put(k, self._derived(k, lambda v: v.copy()))
        return out

    def remove_column(self, name):
        del self._cols[name]

    def sort(self, keys):
        if isinstance(keys, str):
            keys = [keys]
        order = np.lexsort([self._cols[k] for k in reversed(keys)])
        for k in list(self._cols.keys()):
            self._cols.put(k, self._derived(k, lambda v: v[order]))

    def toRows(self):
        return [r.asDict() for r in self]


class RowsOf(Table):
    """Rows [lo, hi) of another table as a table of their own, cut out (copied) when first looked at: a tile batch has hundreds of
    per-image catalogs ('<label>#<tileName>' -> {'catalog': ...}, the reference's layout) and a run usually reads only the
    stacked rows.  The parent must not change afterwards (the batched driver's flat tables never do: vstack, row selection
    and sort all make new tables)."""

    def __init__(self, parent, lo, hi):
        self._parent, self._lo, self._hi = parent, int(lo), int(hi)
        self._made = None

    @property
    def _cols(self):
        if self._made is None:
            made = _LazyDict()
            lo, hi = self._lo, self._hi
            for k in self._parent.keys():
                made.put(k, self._parent._derived(k, lambda v: v[lo:hi].copy()))
            self._made = made
            self._parent = None
        return self._made

    def __len__(self):
        return self._hi - self._lo


def vstack(tables):
    tables = [t for t in tables if len(t) > 0]
    out = Table()
    if not tables:
        return out
    cols = [t._cols for t in tables]
    lengths = None
    for k in tables[0].keys():
        if all(k in c and c.isDeferred(k) for c in cols):
            parts = [c.raw(k) for c in cols]
            out._cols.defer(k, lambda parts=parts: np.concatenate([p.value() for p in parts]))
        elif all(k in c for c in cols):
            out._cols.put(k, np.concatenate([c[k] for c in cols]))
        else:
            if lengths is None:
                lengths = [len(t) for t in tables]
            out._cols.put(k, np.concatenate([c[k] if k in c else np.full(m, -99) for c, m in zip(cols, lengths)]))
    return out


def _splitFraction(text):
    """('12', 0.345) from '12.345': integer digits and the fractional part re-read from its printed digits, which is
    how the reference derives minutes and seconds (catalogs.py:295-346) -- truncation, never rounding."""
    whole, frac = text.split('.')
    return whole, float('.' + frac)


def _nameRA(RADeg, longFormat):
    hoursText = "%.10f" % ((RADeg / 360) * 24) if not longFormat else str((RADeg / 360) * 24)
    whole, frac = _splitFraction(hoursText)
    out = whole.rjust(2, '0')[:2]
    mins = frac * 60
    if not longFormat:
        minsText = "%.10f" % mins
        return out + (minsText[:3].rjust(4, '0') if mins < 10 else minsText[:4])       # MM.m
    minsText = str(mins)
    out += minsText[:1].rjust(2, '0') if mins < 10 else minsText[:2]
    secs = _splitFraction(minsText)[1] * 60
    secsText = str(secs)
    return out + (secsText[:3].rjust(4, '0') if secs < 10 else secsText[:4])           # SS.s


def _nameDec(decDeg, longFormat):
    text = str(decDeg)
    sign = '+' if decDeg > 0 else '-'
    whole, frac = _splitFraction(text.lstrip('-'))
    out = sign + whole.rjust(2, '0')[:2]
    mins = frac * 60
    minsText = str(mins)
    out += minsText[:1].rjust(2, '0') if mins < 10 else minsText[:2]
    if longFormat:
        secs = _splitFraction(minsText)[1] * 60
        secsText = str(secs)
        out += secsText[:3].rjust(4, '0') if secs < 10 else secsText[:4]
    return out


def makeNames(RADeg, decDeg, prefix='ACT-CL', longNames=False):
    """Names of many objects at once (object array of str)."""
    namer = makeLongName if longNames else makeName
    return np.array([namer(r, d, prefix=prefix) for r, d in zip(np.atleast_1d(RADeg), np.atleast_1d(decDeg))], dtype=object)


def makeName(RADeg, decDeg, prefix='ACT-CL'):
    """Object name `prefix JHHMM.m+/-DDMM`, digits truncated as the reference does (catalogs.py:259-274, 295-346)."""
    return prefix + " J" + _nameRA(float(RADeg), False) + _nameDec(float(decDeg), False)


def makeLongName(RADeg, decDeg, prefix='ACT-CL'):
    """Object name `prefix JHHMMSS.s+/-DDMMSS.s` (catalogs.py:277-292, 349-420)."""
    return prefix + " J" + _nameRA(float(RADeg), True) + _nameDec(float(decDeg), True)


def catalogListToTab(catalogList, keysToWrite=COLUMN_NAMES):
    """List of per-object dicts -> table with the reference's column order."""
    availKeys = list(catalogList[0].keys())
    tab = Table()
    for key in keysToWrite:
        if key in availKeys:
            arr = [obj[key] if obj[key] is not None else -99 for obj in catalogList]
            tab.add_column(Column(arr, key))
    return tab


def _candidatePairs(ra, dec, radiusDeg):
    """Index pairs (i, j), each once, that may lie within radiusDeg of each other: a superset cut on |d dec| and |d RA|
    (a great-circle or tangent-plane separation below the radius implies both), found by a sweep over the declination-sorted
    rows with searchsorted -- no tree to build."""
    n = len(ra)
    order = np.argsort(dec, kind='stable')
    ds = dec[order]
    hi = np.searchsorted(ds, ds + radiusDeg, side='right')      # rows order[i + 1 : hi[i]] are within the declination window of row i
    idx = np.arange(n)
    cnt = hi - (idx + 1)
    total = int(cnt.sum())
    if total == 0:
        return np.zeros((0, 2), dtype=np.intp)
    first = np.repeat(idx, cnt)
    start = np.cumsum(cnt) - cnt
    second = first + 1 + (np.arange(total) - np.repeat(start, cnt))
    i, j = order[first], order[second]
    dRA = np.abs(ra[i] - ra[j])
    dRA = np.minimum(dRA, 360.0 - dRA)
    cosMin = np.cos(np.radians(min(max(abs(ds[0]), abs(ds[-1])) + radiusDeg, 90.0)))     # smallest cos(dec) any pair can see
    if cosMin > 1e-6:
        keep = dRA <= radiusDeg / cosMin * 1.0001
        i, j = i[keep], j[keep]
    return np.stack([i, j], axis=1)


def _candidatePairsGrid(ra, dec, radiusDeg):
    """The same superset of close pairs as _candidatePairs, from a hash of the sky into cells one radius high and one radius
    (at the largest |dec| present) wide: a pair within the radius lies in the same or in adjacent cells.  The number of
    candidates then follows the local density, not the number of rows in a declination strip -- with thousands of rows from
    many overlapping tiles the strip sweep generates hundreds of thousands of pairs.  Returns None when the cells would wrap
    onto themselves in RA (near a pole); the caller falls back to the strip sweep."""
    n = len(ra)
    cosMin = np.cos(np.radians(min(np.abs(dec).max() + radiusDeg, 90.0)))
    if cosMin <= 1e-6:
        return None
    cellRA = radiusDeg / cosMin * 1.0001
    NX = int(np.floor(360.0 / cellRA))
    if NX < 4:
        return None
    cellRA = 360.0 / NX                                  # cells tile the circle exactly; still >= the radius in RA
    ix = np.minimum((np.mod(ra, 360.0) / cellRA).astype(np.int64), NX - 1)
    iy = np.floor((dec + 90.0) / radiusDeg).astype(np.int64)
    key = iy * NX + ix
    order = np.argsort(key, kind='stable')
    ks = key[order]
    idx = np.arange(n)
    out = []
    # every unordered pair of cells once: the cell itself (rows after the current one), its east neighbour, and the three cells
    # of the row above; RA neighbours wrap around
    for dy, dx in ((0, 0), (0, 1), (1, -1), (1, 0), (1, 1)):
        if dy == 0 and dx == 0:
            lo = idx + 1
            hi = np.searchsorted(ks, ks, side='right')
        else:
            nkey = (iy[order] + dy) * NX + np.mod(ix[order] + dx, NX)
            if dx != 0 and NX < 3:
                continue
            lo = np.searchsorted(ks, nkey, side='left')
            hi = np.searchsorted(ks, nkey, side='right')
        cnt = np.maximum(hi - lo, 0)
        total = int(cnt.sum())
        if total == 0:
            continue
        first = np.repeat(idx, cnt)
        start = np.cumsum(cnt) - cnt
        second = np.repeat(lo, cnt) + (np.arange(total) - np.repeat(start, cnt))
        out.append(np.stack([order[first], order[second]], axis=1))
    if not out:
        return np.zeros((0, 2), dtype=np.intp)
    return np.concatenate(out, axis=0)


def _crossMatchKeepBest(ra, dec, snr, radiusDeg):
    """Boolean mask of rows to remove: for every row, all rows within radiusDeg of it except the one with the highest
    S/N (first one on ties) are dropped.  Same outcome as the reference's O(n^2) loop (catalogs.py:171-183): candidate
    pairs come from a declination-sorted sweep, the reference's (tangent-plane, centred on the row) separation decides,
    and the rule is applied to all rows at once."""
    n = len(ra)
    toRemove = np.zeros(n, dtype=bool)
    if n < 2:
        return toRemove
    # the strip sweep is the cheaper one until the strips get crowded (its candidates grow with n^2 / declination range)
    pairs = None
    if float(n) * n * radiusDeg / max(float(dec.max() - dec.min()), radiusDeg) > 5e4:
        pairs = _candidatePairsGrid(ra, dec, radiusDeg * 1.01)
    if pairs is None:
        pairs = _candidatePairs(ra, dec, radiusDeg * 1.01)
    if len(pairs) == 0:
        return toRemove
    # directed edges centre -> member, both directions; the separation is measured about the centre, as the reference does
    centre = np.concatenate([pairs[:, 0], pairs[:, 1]])
    member = np.concatenate([pairs[:, 1], pairs[:, 0]])
    near = _wcs.calcAngSepDeg(ra[centre], dec[centre], ra[member], dec[member]) < radiusDeg
    centre, member = centre[near], member[near]
    if len(centre) == 0:
        return toRemove
    # every centre with at least one neighbour is a member of its own group
    own = np.unique(centre)
    centre = np.concatenate([centre, own])
    member = np.concatenate([member, own])
    # best member of each group: highest S/N, lowest row index on ties (np.argmax over ascending indices)
    order = np.lexsort((member, -snr[member], centre))
    centre, member = centre[order], member[order]
    first = np.ones(len(centre), dtype=bool)
    first[1:] = centre[1:] != centre[:-1]
    best = member[np.flatnonzero(first)[np.cumsum(first) - 1]]
    toRemove[member[member != best]] = True
    return toRemove


def withinImage(RADeg, decDeg, shape, wcs, mask=None):
    """Boolean selection of the positions catalogs.getCatalogWithinImage keeps (catalogs.py:874-892): 0 <= x < W - 1 and
    0 <= y < H - 1, and mask == 1 at the rounded pixel when a mask is given."""
    x, y = wcs.wcs2pix(np.asarray(RADeg, dtype=np.float64), np.asarray(decDeg, dtype=np.float64))
    x, y = np.atleast_1d(np.asarray(x, dtype=np.float64)), np.atleast_1d(np.asarray(y, dtype=np.float64))
    with np.errstate(invalid='ignore'):
        sel = (x >= 0) & (x < shape[1] - 1) & (y >= 0) & (y < shape[0] - 1)
    if mask is not None:
        idx = np.flatnonzero(sel)
        sel[idx] = np.asarray(mask)[np.rint(y[idx]).astype(int), np.rint(x[idx]).astype(int)] == 1
    return sel


def getCatalogWithinImage(tab, shape, wcs, mask=None):
    """The subset of the catalog inside the image (catalogs.py:855-892)."""
    return tab[withinImage(tab['RADeg'], tab['decDeg'], shape, wcs, mask=mask)]


def crossMatchKeepBest(ra, dec, snr, radiusDeg, ops=None, comm=None):
    """Boolean mask of the rows that lose to a higher-S/N row (see _crossMatchKeepBest).  With `ops` (nemo_b200.ops.Ops) the rule is
    evaluated on the device (nb_xmatch_keep_best: every centre against every row).  With `comm` as well (all ranks hold the same
    rows) the ranks share the centres and combine their masks with one all-reduce."""
    if ops is None:
        return _crossMatchKeepBest(ra, dec, snr, radiusDeg)
    from . import device
    n = len(ra)
    if n < 2:
        return np.zeros(n, dtype=bool)
    if comm is not None and comm.size > 1:
        per = (n + comm.size - 1) // comm.size
        first = min(comm.rank * per, n)
        mask = ops.xmatch_keep_best(ra, dec, snr, radiusDeg, first, min(per, n - first))
        mask = comm.maxReduceDevice(mask)
    else:
        mask = ops.xmatch_keep_best(ra, dec, snr, radiusDeg)
    return device.toHost(mask) != 0


def makeOptimalCatalog(catalogDict, constraintsList=[], ops=None, stacked=None):
    """Keeps, for every group of detections within XMATCH_RADIUS_DEG of each other across filters, the one with
    the highest S/N; sorted by (RADeg, decDeg)  (catalogs.py:149-189).  ops: evaluate the cross-match on the device.
    stacked: tables that hold the rows of catalogDict's catalogs already stacked, in catalogDict's order (the batched driver
    has one per tile batch) -- used instead of stacking the per-image tables."""
    if stacked is not None:
        allCatalogs = [t for t in stacked if len(t) > 0]
    else:
        allCatalogs = [catalogDict[key]['catalog'] for key in catalogDict.keys() if len(catalogDict[key]['catalog']) > 0]
    if len(allCatalogs) == 0:
        return Table(names=('SNR', 'RADeg', 'decDeg'))
    allCat = vstack(allCatalogs)
    ra, dec, snr = allCat['RADeg'].astype(float), allCat['decDeg'].astype(float), allCat['SNR'].astype(float)
    merged = allCat[~crossMatchKeepBest(ra, dec, snr, XMATCH_RADIUS_DEG, ops=ops)]
    merged.sort(['RADeg', 'decDeg'])
    if constraintsList:
        raise Exception("catalogCuts constraints are not supported yet")
    return merged


def removeDuplicates(tab, ops=None, comm=None):
    """Drops lower-S/N duplicates of the same object found in overlapping tiles (catalogs.py:520-565)."""
    if len(tab) < 2:
        return tab
    ra, dec, snr = tab['RADeg'].astype(float), tab['decDeg'].astype(float), tab['SNR'].astype(float)
    return tab[~crossMatchKeepBest(ra, dec, snr, XMATCH_RADIUS_DEG, ops=ops, comm=comm)]
