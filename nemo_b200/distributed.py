"""
Multi-GPU plumbing: one process per GPU, tiles sharded over ranks, and the two collectives of the reference's
pipeline expressed with ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests):

  * catalog all-gather   -- replaces ``config.comm.allgather(optimalCatalog)``  (nemo/pipelines.py:322)
  * noise-table gather   -- replaces ``config.comm.gather(selFnCollection)``    (nemo/pipelines.py:405)

Payloads are kilobytes, so they are packed into one fixed-width buffer per call (numeric columns as fp64, string
columns as fixed-width bytes) and moved with a count exchange + one padded all_gather: latency-bound, never
bandwidth-bound.  There is no collective on the data path -- (tile, scale) work items are independent.
"""

import os
import numpy as np
from collections.abc import Mapping

from . import catalogs

STR_WIDTH = 48


class Comm(object):
    """Thin wrapper over the default ``torch.distributed`` process group."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        if not dist.is_initialized():
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
            dist.init_process_group(backend=backend)
        self.rank = dist.get_rank()
        self.size = dist.get_world_size()
        if device is None:
            if dist.get_backend() == 'nccl':
                device = torch.device('cuda', torch.cuda.current_device())
            else:
                device = torch.device('cpu')
        self.device = device

    def barrier(self):
        self.dist.barrier()

    def allGatherArrays(self, arr):
        """All ranks contribute a 2-D array with the same number of columns and dtype; returns the list of every
        rank's array, in rank order.

        One collective in the usual case: every rank sends a fixed-capacity byte buffer whose first eight bytes carry its row
        count.  The capacity (kept per row size, the same on every rank because every rank sees every count) only grows: when
        some rank has more rows than fit, all ranks see that in the gathered headers, double the capacity and repeat."""
        torch, dist = self.torch, self.dist
        arr = np.ascontiguousarray(arr)
        n, ncols = arr.shape
        rowBytes = ncols * arr.dtype.itemsize
        caps = self.__dict__.setdefault('_gatherCaps', {})
        cap = caps.get(rowBytes, max(4, 65536 // max(rowBytes, 1)))       # rows; about 64 KB to begin with
        raw = arr.view(np.uint8).reshape(-1) if arr.size else np.zeros(0, dtype=np.uint8)
        while True:
            host = np.zeros(8 + cap * rowBytes, dtype=np.uint8)
            host[:8] = np.frombuffer(np.int64(n).tobytes(), dtype=np.uint8)
            m = min(n, cap) * rowBytes
            host[8:8 + m] = raw[:m]
            buf = torch.from_numpy(host).to(self.device)
            out = torch.empty(self.size * host.size, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(out, buf)
            got = out.cpu().numpy().reshape(self.size, host.size)
            counts = [int(np.frombuffer(got[r, :8].tobytes(), dtype=np.int64)[0]) for r in range(self.size)]
            if max(counts) <= cap:
                caps[rowBytes] = cap
                return [np.frombuffer(got[r, 8:8 + c * rowBytes].tobytes(), dtype=arr.dtype).reshape(c, ncols) for r, c in enumerate(counts)]
            while cap < max(counts):
                cap *= 2
            caps[rowBytes] = cap

    def allGatherBlobs(self, blob, key='blob'):
        """Every rank contributes a byte string (1-D uint8 array) of its own length; returns the list of every rank's bytes, in
        rank order.  One collective in the usual case, as allGatherArrays: fixed-capacity buffers (kept per `key`, growing only)
        whose first eight bytes carry the length."""
        torch, dist = self.torch, self.dist
        raw = np.ascontiguousarray(blob, dtype=np.uint8).reshape(-1)
        n = raw.size
        caps = self.__dict__.setdefault('_blobCaps', {})
        cap = caps.get(key, 65536)
        while True:
            host = np.zeros(8 + cap, dtype=np.uint8)
            host[:8] = np.frombuffer(np.int64(n).tobytes(), dtype=np.uint8)
            m = min(n, cap)
            host[8:8 + m] = raw[:m]
            buf = torch.from_numpy(host).to(self.device)
            out = torch.empty(self.size * host.size, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(out, buf)
            got = out.cpu().numpy().reshape(self.size, host.size)
            counts = [int(np.frombuffer(got[r, :8].tobytes(), dtype=np.int64)[0]) for r in range(self.size)]
            if max(counts) <= cap:
                caps[key] = cap
                return [got[r, 8:8 + c] for r, c in enumerate(counts)]
            while cap < max(counts):
                cap *= 2
            caps[key] = cap

    def maxOverRanks(self, value):
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def maxReduceDevice(self, t):
        """Element-wise maximum of a device tensor over ranks, in place (the OR of u8 masks): one all-reduce."""
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t

    def sumOverRanks(self, value):
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def bindProcessToDeviceCpus(deviceIndex):
    """Pins this process to the CPU cores NVML lists as local to GPU `deviceIndex` (same NUMA node / PCIe root), so that the pinned
    staging buffers it allocates afterwards are first-touched next to the GPU that will read them.  With eight ranks uploading
    at once, host memory placement decides the copy bandwidth each gets.  Returns the core list, or None when NVML or the
    affinity call is unavailable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get('CUDA_VISIBLE_DEVICES')
        index = int(visible.split(',')[deviceIndex]) if visible and all(v.strip().isdigit() for v in visible.split(',')) else deviceIndex
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        nCpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (nCpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and 64 * w + b < nCpu]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shardTiles(tileNames, rank, size):
    """Round-robin split of tiles over ranks; all scales of a tile stay on one rank."""
    return list(tileNames)[rank::size]


def packTable(tab, keys, labels=None, tileNames=None):
    """Table (or []) -> (numeric [n, nNum] fp64, strings [n, nStr * STR_WIDTH] u8) for a fixed column list.

    When `labels` (filter labels) and `tileNames` are given -- lists every rank holds -- the 'template' and 'tileName'
    columns travel as indices into them and 'name' is not sent at all (it is a function of RADeg, decDeg and is
    re-attached, deferred, on the receiving side): the payload is then purely numeric."""
    n = len(tab)
    coded = labels is not None and tileNames is not None
    strKeys = [] if coded else [k for k in keys if k in ('name', 'template', 'tileName')]
    numKeys = [k for k in keys if k not in ('name', 'template', 'tileName')]
    if coded:
        numKeys = numKeys + [k for k in ('template', 'tileName') if k in keys]
    num = np.zeros((n, len(numKeys)), dtype=np.float64)
    strs = np.zeros((n, len(strKeys) * STR_WIDTH), dtype=np.uint8)
    if n > 0:
        for j, k in enumerate(numKeys):
            if coded and k == 'template':
                lut = {v: i for i, v in enumerate(labels)}
                num[:, j] = [lut[v] for v in tab[k]]
            elif coded and k == 'tileName':
                lut = {v: i for i, v in enumerate(tileNames)}
                num[:, j] = [lut[v] for v in tab[k]]
            else:
                num[:, j] = np.asarray(tab[k], dtype=np.float64) if k in tab.keys() else -99
        for j, k in enumerate(strKeys):
            for i, s in enumerate(tab[k]):
                b = str(s).encode()[:STR_WIDTH]
                strs[i, j * STR_WIDTH:j * STR_WIDTH + len(b)] = np.frombuffer(b, dtype=np.uint8)
    return num, strs, numKeys, strKeys


def _objectArray(values):
    out = np.empty(len(values), dtype=object)
    out[:] = list(values)
    return out


def unpackTable(num, strs, numKeys, strKeys, labels=None, tileNames=None, objIdent='ACT-CL', longNames=False):
    tab = catalogs.Table()
    n = num.shape[0]
    cols = {}
    coded = labels is not None and tileNames is not None
    for j, k in enumerate(numKeys):
        if coded and k == 'template':
            cols[k] = _objectArray(labels)[num[:, j].astype(np.intp)]
        elif coded and k == 'tileName':
            cols[k] = _objectArray(tileNames)[num[:, j].astype(np.intp)]
        else:
            cols[k] = num[:, j].copy()
    for j, k in enumerate(strKeys):
        cols[k] = np.array([bytes(strs[i, j * STR_WIDTH:(j + 1) * STR_WIDTH]).rstrip(b'\x00').decode() for i in range(n)], dtype=object)
    if coded and 'RADeg' in cols:
        ra, dec = cols['RADeg'], cols['decDeg']
        cols['name'] = catalogs.deferred(lambda: catalogs.makeNames(ra, dec, prefix=objIdent, longNames=longNames))
    return catalogs.Table.fromColumns([(k, cols[k]) for k in (['name'] if 'name' in cols and coded else []) + list(numKeys) + list(strKeys)])


def allGatherCatalog(comm, tab, keys, labels=None, tileNames=None, objIdent='ACT-CL', longNames=False, ops=None):
    """Every rank gets the concatenation of all ranks' catalogs, with duplicates across tiles removed (the
    reference strips them after its allgather, pipelines.py:329-331) and rows sorted by (RADeg, decDeg)."""
    num, strs, numKeys, strKeys = packTable(tab, keys, labels, tileNames)
    num = np.concatenate(comm.allGatherArrays(num), axis=0)
    if len(strKeys) > 0:
        strs = np.concatenate(comm.allGatherArrays(strs), axis=0)
    else:
        strs = np.zeros((num.shape[0], 0), dtype=np.uint8)
    merged = unpackTable(num, strs, numKeys, strKeys, labels, tileNames, objIdent, longNames)
    ordered = catalogs.Table.fromColumns([(k, merged._cols.raw(k)) for k in keys if k in merged.keys()])
    if len(ordered) > 0:
        # every rank holds the same gathered rows: with `ops` the ranks share the cross-match (each takes 1 / size of the centres
        # on its GPU, one all-reduce of the byte mask) instead of every rank repeating the whole of it on the host
        ordered = catalogs.removeDuplicates(ordered, ops=ops, comm=comm if ops is not None else None)
        ordered.sort(['RADeg', 'decDeg'])
    return ordered


class GatheredTables(Mapping):
    """Read-only {'<label>#<tileName>': table} over the row blocks a gather returned (columns: label index, tile index, flattened
    table); the key index is built the first time it is needed."""

    def __init__(self, blocks, labels, tileNames, shape):
        self._blocks, self._labels, self._tiles, self._shape = blocks, labels, tileNames, shape
        self._index = None

    def _idx(self):
        if self._index is None:
            self._index = {}
            for b, fl in enumerate(self._blocks):
                li, ti = fl[:, 0].astype(int), fl[:, 1].astype(int)
                for r in range(fl.shape[0]):
                    self._index[self._labels[li[r]] + "#" + self._tiles[ti[r]]] = (b, r)
        return self._index

    def __getitem__(self, key):
        b, r = self._idx()[key]
        row = self._blocks[b][r, 2:]
        return row.reshape(self._shape) if self._shape is not None else row.copy()

    def __iter__(self):
        return iter(self._idx())

    def __len__(self):
        return sum(fl.shape[0] for fl in self._blocks)


def noiseTableBlock(tab, labelIndex, tileIndex):
    """Rows (label index, tile index, flattened table) for a stack of noise tables tab [n, nCells, nBins] -- what GatheredTables
    reads and gatherNoiseBlocks sends."""
    tab = np.asarray(tab, dtype=np.float64)
    n = tab.shape[0]
    return np.concatenate([np.asarray(labelIndex, dtype=np.float64).reshape(n, 1), np.asarray(tileIndex, dtype=np.float64).reshape(n, 1),
                           tab.reshape(n, -1)], axis=1)


def gatherNoiseBlocks(comm, blocks, labels, tileNames, shape):
    """The noise tables of every rank as one read-only {'<label>#<tileName>': table} (GatheredTables) on every rank.  blocks: this
    rank's noiseTableBlock arrays (one per tile batch, possibly none); shape: (nCells, nBins) of a table, or None on a rank that has
    none.  ONE collective: the rows travel as bytes behind a header (rows, columns, table shape), so ranks need not agree on the
    width beforehand (a rank without tiles has no tables to know it from).  comm None: this rank's blocks only."""
    blocks = [b for b in blocks if b.shape[0] > 0]
    if comm is None or comm.size == 1:
        return GatheredTables(blocks, labels, tileNames, tuple(shape) if shape is not None else None) if blocks else {}
    local = np.ascontiguousarray(np.concatenate(blocks, axis=0)) if blocks else np.zeros((0, 0))
    sh = tuple(shape) if shape is not None and len(shape) == 2 else (0, 0)
    header = np.array([local.shape[0], local.shape[1], sh[0], sh[1]], dtype=np.int64)
    blob = np.concatenate([header.view(np.uint8), local.view(np.uint8).reshape(-1)])
    out, outShape = [], None
    for raw in comm.allGatherBlobs(blob, key='noiseTables'):
        h = np.frombuffer(raw[:32].tobytes(), dtype=np.int64)
        if h[0] == 0:
            continue
        out.append(np.frombuffer(raw[32:32 + h[0] * h[1] * 8].tobytes(), dtype=np.float64).reshape(int(h[0]), int(h[1])))
        if h[2] * h[3] == h[1] - 2:
            outShape = (int(h[2]), int(h[3]))
    return GatheredTables(out, labels, tileNames, outShape) if out else {}


def gatherNoiseTables(comm, noiseTables, labels=None, tileNames=None):
    """{label#tile: [nCells, nBins] array} from every rank -> the union on every rank (rank 0 is the consumer in the
    reference, pipelines.py:405-420; an all-gather costs the same for these sizes).  With `labels` / `tileNames` the keys
    travel as two index columns in front of the flattened table: one numeric collective."""
    keysSorted = sorted(noiseTables.keys())
    coded = labels is not None and tileNames is not None
    if len(keysSorted) > 0:
        flat = np.stack([np.asarray(noiseTables[k], dtype=np.float64).ravel() for k in keysSorted])
        shape = noiseTables[keysSorted[0]].shape
    else:
        flat = np.zeros((0, 0))
        shape = (0, 0)
    width = int(comm.maxOverRanks(flat.shape[1] if flat.size else 0))
    if flat.shape[1] != width:
        flat = np.zeros((0, width))
    out = {}
    if coded:
        codes = np.zeros((flat.shape[0], 2))
        lutL = {v: i for i, v in enumerate(labels)}
        lutT = {v: i for i, v in enumerate(tileNames)}
        for i, k in enumerate(keysSorted[:flat.shape[0]]):
            lab, tile = k.split('#', 1)
            codes[i] = (lutL[lab], lutT[tile])
        # thousands of tables at 8 ranks x 35 tiles x 16 scales: the rows stay in the gathered arrays and the
        # '<label>#<tile>' dictionary is a view made on first use (building it eagerly cost 10 ms per step)
        blocks = [fl for fl in comm.allGatherArrays(np.concatenate([codes, flat], axis=1)) if fl.shape[0] > 0]
        return GatheredTables(blocks, labels, tileNames, shape if int(np.prod(shape)) == width else None)
    names = np.zeros((len(keysSorted), STR_WIDTH * 2), dtype=np.uint8)
    for i, s in enumerate(keysSorted):
        b = s.encode()[:STR_WIDTH * 2]
        names[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    allFlat = comm.allGatherArrays(flat)
    allNames = comm.allGatherArrays(names)
    for fl, nm in zip(allFlat, allNames):
        for i in range(fl.shape[0]):
            key = bytes(nm[i]).rstrip(b'\x00').decode()
            out[key] = fl[i].reshape(shape) if int(np.prod(shape)) == fl.shape[1] else fl[i]
    return out


def gatherRMSTabs(comm, RMSTabs, labels, tileNames):
    """{label#tile: (y0RMS, areaDeg2)} from every rank -> the union on every rank (the reference gathers its per-tile RMSTab
    objects to rank 0, pipelines.py:395-405).  Rows travel as (label index, tile index, y0RMS, areaDeg2): one collective."""
    lutL = {v: i for i, v in enumerate(labels)}
    lutT = {v: i for i, v in enumerate(tileNames)}
    rows = [np.zeros((0, 4))]
    for key in sorted(RMSTabs.keys()):
        lab, tile = key.split('#', 1)
        values, areas = RMSTabs[key]
        rows.append(np.stack([np.full(len(values), lutL[lab], dtype=np.float64), np.full(len(values), lutT[tile], dtype=np.float64),
                              np.asarray(values, dtype=np.float64), np.asarray(areas, dtype=np.float64)], axis=1))
    out = {}
    for part in comm.allGatherArrays(np.concatenate(rows, axis=0)):
        if part.shape[0] == 0:
            continue
        code = part[:, 0].astype(np.int64) * len(tileNames) + part[:, 1].astype(np.int64)
        for c in np.unique(code):
            sel = code == c
            out[labels[int(c) // len(tileNames)] + "#" + tileNames[int(c) % len(tileNames)]] = (part[sel, 2].copy(), part[sel, 3].copy())
    return out
