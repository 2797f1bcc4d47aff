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
        rank's array, in rank order."""
        torch, dist = self.torch, self.dist
        arr = np.ascontiguousarray(arr)
        ncols = arr.shape[1]
        count = torch.tensor([arr.shape[0]], dtype=torch.int64, device=self.device)
        counts = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(self.size)]
        dist.all_gather(counts, count)
        counts = [int(c.item()) for c in counts]
        nmax = max(max(counts), 1)
        buf = torch.zeros((nmax, ncols), dtype=torch.from_numpy(arr[:0]).dtype, device=self.device)
        if arr.shape[0] > 0:
            buf[:arr.shape[0]] = torch.from_numpy(arr).to(self.device)
        out = [torch.empty_like(buf) for _ in range(self.size)]
        dist.all_gather(out, buf)
        return [o[:n].cpu().numpy() for o, n in zip(out, counts)]

    def maxOverRanks(self, value):
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sumOverRanks(self, value):
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def shardTiles(tileNames, rank, size):
    """Round-robin split of tiles over ranks; all scales of a tile stay on one rank."""
    return list(tileNames)[rank::size]


def packTable(tab, keys):
    """Table (or []) -> (numeric [n, nNum] fp64, strings [n, nStr * STR_WIDTH] u8) for a fixed column list."""
    n = len(tab)
    numKeys = [k for k in keys if k not in ('name', 'template', 'tileName')]
    strKeys = [k for k in keys if k in ('name', 'template', 'tileName')]
    num = np.zeros((n, len(numKeys)), dtype=np.float64)
    strs = np.zeros((n, len(strKeys) * STR_WIDTH), dtype=np.uint8)
    if n > 0:
        for j, k in enumerate(numKeys):
            num[:, j] = np.asarray(tab[k], dtype=np.float64) if k in tab.keys() else -99
        for j, k in enumerate(strKeys):
            for i, s in enumerate(tab[k]):
                b = str(s).encode()[:STR_WIDTH]
                strs[i, j * STR_WIDTH:j * STR_WIDTH + len(b)] = np.frombuffer(b, dtype=np.uint8)
    return num, strs, numKeys, strKeys


def unpackTable(num, strs, numKeys, strKeys):
    tab = catalogs.Table()
    n = num.shape[0]
    cols = {}
    for j, k in enumerate(numKeys):
        cols[k] = num[:, j].copy()
    for j, k in enumerate(strKeys):
        cols[k] = np.array([bytes(strs[i, j * STR_WIDTH:(j + 1) * STR_WIDTH]).rstrip(b'\x00').decode() for i in range(n)], dtype=object)
    for k in list(numKeys) + list(strKeys):
        tab.add_column(catalogs.Column(cols[k], k))
    return tab


def allGatherCatalog(comm, tab, keys):
    """Every rank gets the concatenation of all ranks' catalogs, with duplicates across tiles removed (the
    reference strips them after its allgather, pipelines.py:329-331) and rows sorted by (RADeg, decDeg)."""
    num, strs, numKeys, strKeys = packTable(tab, keys)
    nums = comm.allGatherArrays(num)
    strss = comm.allGatherArrays(strs)
    num = np.concatenate(nums, axis=0)
    strs = np.concatenate(strss, axis=0)
    merged = unpackTable(num, strs, numKeys, strKeys)
    ordered = catalogs.Table()
    for k in keys:
        ordered.add_column(catalogs.Column(merged[k], k))
    if len(ordered) > 0:
        ordered = catalogs.removeDuplicates(ordered)
        ordered.sort(['RADeg', 'decDeg'])
    return ordered


def gatherNoiseTables(comm, noiseTables):
    """{label#tile: [nCells, nBins] array} from every rank -> the union on every rank (rank 0 is the consumer in the
    reference, pipelines.py:405-420; an all-gather costs the same for these sizes)."""
    labels = sorted(noiseTables.keys())
    if len(labels) > 0:
        flat = np.stack([np.asarray(noiseTables[k], dtype=np.float64).ravel() for k in labels])
        shape = noiseTables[labels[0]].shape
    else:
        flat = np.zeros((0, 0))
        shape = (0, 0)
    width = int(comm.maxOverRanks(flat.shape[1] if flat.size else 0))
    if flat.shape[1] != width:
        flat = np.zeros((0, width))
    names = np.zeros((len(labels), STR_WIDTH * 2), dtype=np.uint8)
    for i, s in enumerate(labels):
        b = s.encode()[:STR_WIDTH * 2]
        names[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    allFlat = comm.allGatherArrays(flat)
    allNames = comm.allGatherArrays(names)
    out = {}
    for fl, nm in zip(allFlat, allNames):
        for i in range(fl.shape[0]):
            key = bytes(nm[i]).rstrip(b'\x00').decode()
            out[key] = fl[i].reshape(shape) if int(np.prod(shape)) == fl.shape[1] else fl[i]
    return out
