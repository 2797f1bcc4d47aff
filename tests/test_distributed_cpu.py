"""
World-size-2 tests of the multi-GPU plumbing on CPU (gloo): tile sharding, the catalog all-gather that replaces
comm.allgather (reference pipelines.py:322-331) and the noise-table gather that replaces comm.gather (:405).
"""

import os
import socket
import sys
import traceback

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _freePort():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rankCatalog(rank):
    from nemo_b200 import photometry
    n = 3 + rank
    # rank 1's first two objects duplicate rank 0's last two (sources in the overlap of two tiles) with a higher S/N
    ra = 10.0 + np.arange(n) * 0.2 + rank * 0.2
    cols = {'id': np.arange(n) + 1, 'x': np.arange(n) * 1.5, 'y': np.arange(n) * 2.5, 'RADeg': ra, 'decDeg': np.full(n, -1.0),
            'galacticLatDeg': np.full(n, -60.0), 'numSigPix': np.full(n, 7.0), 'SNR': 5.0 + rank * 3 + np.arange(n) * 0.1,
            'flags': np.zeros(n, int), 'template': 'Arnaud_M2e14_z0p4', 'tileName': 'tile%d' % rank, 'objIdent': 'ACT-CL', 'longNames': False}
    return photometry.columnsToTable(cols, {'y_c': np.arange(n) * 1e-1 + rank})


def _worker(rank, world, port, q):
    try:
        os.environ.update({'MASTER_ADDR': '127.0.0.1', 'MASTER_PORT': str(port), 'RANK': str(rank), 'WORLD_SIZE': str(world)})
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        dist.init_process_group(backend='gloo', rank=rank, world_size=world)
        from nemo_b200 import distributed, pipelines
        comm = distributed.Comm()
        assert comm.size == world and comm.rank == rank
        tiles = ['t%d' % i for i in range(7)]
        mine = distributed.shardTiles(tiles, rank, world)
        counts = comm.allGatherArrays(np.array([[len(mine)]], dtype=np.float64))
        assert sum(int(c[0, 0]) for c in counts) == len(tiles)
        # more rows on one rank than the gather buffer holds at first (256): the capacity grows on every rank alike, and
        # ragged and empty contributions come back exact
        big = np.arange((700 if rank == 1 else 0) * 3, dtype=np.float64).reshape(-1, 3) + rank
        parts = comm.allGatherArrays(big)
        assert [p.shape for p in parts] == [(0, 3), (700, 3)] and np.array_equal(parts[1], np.arange(2100.0).reshape(-1, 3) + 1)
        parts = comm.allGatherArrays(np.full((2 + rank, 5), 7 + rank, dtype=np.uint8))
        assert [p.shape for p in parts] == [(2, 5), (3, 5)] and parts[1].dtype == np.uint8 and int(parts[1][2, 4]) == 8
        keys = pipelines.catalogKeys(False, 'yc', True)
        tab = _rankCatalog(rank) if rank == 0 else _rankCatalog(rank)
        # err columns are not in this toy table: the packer writes -99 for missing numeric columns
        # coded form (what the pipeline uses): template / tileName as indices, names re-derived on arrival
        labels, tileNames = ['Arnaud_M2e14_z0p4', 'other'], ['tile0', 'tile1']
        merged = distributed.allGatherCatalog(comm, tab, keys, labels, tileNames)
        noise = distributed.gatherNoiseTables(comm, {'F#tile%d' % rank: np.full((4, 2), float(rank + 1))}, ['F'], tileNames)
        # plain form with string columns; one rank has nothing to contribute
        empty = distributed.allGatherCatalog(comm, [] if rank == 1 else tab, keys)
        noise2 = distributed.gatherNoiseTables(comm, {'F#tile%d' % rank: np.full((4, 2), float(rank + 1))})
        assert sorted(noise2.keys()) == sorted(noise.keys()) and np.array_equal(noise2['F#tile1'], noise['F#tile1'])
        # block form (what the batched driver sends): stacks of tables with (filter, tile) index columns, one collective of bytes;
        # rank 0 has two batches, rank 1 none at all (and therefore no table shape to offer)
        blocks = [distributed.noiseTableBlock(np.full((2, 3, 2), 1.0 + k), [0, 1], [k, k]) for k in range(2)] if rank == 0 else []
        view = distributed.gatherNoiseBlocks(comm, blocks, labels, tileNames, (3, 2) if rank == 0 else None)
        assert sorted(view.keys()) == ['Arnaud_M2e14_z0p4#tile0', 'Arnaud_M2e14_z0p4#tile1', 'other#tile0', 'other#tile1'] and len(view) == 4
        assert view['other#tile1'].shape == (3, 2) and np.all(view['other#tile1'] == 2.0) and np.all(view['Arnaud_M2e14_z0p4#tile0'] == 1.0)
        assert distributed.gatherNoiseBlocks(comm, [], labels, tileNames, None) == {}
        alone = distributed.gatherNoiseBlocks(None, [distributed.noiseTableBlock(np.full((1, 3, 2), 5.0), [1], [0])], labels, tileNames, (3, 2))
        assert list(alone.keys()) == ['other#tile0'] and alone['other#tile0'].shape == (3, 2)
        tabs = distributed.gatherRMSTabs(comm, {'Arnaud_M2e14_z0p4#tile%d' % rank: (np.arange(3 + rank) + 1.0, np.full(3 + rank, 0.5 + rank))},
                                         labels, tileNames)
        assert sorted(tabs.keys()) == ['Arnaud_M2e14_z0p4#tile0', 'Arnaud_M2e14_z0p4#tile1']
        assert np.array_equal(tabs['Arnaud_M2e14_z0p4#tile1'][0], [1.0, 2.0, 3.0, 4.0]) and np.array_equal(tabs['Arnaud_M2e14_z0p4#tile1'][1], [1.5] * 4)
        result = {'n': len(merged), 'ra': list(merged['RADeg']), 'snr': list(merged['SNR']), 'names': list(merged['name']),
                  'tiles': list(merged['tileName']), 'noiseKeys': sorted(noise.keys()), 'noise1': noise['F#tile1'].tolist(),
                  'nEmpty': len(empty), 'max': comm.maxOverRanks(rank + 1.5), 'sum': comm.sumOverRanks(rank + 1.0)}
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, result))
    except Exception:
        q.put((rank, traceback.format_exc()))


@pytest.mark.timeout(300)
def test_catalog_and_noise_table_gather_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _freePort()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = {}
    for _ in procs:
        rank, res = q.get(timeout=240)
        results[rank] = res
    for p in procs:
        p.join(timeout=60)
    for rank, res in results.items():
        assert isinstance(res, dict), "rank %d failed:\n%s" % (rank, res)
    a, b = results[0], results[1]
    assert a == b, "ranks disagree after the all-gather"
    # 3 + 4 objects, two duplicate pairs (RA 10.2, 10.4) of which the higher S/N (rank 1's) survives; sorted by RA
    assert a['n'] == 5 and a['ra'] == sorted(a['ra'])
    assert [round(r, 6) for r in a['ra']] == [10.0, 10.2, 10.4, 10.6, 10.8]
    assert a['tiles'] == ['tile0', 'tile1', 'tile1', 'tile1', 'tile1'] and abs(a['snr'][1] - 8.0) < 1e-12
    assert a['names'][0] == 'ACT-CL J0040.0-0100'
    assert a['noiseKeys'] == ['F#tile0', 'F#tile1'] and a['noise1'] == [[2.0, 2.0]] * 4
    assert a['nEmpty'] == 3 and a['max'] == 2.5 and a['sum'] == 3.0


def test_tile_sharding_covers_every_tile_once():
    from nemo_b200 import distributed, pipelines
    tiles = ['%d_%d' % (i // 4, i % 4) for i in range(18)]
    for size in (1, 2, 4, 8):
        seen = []
        for rank in range(size):
            seen += distributed.shardTiles(tiles, rank, size)
        assert sorted(seen) == sorted(tiles)
    maps = {t: None for t in tiles}
    cfg = pipelines.SimpleConfig({'mapFilters': []}, maps, rank=3, size=8)
    assert cfg.tileNames == tiles[3::8] and cfg.allTileNames == tiles


def test_pack_unpack_round_trip():
    from nemo_b200 import distributed, pipelines
    tab = _rankCatalog(1)
    keys = pipelines.catalogKeys(False, 'yc', True)
    num, strs, numKeys, strKeys = distributed.packTable(tab, keys)
    assert num.shape == (4, len(numKeys)) and strs.shape == (4, len(strKeys) * distributed.STR_WIDTH)
    back = distributed.unpackTable(num, strs, numKeys, strKeys)
    assert list(back['name']) == list(tab['name']) and np.array_equal(back['SNR'], tab['SNR'])
    assert np.all(back['err_y_c'] == -99)
    num0, strs0, _, _ = distributed.packTable([], keys)
    assert num0.shape == (0, len(numKeys)) and strs0.shape[0] == 0
