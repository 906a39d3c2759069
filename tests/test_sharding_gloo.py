"""Host-side logic of the multi-GPU path on CPU: band arithmetic, and a 2-rank gloo run in which every rank
produces its row band (with the CPU oracle standing in for the kernels -- the oracle is the checker here, the
thing under test is the sharding / gather plumbing) and the gathered field equals the single-process field."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from grid_interpolation_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_range_properties():
    for n in (0, 1, 7, 79, 16384, 100_000_001):
        for world in (1, 2, 3, 4, 8):
            parts = sharding.split_range(n, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.row_band(79, 1, 2) == (40, 79)
    with pytest.raises(ValueError):
        sharding.split_range(5, 0)


def test_gather_bands_properties():
    """Uneven row bands for a gather onto one rank: contiguous, ordered, covering, the destination's share honoured,
    the others balanced, edges on strip boundaries for large fields and no empty band for small ones."""
    for len_y in (79, 1000, 16384):
        for world in (1, 2, 3, 8):
            for dest in range(world):
                for share in (0.0, 1.0 / world, 0.36, 0.5, 1.0):
                    b = sharding.gather_bands(len_y, world, dest, share)
                    assert len(b) == world and b[0][0] == 0 and b[-1][1] == len_y
                    assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
                    sizes = [hi - lo for lo, hi in b]
                    assert all(sz >= 0 for sz in sizes)
                    if world > 1 and 0.0 < share < 1.0:
                        assert abs(sizes[dest] - share * len_y) <= 32
                        others = [sz for r, sz in enumerate(sizes) if r != dest]
                        assert max(others) - min(others) <= 64
                        if len_y >= 4 * 32 * world:
                            assert sum(sz % 32 != 0 for sz in sizes) <= 1     # whole strips, one remainder
                        else:
                            assert min(sizes) > 0
    assert sharding.gather_bands(16384, 8, 0, 0.125) == sharding.split_range(16384, 8)
    with pytest.raises(ValueError):
        sharding.gather_bands(10, 2, 2, 0.5)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import workloads
        from oracle import oracle as orc
        w = workloads.c1(num_points=1500)
        ly = w.y_axis.size
        full = np.ascontiguousarray(orc.idw_esrg_nearest(w.points, w.data_pts, w.x_axis, w.y_axis, w.grid_spac, 6))
        lo, hi = sharding.row_band(ly, rank, world)
        band = torch.from_numpy(full[lo:hi].copy())          # what this rank's band call would return
        field = sharding.all_gather_field(band, ly)
        ok = torch.equal(field, torch.from_numpy(full))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ret[rank] = bool(ok) and t.item() == float(world)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_band_gather(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
