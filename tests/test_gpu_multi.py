"""The multi-GPU path on real devices (GPU): row bands + the gather fused into the kernels' stores.

  * one process, two destination fields on the same device through the C ABI (the multi-destination stores, both
    memory orders, the band-limited pack and its GI_ERR_PACK_WINDOW report);
  * 2 and 3 processes (one rank each, CUDA IPC peer mappings; the ranks share cuda:0 when the box has one GPU, the
    process group is gloo so that this also runs there): every rank ends up with the whole field, equal to the
    oracle's, after `PeerField.wait()`; the NCCL collectives themselves need one GPU per rank and are exercised by
    bench.py --gpus N.
"""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from grid_interpolation_b200 import _lib, sharding  # noqa: E402
from grid_interpolation_b200 import interpolation as ip  # noqa: E402
import workloads  # noqa: E402
from oracle import oracle as orc  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev(a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")


@pytest.mark.parametrize("order", ["C", "F"])
def test_gather_entry_point_stores_every_band_into_every_field(c1, order):
    import torch
    want = orc.barycentric_interpolation(c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours)
    ly, lx = want.shape
    d = [dev(a) for a in (c1.points, c1.data_pts, c1.x_axis, c1.y_axis, c1.simplices, c1.neighbours)]
    shape = (ly, lx) if order == "C" else (lx, ly)
    fields = [torch.full(shape, np.nan, dtype=torch.float64, device="cuda:0") for _ in range(3)]
    ptrs = (C.c_void_p * 3)(*[f.data_ptr() for f in fields])
    flags = _lib.GI_POINTS_ROWMAJOR | _lib.GI_TOPO_ROWMAJOR | (_lib.GI_FIELD_ROWMAJOR if order == "C" else 0)
    st = torch.cuda.current_stream().cuda_stream
    for lo, hi in sharding.split_range(ly, 3):
        rc = _lib.lib().gi_barycentric_interpolation_gather_f64(
            d[0].data_ptr(), d[0].shape[0], d[1].data_ptr(), d[2].data_ptr(), lx, d[3].data_ptr(), ly,
            d[4].data_ptr(), d[5].data_ptr(), d[4].shape[0], lo, hi, 8, ptrs, 3, flags, st)
        _lib.check(rc, "gather")
    _lib.check(_lib.lib().gi_stream_status(st), "gather")
    for f in fields:
        got = f.cpu().numpy() if order == "C" else f.cpu().numpy().T
        assert np.array_equal(got, want)


def test_band_limited_pack_gives_the_same_band_or_asks_for_a_retry():
    """Only the triangles near the band are packed (GI_BARY_BAND_PACK): the band equals the same rows of the one-shot
    field; a mesh whose triangles are far taller than the margin makes walks leave the packed set, which must be
    reported as GI_ERR_PACK_WINDOW (never a silently wrong value) and is handled by the sharding helper."""
    w = workloads.c4(40_000, 1024, seed=11)
    d = [dev(a) for a in (w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours)]
    full = ip.barycentric_interpolation(*d, order="C").cpu().numpy()
    for rows in ((0, 128), (384, 512), (896, 1024)):
        band = ip.barycentric_interpolation_ex(*d, rows=rows, halo=8, order="C", band_pack=True).cpu().numpy()
        assert np.array_equal(band, full[rows[0]:rows[1]])
    host = ip.barycentric_interpolation_ex(w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours,
                                           rows=(384, 512), halo=8, order="C")          # decides by itself
    assert np.array_equal(host, full[384:512])
    # 300 points over the same 1024 x 1024 grid: triangles ~ 80 rows tall, walks cross the 48-row margin and the hull
    # fill looks further than any small halo: the raw call must report one or the other, the helper retries both
    w = workloads.c4(300, 1024, seed=12)
    d = [dev(a) for a in (w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours)]
    full = ip.barycentric_interpolation(*d, order="C").cpu().numpy()
    rows = (500, 516)
    try:
        band = ip.barycentric_interpolation_ex(*d, rows=rows, halo=8, order="C", band_pack=True).cpu().numpy()
        assert np.array_equal(band, full[rows[0]:rows[1]])       # the walks happened to stay inside: still right
    except _lib.GiError as e:
        assert e.status in (_lib.GI_ERR_PACK_WINDOW, _lib.GI_ERR_HALO)
    (lo, hi), band = sharding.barycentric_band(ip, *d, rank=31, world=64, order="C")     # 16-row band, retries
    assert np.array_equal(band.cpu().numpy(), full[lo:hi])
    host = ip.barycentric_interpolation_ex(w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours,
                                           rows=rows, halo=400, order="C")
    assert np.array_equal(host, full[rows[0]:rows[1]])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, order, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = workloads.c1(num_points=3000)
        want = orc.barycentric_interpolation(w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours)
        d = [torch.as_tensor(np.ascontiguousarray(a), device="cuda") for a in
             (w.points, w.data_pts, w.x_axis, w.y_axis, w.simplices, w.neighbours)]
        ok = True
        with sharding.PeerField(want.shape[0], want.shape[1], order=order) as field:
            for step in range(3):                      # repeated steps: the sequence numbers of signal / wait advance
                sharding.barycentric_gather(field, *d)
                field.wait()
                torch.cuda.synchronize()
                _lib.check(_lib.lib().gi_stream_status(torch.cuda.current_stream().cuda_stream), "gather")
                ok = ok and np.array_equal(field.tensor.cpu().numpy(), want)
                dist.barrier()
                field.tensor.zero_()
                torch.cuda.synchronize()
                dist.barrier()
            # the gather onto ONE rank: every rank stores into its own copy and rank 1's; only rank 1 waits
            root = 1
            r0, r1 = sharding.barycentric_gather(field, *d, dest=root)
            if rank == root:
                field.wait()
            torch.cuda.synchronize()
            _lib.check(_lib.lib().gi_stream_status(torch.cuda.current_stream().cuda_stream), "gather to one rank")
            got = field.tensor.cpu().numpy()
            ok = ok and (np.array_equal(got, want) if rank == root else
                         (np.array_equal(got[r0:r1], want[r0:r1]) and not got[:r0].any() and not got[r1:].any()))
            dist.barrier()
            # the same with uneven bands: the gathering rank computes 60 % of the rows (sharding.gather_bands)
            field.tensor.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            bands = sharding.gather_bands(want.shape[0], world, root, 0.6)
            r0, r1 = sharding.barycentric_gather(field, *d, dest=root, bands=bands)
            assert (r0, r1) == bands[rank] and bands[root][1] - bands[root][0] > want.shape[0] // world
            if rank == root:
                field.wait()
            torch.cuda.synchronize()
            _lib.check(_lib.lib().gi_stream_status(torch.cuda.current_stream().cuda_stream), "uneven gather")
            got = field.tensor.cpu().numpy()
            ok = ok and (np.array_equal(got, want) if rank == root else np.array_equal(got[r0:r1], want[r0:r1]))
            dist.barrier()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,order", [(2, "C"), (3, "F")])
def test_ranks_store_their_bands_into_every_ranks_field(world, order):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), order, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
