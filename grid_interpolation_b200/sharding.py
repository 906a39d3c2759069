"""Row-band / point-range sharding over ``torch.distributed`` ranks (one process per GPU of one node).

The hot path shards naturally: regular-grid targets are split into contiguous row bands,
scattered bilinear targets into contiguous point ranges; the source mesh / points are replicated on every rank
(``broadcast_inputs``: one host -> device copy on rank 0, then NCCL over NVLink).  The final gather of the output
field is fused into the kernels: every rank stores its band straight into the field of every rank over NVLink
(``PeerField`` + ``barycentric_gather``, CUDA IPC peer mappings handed out by libgridinterp), so no collective
runs after the compute.  ``all_gather_field`` is the plain collective form (NCCL ``all_gather_into_tensor``; gloo
on CPU tensors for the host-side logic tests) that the fused form is measured against.

PyTorch is plumbing here: process group, device buffers, streams.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np


def split_range(n: int, world: int) -> List[Tuple[int, int]]:
    """``world`` contiguous, balanced, ordered ranges covering [0, n) (sizes differ by at most 1)."""
    if world < 1 or n < 0:
        raise ValueError("world must be >= 1 and n >= 0")
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def row_band(len_y: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [begin, end) of the (len_y, len_x) field owned by ``rank``."""
    return split_range(len_y, world)[rank]


def gather_bands(len_y: int, world: int, dest: int, dest_share: float, align: int = 32) -> List[Tuple[int, int]]:
    """Row bands for a gather onto rank ``dest``: that rank computes ``dest_share`` of the rows, the others split the
    rest evenly (band heights are multiples of ``align`` rows, the strip height of the locate / evaluate kernel,
    except for one remainder band).

    Why uneven: when the field is gathered onto ONE rank, that rank's NVLink ingress carries (1 - share) of the field
    (2.1 GB at 16384 x 16384 doubles: 2.1 ms at 900 GB/s for 7/8 of it) while its SMs finish a 1/N band in a fraction
    of that.  Peer stores need no SM on the receiving side, so the gathering rank can compute a larger band while the
    rest arrives: the step is shortest where compute(share) = ingress(1 - share)."""
    if not 0 <= dest < world:
        raise ValueError("dest must be a rank of the group")
    if world == 1:
        return [(0, len_y)]
    if len_y < 4 * align * world:      # small fields: no point in aligning, and no rank should end up empty
        align = 1
    share = min(max(float(dest_share), 0.0), 1.0)
    n_dest = int(round(share * len_y / align)) * align
    n_dest = min(max(n_dest, min(align, len_y)), len_y)
    rest_rows = len_y - n_dest
    units = split_range((rest_rows + align - 1) // align, world - 1)        # the other ranks, in units of `align` rows
    other, used = [], 0
    for lo_u, hi_u in units:                                               # the last of them absorbs the remainder
        sz = min((hi_u - lo_u) * align, rest_rows - used)
        other.append(sz)
        used += sz
    out, lo, j = [], 0, 0
    for r in range(world):
        if r == dest:
            sz = n_dest
        else:
            sz = other[j]
            j += 1
        out.append((lo, lo + sz))
        lo += sz
    return out


def halo_for_fill(default: int = 8) -> int:
    """Extra rows walked on each side of a band so the hull fill sees neighbouring bands (GI_ERR_HALO if too
    small).  8 covers every mesh whose hull leaves less than 8 empty grid rows inside a band boundary."""
    return default


def _retry_band_call(call, halo, max_halo, band_pack):
    """Run ``call(halo, band_pack)`` synchronously; on GI_ERR_HALO double the halo, on GI_ERR_PACK_WINDOW drop the
    band-limited pack.  The decision is made on the status code the exception carries."""
    from . import _lib
    h = halo
    while True:
        try:
            return h, call(h, band_pack)
        except _lib.GiError as e:
            if e.status == _lib.GI_ERR_PACK_WINDOW and band_pack:
                band_pack = False
            elif e.status == _lib.GI_ERR_HALO and h < max_halo:
                h = max(2 * h, 1)
            else:
                raise


def barycentric_band(ip, points, data_in, x_axis, y_axis, simplices, neighbours, rank, world, halo=None,
                     max_halo=4096, band_pack=True, **kw):
    """This rank's band of ``barycentric_interpolation``.  Only the triangles near the band are packed; the halo is
    doubled on GI_ERR_HALO and the whole mesh packed on GI_ERR_PACK_WINDOW (both need the call's status, so the
    call is synchronous whatever ``sync`` says)."""
    band = row_band(len(y_axis), rank, world)
    h = halo_for_fill() if halo is None else halo
    kw = dict(kw)
    kw["sync"] = True
    is_torch = type(points).__module__.split(".")[0] == "torch"

    def call(hh, bp):
        extra = {"band_pack": bp} if (is_torch and world > 1) else {}
        return ip.barycentric_interpolation_ex(points, data_in, x_axis, y_axis, simplices, neighbours, rows=band,
                                               halo=hh if world > 1 else 0, **extra, **kw)

    _, res = _retry_band_call(call, h, max_halo, band_pack)
    return band, res


def all_gather_field(band_tensor, len_y: int, group=None):
    """Assemble the (len_y, len_x) field on every rank from the per-rank row bands (C-order tensors).

    Bands may differ by one row; they are padded to the largest band for ``all_gather_into_tensor``.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bands = split_range(len_y, world)
    rows_max = max(hi - lo for lo, hi in bands)
    len_x = band_tensor.shape[1]
    padded = band_tensor
    if band_tensor.shape[0] != rows_max:
        padded = torch.zeros((rows_max, len_x), dtype=band_tensor.dtype, device=band_tensor.device)
        padded[: band_tensor.shape[0]] = band_tensor
    gathered = torch.empty((world * rows_max, len_x), dtype=band_tensor.dtype, device=band_tensor.device)
    dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
    if all(hi - lo == rows_max for lo, hi in bands):
        return gathered
    parts = [gathered[r * rows_max: r * rows_max + (hi - lo)] for r, (lo, hi) in enumerate(bands)]
    return torch.cat(parts, 0)


def broadcast_inputs(arrays, src: int = 0, device=None, group=None, meta=None):
    """The initial broadcast of the replicated inputs (mesh / points / axes): rank ``src`` passes host (numpy, ideally
    pinned) or device arrays, the other ranks pass None; every rank gets device tensors.  One host -> device copy
    on ``src``, then NCCL over NVLink -- instead of every rank pulling its own copy over PCIe.  ``meta`` = list of
    (shape, numpy dtype) known to every rank skips the (slow, pickled) exchange of the shapes."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    device = device or torch.device("cuda", torch.cuda.current_device())
    if meta is None:
        box = [None]
        if rank == src:
            box = [[(tuple(a.shape), str(a.dtype).replace("torch.", "")) for a in arrays]]
        dist.broadcast_object_list(box, src=src, group=group)
        meta = box[0]
    out = []
    for i, (shape, dt) in enumerate(meta):
        if rank == src:
            a = arrays[i]
            t = a.to(device) if hasattr(a, "data_ptr") else torch.as_tensor(a).to(device, non_blocking=True)
        else:
            t = torch.empty(tuple(shape), dtype=getattr(torch, str(np.dtype(dt))), device=device)
        dist.broadcast(t, src=src, group=group)
        out.append(t)
    return out


class _DeviceView:
    """``__cuda_array_interface__`` of a raw device pointer (so torch can wrap library-owned memory zero-copy)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerField:
    """A (len_y, len_x) output field that exists once per rank and that every rank can store into.

    Each rank allocates its copy (+ a small flag array) through libgridinterp (``gi_peer_alloc``), exports the CUDA
    IPC handles, all-gathers them over the process group and maps the other ranks' copies (``gi_peer_open``).
    ``ptrs[0]`` is the local copy, the rest are peers; ``tensor`` views the local copy as a torch tensor.
    ``signal()`` / ``wait()`` turn "every band has arrived in my copy" into a stream-ordered event.
    """

    def __init__(self, len_y: int, len_x: int, dtype=np.float64, order: str = "C", group=None):
        import torch
        import torch.distributed as dist
        from . import _lib
        self._lib = _lib.lib()
        self._check = _lib.check
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.GI_MAX_GATHER_FIELDS:
            raise ValueError(f"at most {_lib.GI_MAX_GATHER_FIELDS} ranks")
        self.dtype = np.dtype(dtype)
        self.shape, self.order = (int(len_y), int(len_x)), order
        nbytes = self.shape[0] * self.shape[1] * self.dtype.itemsize
        self._local = self._lib.gi_peer_alloc(nbytes)
        self._flags = self._lib.gi_peer_alloc(4 * _lib.GI_MAX_GATHER_FIELDS)
        if not self._local or not self._flags:
            raise MemoryError(_lib.last_error())
        hb = _lib.GI_PEER_HANDLE_BYTES
        mine = (C.c_char * (2 * hb))()
        self._check(self._lib.gi_peer_export(self._local, C.addressof(mine)), "gi_peer_export")
        self._check(self._lib.gi_peer_export(self._flags, C.addressof(mine) + hb), "gi_peer_export")
        dev = torch.device("cuda", torch.cuda.current_device())
        use_cuda = dist.get_backend(group) == "nccl"
        send = torch.frombuffer(bytearray(bytes(mine)), dtype=torch.uint8)
        send = send.to(dev) if use_cuda else send
        recv = [torch.empty(2 * hb, dtype=torch.uint8, device=send.device) for _ in range(self.world)]
        dist.all_gather(recv, send, group=group)
        handles = [r.cpu().numpy() for r in recv]
        self._opened = []
        fields, flags = [self._local], [self._flags]
        for r in list(range(self.rank + 1, self.world)) + list(range(self.rank)):   # rotate: spread NVLink targets
            fp, gp = C.c_void_p(), C.c_void_p()
            h = handles[r].tobytes()
            self._check(self._lib.gi_peer_open(h[:hb], C.byref(fp)), "gi_peer_open")
            self._check(self._lib.gi_peer_open(h[hb:], C.byref(gp)), "gi_peer_open")
            self._opened += [fp.value, gp.value]
            fields.append(fp.value); flags.append(gp.value)
        self.ptrs = (C.c_void_p * self.world)(*fields)
        self.flag_ptrs = (C.c_void_p * self.world)(*flags)
        order_r = [self.rank] + list(range(self.rank + 1, self.world)) + list(range(self.rank))
        self.ptr_of = {r: fields[i] for i, r in enumerate(order_r)}     # rank -> its field / flags as mapped here
        self.flag_of = {r: flags[i] for i, r in enumerate(order_r)}
        self._seq = 0
        shape = self.shape if order == "C" else self.shape[::-1]
        t = torch.as_tensor(_DeviceView(self._local, shape, self.dtype.str), device=dev)
        self.tensor = t if order == "C" else t.t()
        dist.barrier(group)

    def signal(self, stream=None, dest=None):
        """After everything enqueued so far on ``stream``: tell every rank (``dest=None``) or rank ``dest`` only that
        this rank's band is in its field."""
        self._seq += 1
        if dest is None:
            self._check(self._lib.gi_peer_signal(self.flag_ptrs, self.world, self.rank, self._seq, _stream(stream)),
                        "gi_peer_signal")
        else:
            one = (C.c_void_p * 1)(self.flag_of[dest])
            self._check(self._lib.gi_peer_signal(one, 1, self.rank, self._seq, _stream(stream)), "gi_peer_signal")
        return self._seq

    def wait(self, seq=None, stream=None, timeout_s: float = 10.0):
        """Enqueue a wait on ``stream`` until all ranks have signalled ``seq`` (default: this rank's last signal)."""
        self._check(self._lib.gi_peer_wait(self._flags, self.world, self._seq if seq is None else seq, timeout_s,
                                           _stream(stream)), "gi_peer_wait")

    def close(self):
        import torch.distributed as dist
        if getattr(self, "_local", None):
            import torch
            torch.cuda.synchronize()
            dist.barrier(self.group)         # nobody may still be storing into a field that is about to go
            for p in self._opened:
                self._lib.gi_peer_close(p)
            dist.barrier(self.group)
            self.tensor = None
            self._lib.gi_peer_free(self._local); self._lib.gi_peer_free(self._flags)
            self._local = self._flags = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def _stream(stream):
    if stream is not None:
        return stream
    import torch
    return torch.cuda.current_stream().cuda_stream


def barycentric_gather(field: PeerField, points, data_in, x_axis, y_axis, simplices, neighbours, *, halo=None,
                       band_pack=True, signal=True, stream=None, dest=None, bands=None):
    """This rank's row band of ``barycentric_interpolation`` (interpolation.f90:362-404,412), stored by the kernels
    into ``field`` on every rank (``dest=None``: the field ends up complete everywhere) or into this rank's own copy
    and rank ``dest``'s (the final gather of the field onto one rank: 1/N of the NVLink traffic per sender).  Inputs
    are torch CUDA tensors in C order (points (n, 2), simplices / neighbours (t, 3) int32); asynchronous on the
    current torch stream.  Check ``lib().gi_stream_status`` (or call ``field.wait()`` + synchronize) before trusting
    the result: GI_ERR_HALO / GI_ERR_PACK_WINDOW ask for a retry with a larger halo / ``band_pack=False``.
    ``bands``: the (begin, end) rows of every rank (default: ``row_band``, even); see ``gather_bands`` for the uneven
    split that suits a gather onto one rank."""
    from . import _lib
    lib = _lib.lib()
    suf = "f64" if field.dtype == np.float64 else "f32"
    if bands is None:
        r0, r1 = row_band(field.shape[0], field.rank, field.world)
    else:
        if len(bands) != field.world or bands[0][0] != 0 or bands[-1][1] != field.shape[0] or any(
                bands[i][1] != bands[i + 1][0] for i in range(field.world - 1)):
            raise ValueError("bands must be one contiguous (begin, end) row range per rank covering the field")
        r0, r1 = bands[field.rank]
    h = (halo_for_fill() if halo is None else halo) if field.world > 1 else 0
    flags = _lib.GI_POINTS_ROWMAJOR | _lib.GI_TOPO_ROWMAJOR | (_lib.GI_FIELD_ROWMAJOR if field.order == "C" else 0)
    if band_pack and field.world > 1:
        flags |= _lib.GI_BARY_BAND_PACK
    for t in (points, data_in, x_axis, y_axis, simplices, neighbours):
        if not (hasattr(t, "is_cuda") and t.is_cuda and t.is_contiguous()):
            raise ValueError("barycentric_gather takes contiguous torch CUDA tensors")
    if dest is None:
        ptrs, n_fields = field.ptrs, field.world
    else:
        dst = [field.ptr_of[field.rank]] + ([field.ptr_of[dest]] if dest != field.rank else [])
        ptrs, n_fields = (C.c_void_p * len(dst))(*dst), len(dst)
    st = getattr(lib, "gi_barycentric_interpolation_gather_" + suf)(
        points.data_ptr(), points.shape[0], data_in.data_ptr(), x_axis.data_ptr(), x_axis.shape[0], y_axis.data_ptr(),
        y_axis.shape[0], simplices.data_ptr(), neighbours.data_ptr(), simplices.shape[0], r0, r1, h, ptrs,
        n_fields, flags, _stream(stream))
    _lib.check(st, "barycentric_gather")
    if signal:
        field.signal(stream, dest=dest)
    return r0, r1
