"""Row-band / point-range sharding over ``torch.distributed`` ranks (one process per GPU).

The hot path shards without any data-path collective: regular-grid targets are split into
contiguous row bands, scattered bilinear targets into contiguous point ranges; the source mesh / points are
replicated on every rank.  The only optional collective is the gather of the output field for callers that
want it assembled (``all_gather_field``).  Works with the NCCL backend on GPUs and with gloo on CPU tensors
(host-side logic tests).
"""
from __future__ import annotations

from typing import List, Tuple


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


def halo_for_fill(default: int = 8) -> int:
    """Extra rows walked on each side of a band so the hull fill sees neighbouring bands (GI_ERR_HALO if too
    small).  8 covers every mesh whose hull leaves less than 8 empty grid rows inside a band boundary."""
    return default


def barycentric_band(ip, points, data_in, x_axis, y_axis, simplices, neighbours, rank, world, halo=None,
                     max_halo=4096, **kw):
    """This rank's band of ``barycentric_interpolation``; the halo is doubled on GI_ERR_HALO."""
    band = row_band(len(y_axis), rank, world)
    h = halo_for_fill() if halo is None else halo
    while True:
        try:
            return band, ip.barycentric_interpolation_ex(points, data_in, x_axis, y_axis, simplices, neighbours,
                                                         rows=band, halo=h if world > 1 else 0, **kw)
        except RuntimeError as e:
            if "halo" not in str(e) or h >= max_halo:
                raise
            h *= 2


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
