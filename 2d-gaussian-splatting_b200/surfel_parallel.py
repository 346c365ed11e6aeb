"""Multi-GPU sharding of the rasterizer hot path.  One process per GPU.

The reference has no multi-device code at all; the two modes below are the ones
BASELINE.json's configs 4 and 5 name.

1. View-parallel (config 4, and bench.py --gpus N): independent camera views are dealt out one per
   rank; every rank holds the full splat set and runs the whole pipeline for its own views.  There
   is NO collective on the data path.  `allreduce_gradients` is only needed when the ranks train one
   shared model.

2. Tile-band partition of ONE oversized frame (config 5): rank r owns the tile rows
   [floor(gy*r/n), floor(gy*(r+1)/n)).  Every rank preprocesses all splats but clips each splat's tile
   rect to its band (C ABI: surfel_settings.tile_row_begin/end), so only its own instances are
   emitted, sorted and blended; sort keys stay bit-identical to the single-GPU run restricted to the
   band.  The one exchange step is an all-gather of the band outputs (10 planes); in the backward the
   per-pixel cotangents are simply sliced (no communication) and the per-splat gradients, which are
   partial sums over the band's pixels, are summed with one all-reduce.

Collectives go through torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist

TILE = 16


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of independent views to ranks."""
    return list(range(rank, num_views, world))


def tile_rows(H: int) -> int:
    return (H + TILE - 1) // TILE


def tile_row_band(H: int, rank: int, world: int) -> Tuple[int, int]:
    """Tile rows [begin, end) owned by `rank` (floor(gy*r/n) .. floor(gy*(r+1)/n))."""
    gy = tile_rows(H)
    return (gy * rank) // world, (gy * (rank + 1)) // world


def band_pixel_rows(H: int, band: Tuple[int, int]) -> Tuple[int, int]:
    return min(H, band[0] * TILE), min(H, band[1] * TILE)


def gather_band_outputs(planes: torch.Tensor, H: int, rank: int, world: int, group=None) -> torch.Tensor:
    """All-gather of band outputs.  `planes` is this rank's (C,H,W) tensor in which only the rows of
    its own band are meaningful; returns the stitched full (C,H,W) frame on every rank.  Bands differ
    by at most one tile row, so each contribution is padded to the largest band and sent with ONE
    equal-sized all_gather (the natural NCCL collective for this split)."""
    C, H_, W = planes.shape
    assert H_ == H
    bands = [tile_row_band(H, r, world) for r in range(world)]
    rows = [band_pixel_rows(H, b) for b in bands]
    max_rows = max(e - s for s, e in rows)
    s, e = rows[rank]
    send = planes.new_zeros((C, max_rows, W))
    send[:, : e - s] = planes[:, s:e]
    recv = planes.new_empty((world * C, max_rows, W))       # concatenation along dim 0
    if world > 1:
        dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    else:
        recv.copy_(send)
    recv = recv.view(world, C, max_rows, W)
    out = planes.new_empty((C, H, W))
    for r, (rs, re) in enumerate(rows):
        out[:, rs:re] = recv[r, :, : re - rs]
    return out


def slice_band_cotangent(grad: torch.Tensor, H: int, rank: int, world: int) -> torch.Tensor:
    """Backward of gather_band_outputs for this rank: keep its own rows, zero elsewhere (no comms)."""
    s, e = band_pixel_rows(H, tile_row_band(H, rank, world))
    out = torch.zeros_like(grad)
    out[:, s:e] = grad[:, s:e]
    return out


def allreduce_gradients(grads: Sequence[torch.Tensor], group=None) -> None:
    """Sum per-splat gradients across ranks in place (tile-band backward, or shared-model training in
    view-parallel mode).  Tensors are flattened into one bucket: launch latency, not link count, is
    what matters on NVSwitch."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


class _BandGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, planes, H, rank, world, group):
        ctx.meta = (H, rank, world)
        return gather_band_outputs(planes, H, rank, world, group)

    @staticmethod
    def backward(ctx, grad):
        H, rank, world = ctx.meta
        return slice_band_cotangent(grad, H, rank, world), None, None, None, None


def rasterize_tile_band(rasterizer_cls, settings, rank: int, world: int, group=None, **inputs) -> Dict[str, torch.Tensor]:
    """One oversized frame split over `world` GPUs: render this rank's tile-row band with the CUDA op,
    then stitch the full frame on every rank with one all-gather.  Differentiable: gradients of the
    per-splat inputs come back as this band's partial sums (call allreduce_gradients on them)."""
    H = int(settings.image_height)
    band = tile_row_band(H, rank, world)
    rs = settings._replace(tile_rows=band)
    color, radii, allmap = rasterizer_cls(rs)(**inputs)
    full = _BandGather.apply(torch.cat([color, allmap], 0), H, rank, world, group)
    return {"render": full[:3], "allmap": full[3:], "radii": radii, "band": band}
