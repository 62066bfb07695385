"""Tile grid, position maps and rank sharding for patch synthesis (host-side plumbing next to the hot path).

Mirrors what the reference's generation dataset feeds the sampler (``dataloader/dataset.py:203-219,242-281``,
``utils/util.py:138-147``): a full 4x1424x2128 packed-raw frame is cut into overlapping ``ps x ps`` crops
(step = ps - ps//4, last row/column snapped to the border); each crop carries a 2-channel position map
(row / (H-1), col / (W-1)) of its location in the frame.  Crops are independent, so ranks take contiguous slices of
the crop list and never communicate.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

FULL_H, FULL_W = 2848 // 2, 4256 // 2          # packed Sony SID frame (dataset.py:203)


def _axis(n: int, ps: int) -> List[int]:
    step = ps - ps // 4
    pts = list(range(0, n - ps + 1, step))
    if n - (pts[-1] + ps) < ps:                  # reference appends the border-aligned origin (even when it repeats)
        pts.append(n - ps)
    return pts


def tile_origins(ps: int, full_h: int = FULL_H, full_w: int = FULL_W) -> List[Tuple[int, int]]:
    """(x, y) crop origins in the reference's order: y outer, x inner."""
    return [(x, y) for y in _axis(full_h, ps) for x in _axis(full_w, ps)]


def position_map(ps_h: int, ps_w: int, x0: int, y0: int, full_h: int = FULL_H, full_w: int = FULL_W,
                 device=None) -> torch.Tensor:
    rows = torch.arange(y0, y0 + ps_h, device=device).float() / (full_h - 1)
    cols = torch.arange(x0, x0 + ps_w, device=device).float() / (full_w - 1)
    return torch.stack((rows[:, None].expand(ps_h, ps_w), cols[None, :].expand(ps_h, ps_w)), dim=0).contiguous()


def synthetic_condition(batch: int, ps: int, seed: int = 1, iso_ratio_idx: int = 24, first_tile: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic generation inputs of the benchmark shape: clean_img ~ U(0, 0.3) packed-Bayer planes,
    position maps cycling over the 88-tile grid, ISO 800 / ratio 250 (index 24 in the reference's combination table)."""
    g = torch.Generator().manual_seed(seed)
    clean = torch.rand((batch, 4, ps, ps), generator=g) * 0.3
    grid = tile_origins(ps)
    pos = torch.stack([position_map(ps, ps, *grid[(first_tile + i) % len(grid)]) for i in range(batch)])
    return {"clean_img": clean, "position": pos, "iso_ratio_idx": torch.full((batch,), iso_ratio_idx, dtype=torch.long)}


def shard(n_items: int, world_size: int, rank: int) -> range:
    """Static contiguous split of `n_items` independent patches over ranks (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def rank_seed(base_seed: int, rank: int, micro_batch_index: int = 0) -> int:
    """Distinct Philox streams per (rank, micro-batch); 2^20 micro-batches per rank before streams could collide."""
    return (int(base_seed) + (int(rank) << 20) + int(micro_batch_index)) & (2 ** 63 - 1)
