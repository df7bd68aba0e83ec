"""Multi-GPU partitioning of the two hot paths: one process per GPU.

* lighting: the framebuffer is cut into `world` contiguous row bands of equal height (the last may be short); every
  rank holds the whole distance field and light list, renders its band, and one all-gather of the bands reassembles
  the lit buffer on every rank (the only collective of the path).
* particles: chunks are independent, so each rank owns a contiguous chunk range; no collective.
"""
from __future__ import annotations

from typing import Tuple


def band_height(height: int, world: int) -> int:
    """Rows per rank; all-gather needs equal-sized shards, so the gathered buffer has band_height*world rows."""
    return (height + world - 1) // world


def row_band(rank: int, world: int, height: int) -> Tuple[int, int]:
    """Rows [begin, end) rendered by `rank` (empty for ranks past the end of a short frame)."""
    h = band_height(height, world)
    return min(rank * h, height), min((rank + 1) * h, height)


def chunk_range(rank: int, world: int, chunks: int) -> Tuple[int, int]:
    """Chunks [begin, end) owned by `rank`: as even as possible, earlier ranks take the remainder."""
    base, rem = divmod(chunks, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
