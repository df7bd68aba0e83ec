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


def rebalance_rows(bounds, times, height: int, quantum: int = 16):
    """Row bands of equal measured COST instead of equal height.  `bounds` = the current band edges [b0=0, b1, ..., bN=height],
    `times[k]` = the measured time of band k.  Lights are spatially clustered, so equal-height bands finish at different
    times and the frame waits for the slowest rank; treating each band's cost as uniform over its rows gives a piecewise
    linear cumulative cost whose N-quantiles are the new edges (rounded to whole tile rows).  Iterating this over a few
    frames of a static scene converges to bands that finish together."""
    n = len(times)
    assert len(bounds) == n + 1
    total = float(sum(times))
    if total <= 0.0:
        return list(bounds)
    new = [0]
    k, acc = 0, 0.0            # acc = cost of the bands before band k
    for j in range(1, n):
        target = total * j / n
        while k < n - 1 and acc + times[k] < target:
            acc += times[k]
            k += 1
        rows = bounds[k + 1] - bounds[k]
        frac = (target - acc) / times[k] if times[k] > 0 else 0.0
        edge = bounds[k] + frac * rows
        edge = int(round(edge / quantum)) * quantum
        edge = min(max(edge, new[-1]), height)
        new.append(edge)
    new.append(height)
    return new
