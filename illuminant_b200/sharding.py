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


class SharedHostFrame:
    """The reassembled frame in HOST memory of one node, written by every rank directly (section 8e, host-to-host leg).

    One process per GPU shares the node's host memory: rank 0 creates a POSIX shared-memory object, every rank maps it and
    page-locks the mapping (`ilb_host_register`), and each rank passes `rows(r0, r1)` of it as `lightmap_out` of its
    `ilb_render_lighting_frame` call -- its band of the lit buffer goes from its GPU into the consumer's frame over its own PCIe
    link, so the frame is reassembled by the copy engines of all GPUs at once instead of being gathered on one GPU and
    downloaded by it alone.  No collective, no GPU barrier: a sequence number per rank in the same mapping says "my band of frame
    s is in place", and the consumer's number says "frame s has been taken, its memory may be overwritten".

    `depth` > 1 keeps that many frames in the mapping (frame s lives in slot s % depth), so that the ranks can work on frame
    s + 1 while the consumer still holds frame s: the producers then never wait for the consumer's hand-shake, only for a slot.

    Layout: 4096-byte header (int64 slot 8*k = sequence number of rank k, slot 8*world = the consumer's), then the frame(s).
    """
    HEADER = 4096

    def __init__(self, name: str, height: int, width: int, channels: int, dtype, rank: int, world: int, ctx=None,
                 timeout_s: float = 60.0, depth: int = 1):
        import mmap
        import os
        import time

        import numpy as np
        self.rank, self.world, self.ctx = int(rank), int(world), ctx
        self.shape, self.dtype = (int(height), int(width), int(channels)), np.dtype(dtype)
        if 8 * (self.world + 1) * 8 > self.HEADER:
            raise ValueError("too many ranks for the header")
        self.depth = max(int(depth), 1)
        frame_bytes = int(np.prod(self.shape)) * self.dtype.itemsize
        frame_bytes = (frame_bytes + 4095) // 4096 * 4096          # every frame starts on a page
        nbytes = self.HEADER + self.depth * frame_bytes
        self.path = os.path.join("/dev/shm", name)
        self.owner = self.rank == 0
        if self.owner:
            try:
                os.unlink(self.path)
            except FileNotFoundError:
                pass
            fd = os.open(self.path + ".tmp", os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            os.ftruncate(fd, nbytes)
            os.rename(self.path + ".tmp", self.path)      # appears under its name only at full size, zero-filled
        else:
            deadline = time.monotonic() + timeout_s
            while True:
                try:
                    fd = os.open(self.path, os.O_RDWR)
                    if os.fstat(fd).st_size == nbytes:
                        break
                    os.close(fd)
                except FileNotFoundError:
                    pass
                if time.monotonic() > deadline:
                    raise TimeoutError(f"shared host frame {self.path} did not appear")
                time.sleep(0.005)
        self._map = mmap.mmap(fd, nbytes)
        os.close(fd)
        self._bytes = np.frombuffer(self._map, dtype=np.uint8)
        self.seq = self._bytes[:self.HEADER].view(np.int64)
        used = int(np.prod(self.shape)) * self.dtype.itemsize
        self.frames = [self._bytes[self.HEADER + k * frame_bytes:self.HEADER + k * frame_bytes + used].view(self.dtype).reshape(self.shape)
                       for k in range(self.depth)]
        self.frame = self.frames[0]
        self.registered = False
        if ctx is not None:
            ctx.host_register(self._bytes.ctypes.data, nbytes)
            self.registered = True

    def frame_of(self, s: int):
        """The slot that holds frame s."""
        return self.frames[s % self.depth]

    def rows(self, r0: int, r1: int, s: int = 0):
        """The view a rank passes as `lightmap_out` for its band [r0, r1) of frame s."""
        return self.frames[s % self.depth][r0:r1]

    @staticmethod
    def _spin(pred, timeout_s: float, what: str):
        import os
        import time
        deadline = None
        n = 0
        while not pred():
            n += 1
            if n & 0x3ff == 0:
                os.sched_yield()
                if deadline is None:
                    deadline = time.monotonic() + timeout_s
                elif time.monotonic() > deadline:
                    raise TimeoutError(what)

    def begin(self, s: int, timeout_s: float = 60.0):
        """Blocks until frame s - depth has been taken by the consumer (its slot is about to be overwritten)."""
        slot = 8 * self.world
        self._spin(lambda: int(self.seq[slot]) >= s - self.depth, timeout_s, f"frame {s - self.depth} was never released")

    def publish(self, s: int):
        """This rank's band of frame s is complete in the shared frame (call after the synchronous frame call returned)."""
        self.seq[8 * self.rank] = s

    def wait_complete(self, s: int, timeout_s: float = 60.0):
        """Consumer: blocks until every rank has published frame s."""
        slots = [8 * k for k in range(self.world)]
        self._spin(lambda: all(int(self.seq[i]) >= s for i in slots), timeout_s, f"frame {s} incomplete")

    def release(self, s: int):
        """Consumer: frame s has been taken."""
        self.seq[8 * self.world] = s

    def close(self):
        import os
        if getattr(self, "_map", None) is None:
            return
        if self.registered and self.ctx is not None:
            try:
                self.ctx.host_unregister(self._bytes.ctypes.data)
            except Exception:   # noqa: BLE001  (context already gone at interpreter exit)
                pass
        self.seq = self.frame = self.frames = self._bytes = None
        try:
            self._map.close()
        except BufferError:
            pass
        self._map = None
        if self.owner:
            try:
                os.unlink(self.path)
            except FileNotFoundError:
                pass

    def __del__(self):
        try:
            self.close()
        except Exception:   # noqa: BLE001
            pass
