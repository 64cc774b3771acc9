"""Static tile assignment of the pair matrix over the GPUs of one box.

Every pair is independent, so the path shards with NO collective on the data path: the sequence
sets are replicated on every GPU (C3: 32 MB), the ordered pair matrix is cut into rectangular
tiles, tiles are dealt to GPUs longest-processing-time-first by DP cells, each GPU aligns its
tiles and its D2H copies land in its slice of the result ON THE HOST.

* One process driving all GPUs (the product path, taxi2_b200/multi.py): the slices belong to one
  (page-locked) numpy array.
* One process per GPU (torchrun): `SharedHostMatrix` is the same thing across processes -- a
  matrix in host shared memory every rank maps and writes its tiles into.  No NCCL on the data
  path; the only collective left is `reduce_subset_statistics`, for callers that aggregate per
  rank instead of over the gathered matrix.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Tile:
    index: int
    x0: int
    nx: int
    y0: int
    ny: int
    cells: int

    @property
    def pairs(self) -> int:
        return self.nx * self.ny


def make_tiles(len_x: np.ndarray, len_y: np.ndarray, tile_x: int, tile_y: int) -> list[Tile]:
    """Row-major tiles of the len(len_x) x len(len_y) ordered pair matrix with their DP-cell cost."""
    len_x = np.asarray(len_x, dtype=np.int64)
    len_y = np.asarray(len_y, dtype=np.int64)
    cx = np.concatenate([[0], np.cumsum(len_x)])
    cy = np.concatenate([[0], np.cumsum(len_y)])
    tiles = []
    for x0 in range(0, len(len_x), tile_x):
        nx = min(tile_x, len(len_x) - x0)
        for y0 in range(0, len(len_y), tile_y):
            ny = min(tile_y, len(len_y) - y0)
            cells = int(cx[x0 + nx] - cx[x0]) * int(cy[y0 + ny] - cy[y0])
            tiles.append(Tile(len(tiles), x0, nx, y0, ny, cells))
    return tiles


def assign_tiles(tiles: list[Tile], world: int) -> list[list[Tile]]:
    """Longest-processing-time-first: heaviest tile to the least loaded rank.  Deterministic
    (ties by tile index, then rank), so every rank computes the same plan without talking."""
    loads = [0] * world
    plan: list[list[Tile]] = [[] for _ in range(world)]
    for tile in sorted(tiles, key=lambda t: (-t.cells, t.index)):
        r = min(range(world), key=lambda k: (loads[k], k))
        plan[r].append(tile)
        loads[r] += tile.cells
    for mine in plan:
        mine.sort(key=lambda t: t.index)
    return plan


class SharedHostMatrix:
    """Host gather across PROCESSES without a collective: a (nx, ny, *item) matrix in shared
    memory (a file under /dev/shm by default) that every rank of a one-node job maps.  Rank 0
    creates it, the others open it once it exists; a rank writes exactly the tiles the static plan
    gives it, so no two ranks touch the same bytes and no locking is needed.  The caller
    synchronises (a barrier) before reading the whole matrix."""

    def __init__(self, path, shape: tuple, dtype, create: bool):
        from pathlib import Path

        self.path = Path(path)
        self.shape, self.dtype = tuple(int(v) for v in shape), np.dtype(dtype)
        if create:
            self.path.parent.mkdir(parents=True, exist_ok=True)
            self.array = np.lib.format.open_memmap(self.path, mode="w+", dtype=self.dtype, shape=self.shape)
        else:
            self.array = np.load(self.path, mmap_mode="r+")
            if self.array.shape != self.shape or self.array.dtype != self.dtype:
                raise ValueError(f"{self.path}: expected {self.shape} {self.dtype}, found {self.array.shape} {self.array.dtype}")

    def tile(self, tile: Tile) -> np.ndarray:
        """The slice a tile's results go to (contiguous when the tile spans all columns)."""
        return self.array[tile.x0:tile.x0 + tile.nx, tile.y0:tile.y0 + tile.ny]

    def flush(self) -> None:
        self.array.flush()

    def unlink(self) -> None:
        self.array = None
        self.path.unlink(missing_ok=True)


def reduce_subset_statistics(sums: np.ndarray, mins: np.ndarray, maxs: np.ndarray, counts: np.ndarray, group=None):
    """All-reduce of per-(metric, subset_x, subset_y) aggregates of
    /root/reference/src/itaxotools/taxi2/tasks/versus_all.py:57-95 (sum, min, max, n) for callers
    that aggregate per rank (NCCL over NVLink on GPUs, gloo on CPU).

    min, max and n are exact.  `sum` is REASSOCIATED: the reference adds the distances of a subset
    pair in row-major pair order (versus_all.py:67), a per-rank partial sum followed by a tree
    reduction adds the same numbers in another order, so the mean can differ from the reference's
    in the last bits (|delta| <= n * 2^-53 * sum|d|) -- invisible at the "{:.4f}" the files are
    written with, but not bit-identical.  The product path (tasks on a MultiEngine) does not use
    this: it runs the row-major aggregator (taxi_aggregate_subsets) over the gathered blocks in
    reference order and is bit-identical on any number of GPUs."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sums, mins, maxs, counts
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    ts = torch.from_numpy(np.ascontiguousarray(sums)).to(device)
    tn = torch.from_numpy(np.ascontiguousarray(mins)).to(device)
    tx = torch.from_numpy(np.ascontiguousarray(maxs)).to(device)
    tc = torch.from_numpy(np.ascontiguousarray(counts)).to(device)
    dist.all_reduce(ts, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(tn, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(tx, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(tc, op=dist.ReduceOp.SUM, group=group)
    return ts.cpu().numpy(), tn.cpu().numpy(), tx.cpu().numpy(), tc.cpu().numpy()
