"""Static tile assignment of the pair matrix over the GPUs of one box.

Every pair is independent, so the path shards with NO collective on the compute path: the
sequence sets are replicated on every GPU (C3: 32 MB), the ordered pair matrix is cut into
rectangular tiles, tiles are dealt to ranks longest-processing-time-first by DP cells, each rank
aligns its tiles and writes into its slice of the result.  The only exchange is the final gather
of per-tile results (or the reduction of per-subset summary statistics), done here with
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
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


def run_sharded(engine_factory, len_x, len_y, tile_x, tile_y, rank: int, world: int, compute_tile) -> dict[int, object]:
    """Run `compute_tile(tile)` for this rank's share of the tiles; returns {tile index: result}."""
    tiles = make_tiles(len_x, len_y, tile_x, tile_y)
    return {tile.index: compute_tile(tile) for tile in assign_tiles(tiles, world)[rank]}


def gather_matrix(local: dict[int, np.ndarray], tiles: list[Tile], shape: tuple, dtype, group=None) -> np.ndarray | None:
    """Gather per-tile result blocks to rank 0 in reference (row-major) order.

    local[tile.index] has shape (nx, ny, *shape).  Returns the full (NX, NY, *shape) array on rank 0,
    None elsewhere.  Uses gather_object-free tensor collectives so it works on NCCL and gloo alike.
    """
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    nx_total = max(t.x0 + t.nx for t in tiles)
    ny_total = max(t.y0 + t.ny for t in tiles)
    plan = assign_tiles(tiles, world)
    out = np.zeros((nx_total, ny_total, *shape), dtype=dtype) if rank == 0 else None
    backend = dist.get_backend(group) if dist.is_initialized() else "gloo"
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    for owner in range(world):
        for tile in plan[owner]:
            block_shape = (tile.nx, tile.ny, *shape)
            if owner == 0:
                if rank == 0:
                    out[tile.x0:tile.x0 + tile.nx, tile.y0:tile.y0 + tile.ny] = local[tile.index]
                continue
            if rank == owner:
                dist.send(torch.from_numpy(np.ascontiguousarray(local[tile.index])).to(device), dst=0, group=group)
            elif rank == 0:
                buf = torch.empty(block_shape, dtype=torch.from_numpy(np.zeros(1, dtype=dtype)).dtype, device=device)
                dist.recv(buf, src=owner, group=group)
                out[tile.x0:tile.x0 + tile.nx, tile.y0:tile.y0 + tile.ny] = buf.cpu().numpy()
    return out


def reduce_subset_statistics(sums: np.ndarray, mins: np.ndarray, maxs: np.ndarray, counts: np.ndarray, group=None):
    """All-reduce of the per-(metric, subset_x, subset_y) aggregates of
    /root/reference/src/itaxotools/taxi2/tasks/versus_all.py:57-95 (sum, min, max, n).
    This is the only collective the path needs (NCCL over NVLink on GPUs)."""
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
