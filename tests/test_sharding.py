"""Multi-GPU host logic on CPU: deterministic static tile plan, and a world_size-2 job over the
gloo backend -- every rank writes its tiles into the shared host matrix (the N>1 data path has no
collective), then the one remaining collective, the subset-statistics reduction."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from taxi2_b200.sharding import SharedHostMatrix, assign_tiles, make_tiles, reduce_subset_statistics


def test_tiles_cover_the_matrix_once():
    rng = np.random.default_rng(0)
    lx, ly = rng.integers(300, 1500, 37), rng.integers(300, 1500, 53)
    tiles = make_tiles(lx, ly, 8, 16)
    seen = np.zeros((37, 53), dtype=int)
    for t in tiles:
        seen[t.x0:t.x0 + t.nx, t.y0:t.y0 + t.ny] += 1
        assert t.cells == int(lx[t.x0:t.x0 + t.nx].sum()) * int(ly[t.y0:t.y0 + t.ny].sum())
    assert (seen == 1).all()
    assert sum(t.cells for t in tiles) == int(lx.sum()) * int(ly.sum())


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_lpt_plan_is_balanced_and_deterministic(world):
    rng = np.random.default_rng(1)
    lx, ly = rng.integers(300, 1500, 200), rng.integers(300, 1500, 300)
    tiles = make_tiles(lx, ly, 16, 32)
    plan = assign_tiles(tiles, world)
    assert sorted(t.index for mine in plan for t in mine) == list(range(len(tiles)))
    loads = [sum(t.cells for t in mine) for mine in plan]
    assert max(loads) <= 1.05 * (sum(loads) / world) + max(t.cells for t in tiles)
    assert plan == assign_tiles(tiles, world)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_path: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        lx, ly = rng.integers(5, 50, 19), rng.integers(5, 50, 23)
        tiles = make_tiles(lx, ly, 4, 6)
        mine = assign_tiles(tiles, world)[rank]
        # stand-in for the per-tile device result: a function of the global pair index
        shared_path = out_path + ".matrix.npy"
        if rank == 0:
            shared = SharedHostMatrix(shared_path, (19, 23, 2), np.int64, create=True)
        dist.barrier()
        if rank != 0:
            shared = SharedHostMatrix(shared_path, (19, 23, 2), np.int64, create=False)
        local = {}
        for t in mine:
            ii, jj = np.meshgrid(np.arange(t.x0, t.x0 + t.nx), np.arange(t.y0, t.y0 + t.ny), indexing="ij")
            local[t.index] = np.stack([ii * 1000 + jj, lx[ii] * ly[jj]], axis=-1).astype(np.int64)
            shared.tile(t)[...] = local[t.index]          # the host gather: each rank writes its own tiles
        shared.flush()
        dist.barrier()
        full = np.array(shared.array) if rank == 0 else None
        # per-subset aggregates: each rank contributes its own pairs
        vals = np.concatenate([local[t.index][..., 1].ravel() for t in mine]).astype(np.float64)
        s, mn, mx, n = reduce_subset_statistics(np.array([vals.sum()]), np.array([vals.min()]), np.array([vals.max()]),
                                                np.array([len(vals)], dtype=np.int64))
        if rank == 0:
            ii, jj = np.meshgrid(np.arange(19), np.arange(23), indexing="ij")
            want = np.stack([ii * 1000 + jj, lx[ii] * ly[jj]], axis=-1)
            ok = np.array_equal(full, want) and n[0] == 19 * 23 and s[0] == want[..., 1].sum() \
                and mn[0] == want[..., 1].min() and mx[0] == want[..., 1].max()
            with open(out_path, "w") as f:
                f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gather_and_reduce_over_gloo(tmp_path):
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"
