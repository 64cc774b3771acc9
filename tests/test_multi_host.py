"""Host logic of the multi-GPU product path on CPU: the tile scheduler of MultiEngine with stand-in
engines (ordering, bounded run-ahead, error propagation, LPT ownership) and the host-side
combination of per-column-tile best matches (first minimum, versus_reference.py:184-188)."""
from __future__ import annotations

import threading
import time

import numpy as np
import pytest

from taxi2_b200.multi import MultiEngine, combine_best
from taxi2_b200.sharding import assign_tiles, make_tiles


class FakeEngine:
    def __init__(self, device):
        self.device = device
        self.n = [0, 0]
        self.calls = []


def fake_multi(ndev: int, lens_x, lens_y=None) -> MultiEngine:
    m = MultiEngine.__new__(MultiEngine)
    m.devices = list(range(ndev))
    m.engines = [FakeEngine(d) for d in range(ndev)]
    m.lens = [np.asarray(lens_x), None if lens_y is None else np.asarray(lens_y)]
    for e in m.engines:
        e.n = [len(lens_x), 0 if lens_y is None else len(lens_y)]
        e.ny = len(lens_x) if lens_y is None else len(lens_y)
    return m


@pytest.mark.parametrize("ndev", [1, 2, 3, 8])
def test_tiles_run_on_their_planned_gpu_and_arrive_in_order(ndev):
    rng = np.random.default_rng(ndev)
    m = fake_multi(ndev, rng.integers(300, 1500, 101))
    tiles = m.row_tiles(rows_per_tile=7)
    assert sum(t.nx for t in tiles) == 101 and all(t.y0 == 0 and t.ny == 101 for t in tiles)
    plan = assign_tiles(tiles, ndev)
    owner = {t.index: k for k, mine in enumerate(plan) for t in mine}

    def fn(engine, tile, slot):
        time.sleep(0.001 * ((tile.index * 7) % 5))
        engine.calls.append(tile.index)
        return (engine.device, tile.index, slot)

    got = list(m.run_tiles(tiles, fn, depth=2, ordered=True))
    assert [t.index for t, _ in got] == list(range(len(tiles)))
    for t, (dev, idx, slot) in got:
        assert dev == owner[t.index] and idx == t.index and 0 <= slot <= 2
    for k, e in enumerate(m.engines):   # each GPU walks its share in row order
        assert e.calls == [t.index for t in plan[k]]
    # LPT keeps the cell load balanced
    loads = [sum(t.cells for t in mine) for mine in plan]
    assert max(loads) <= sum(loads) / ndev + max(t.cells for t in tiles)


def test_run_ahead_is_bounded_by_depth():
    m = fake_multi(2, np.full(40, 100))
    tiles = m.row_tiles(rows_per_tile=1)
    outstanding = [0, 0]
    peak = [0, 0]
    lock = threading.Lock()

    def fn(engine, tile, slot):
        with lock:
            outstanding[engine.device] += 1
            peak[engine.device] = max(peak[engine.device], outstanding[engine.device])
        return engine.device

    for tile, dev in m.run_tiles(tiles, fn, depth=3, ordered=True):
        time.sleep(0.002)          # a slow consumer: the GPUs must wait, not pile results up
        with lock:
            outstanding[dev] -= 1
    assert max(peak) <= 3 + 1      # `depth` finished tiles + the one being handed over


def test_worker_errors_reach_the_consumer_and_threads_stop():
    m = fake_multi(2, np.full(30, 100))
    tiles = m.row_tiles(rows_per_tile=1)

    def fn(engine, tile, slot):
        if tile.index == 11:
            raise RuntimeError("boom")
        return tile.index

    before = threading.active_count()
    with pytest.raises(RuntimeError, match="boom"):
        list(m.run_tiles(tiles, fn))
    assert threading.active_count() == before
    # an abandoned iteration also winds its threads down
    it = m.run_tiles(tiles, lambda e, t, s: t.index)
    next(it)
    it.close()
    assert threading.active_count() == before


def test_unordered_yields_everything_once():
    m = fake_multi(4, np.arange(10, 73))
    tiles = m.row_tiles(rows_per_tile=2, col_tiles=3)
    assert len({(t.x0, t.y0) for t in tiles}) == len(tiles) and max(t.y0 + t.ny for t in tiles) == 63
    got = sorted(t.index for t, _ in m.run_tiles(tiles, lambda e, t, s: None, ordered=False))
    assert got == list(range(len(tiles)))


def test_default_tiling_gives_every_gpu_several_tiles():
    m = fake_multi(8, np.full(200_000, 650), np.full(20_000, 650))   # BASELINE C4 geometry
    tiles = m.row_tiles()
    assert all(t.ny == 20_000 for t in tiles) and sum(t.nx for t in tiles) == 200_000
    plan = assign_tiles(tiles, 8)
    assert min(len(p) for p in plan) >= 4
    loads = [sum(t.cells for t in p) for p in plan]
    assert max(loads) / (sum(loads) / 8) < 1.01


def test_combine_best_is_the_first_minimum_in_reference_order():
    rng = np.random.default_rng(3)
    nx, ny, metric = 200, 37, 2
    d = rng.integers(0, 6, (nx, ny, 4)).astype(np.float64) / 4     # many ties
    d[rng.random((nx, ny)) < 0.3] = np.nan
    d[:5] = np.nan                                                  # queries with no defined distance
    counts = rng.integers(0, 100, (nx, ny, 4)).astype(np.int32)
    want = np.full(nx, -1)
    for i in range(nx):
        best = np.inf
        for j in range(ny):
            v = d[i, j, metric]
            if not np.isnan(v) and v < best:
                best, want[i] = v, j
    bounds = [0, 5, 6, 20, 37]
    for order in ([0, 1, 2, 3], [3, 1, 0, 2], [2, 3, 1, 0]):
        index = np.full(nx, -1, dtype=np.int32)
        best = np.full((nx, 4), np.nan)
        cnt = np.zeros((nx, 4), dtype=np.int32)
        for k in order:
            lo, hi = bounds[k], bounds[k + 1]
            sub = d[:, lo:hi]
            col = np.where(np.isnan(sub[..., metric]), np.inf, sub[..., metric])
            arg = col.argmin(axis=1)
            none = np.isinf(col.min(axis=1))
            t_index = np.where(none, -1, arg + lo).astype(np.int32)
            t_best = np.where(none[:, None], np.nan, sub[np.arange(nx), arg])
            t_cnt = np.where(none[:, None], 0, counts[np.arange(nx), arg + lo]).astype(np.int32)
            combine_best(index, best, cnt, t_index, t_best, t_cnt, metric)
        assert np.array_equal(index, want), order
        ok = want >= 0
        assert np.array_equal(best[ok], d[np.arange(nx)[ok], want[ok]], equal_nan=True)
        assert np.array_equal(cnt[ok], counts[np.arange(nx)[ok], want[ok]])
        assert np.isnan(best[~ok]).all()


@pytest.mark.parametrize("gpus,n,block,max_cols", [(1, 37, 8, None), (3, 101, None, None), (2, 64, 16, 16), (4, 5, 2048, None), (1, 1, None, None)])
def test_symmetric_tiles_cover_every_unordered_pair_once(gpus, n, block, max_cols):
    """The upper-triangle tiling behind versusAll's one-alignment-per-unordered-pair route: diagonal
    squares plus rectangles strictly to their right cover each unordered pair exactly once, in
    row-block-major order, and the symmetric matrix assembled from oracle-backed engines equals the
    matrix of every ordered pair."""
    from fake_engine import oracle_multi

    rng = np.random.default_rng(n)
    seqs = ["".join(rng.choice(list("ACGT"), size=int(rng.integers(5, 12)))) for _ in range(n)]
    multi = oracle_multi(gpus)
    multi.load(seqs, 0)
    tiles = multi.symmetric_tiles(block, max_cols)
    seen = np.zeros((n, n), dtype=np.int32)
    for t in tiles:
        assert t.y0 >= t.x0 and (t.y0 == t.x0 and t.nx == t.ny or t.y0 >= t.x0 + t.nx)
        seen[t.x0:t.x0 + t.nx, t.y0:t.y0 + t.ny] += 1
        if t.y0 != t.x0:
            seen[t.y0:t.y0 + t.ny, t.x0:t.x0 + t.nx] += 1      # the mirror image the rectangle fills as well
    assert (seen == 1).all()
    assert [t.index for t in tiles] == list(range(len(tiles))) and [t.x0 for t in tiles] == sorted(t.x0 for t in tiles)
    want = multi.engines[0].align_rect(0, n, 0, n)
    got = multi.align_matrix_symmetric(block=block, max_cols=max_cols)
    for key in ("score", "counts"):
        assert np.array_equal(got[key], want[key])
    assert np.array_equal(got["metrics"], want["metrics"], equal_nan=True)
