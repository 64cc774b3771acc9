"""Shared pieces of the task classes."""
from __future__ import annotations

from enum import Enum
from math import isinf, isnan
from pathlib import Path
from typing import NamedTuple

import numpy as np

from ..distances import DistanceMetric
from ..sequences import Sequence


class Results(NamedTuple):
    output_directory: Path
    seconds_taken: float


def console_report(caption, index, total):
    """Default progress handler (versus_all.py:25-30)."""
    if caption == "Finalizing...":
        print(f"\rCalculating... {total}/{total} = {100:.2f}%", end="")
        print("\nFinalizing...")
    else:
        print(f"\rCalculating... {index}/{total} = {100 * index / total:.2f}%", end="")


class ComparisonType(Enum):
    """plot.py:15-27; only the labels are needed here (summary.tsv)."""

    Unknown = "no info"
    IntraSpecies = "intra-species"
    InterSpecies = "inter-species"
    IntraGenus = "intra-genus"
    InterGenus = "inter-genus"

    @property
    def label(self) -> str:
        return self.value


def create_parents(path: Path) -> None:
    if path.suffix:
        path = path.parent
    path.mkdir(parents=True, exist_ok=True)


def number_or_none(v: float) -> float | None:
    """DistanceMetric._is_number as a mapping (distances.py:290-292)."""
    v = float(v)
    return None if (isnan(v) or isinf(v)) else v


def metric_columns(metrics: list[DistanceMetric]) -> list[int]:
    cols = []
    for metric in metrics:
        if getattr(metric, "column", None) is None:
            raise NotImplementedError(f"metric {metric} is outside the B200 hot path (p, p-gaps, jc, k2p only)")
        cols.append(metric.column)
    return cols


class PairBlock(NamedTuple):
    """Results of one block of rows of a row-major pair product."""

    x0: int
    nx: int
    metrics: np.ndarray          # (nx, ny, 4) float64, NaN = undefined
    aligned: list | None         # nx*ny (aligned_x, aligned_y) strings, or None
    aligned_raw: tuple | None = None   # (aln_x, aln_y, start, off) arrays of Engine.align_strings_raw, or None
    extra: object = None         # whatever the caller's `extra(engine, block)` returned (computed on the block's GPU thread)


_multi_engines: dict = {}

# pairs per device block of iter_pair_blocks (whole rows); tests shrink it to force many blocks
MAX_BLOCK_PAIRS = 1 << 20
# versusAll without gapped strings keeps the whole n x n x 4 metric matrix on the host (the mirrored half of a
# tile belongs to rows that are written later) as long as it fits this many bytes; larger jobs align every
# ordered pair block by block as before
SYMMETRIC_MAX_BYTES = 4 << 30


def task_engine(task):
    """The process-wide MultiEngine for a task's device selection: `task.devices` (a list of CUDA
    device indices, or "all") when set, else the single `task.device`.  One context + one host
    thread per GPU (taxi2_b200/multi.py); raises without a GPU -- there is no CPU path."""
    from ..multi import MultiEngine

    devices = getattr(task, "devices", None)
    if devices == "all":
        key = "all"
        devices = None
    elif devices:
        key = tuple(int(d) for d in devices)
        devices = list(key)
    else:
        key = (int(task.device),)
        devices = list(key)
    eng = _multi_engines.get(key)
    if eng is None:
        eng = _multi_engines[key] = MultiEngine(devices)
    return eng


def iter_pair_blocks(engine, xs: list[Sequence], ys: list[Sequence] | None, align: bool, want_strings: bool,
                     scores, max_pairs: int | None = None, raw_strings: bool = False, extra=None):
    """Drive the device(s) over the row-major product xs x ys (ys=None: xs x xs) in blocks of whole
    rows, yielding PairBlock in reference order.  One launch per block: gapped strings and
    distances come from the same alignment (the reference aligns each pair once,
    versus_all.py:527-552).

    `engine` is a MultiEngine (or a single Engine, wrapped): blocks are dealt to the GPUs by the
    static LPT plan and computed ahead of the consumer (two finished blocks per GPU at most), so
    the caller's formatting / writing of block k overlaps the alignment of the next blocks.
    raw_strings: hand the gapped strings over as the library's arrays (for the native pair writer)
    instead of one Python string per sequence.  extra(engine, block): optional per-block work that
    needs the block's own context (runs on its GPU thread); its result is block.extra.  (On the
    one-alignment-per-unordered-pair route it runs before the block's metrics are complete and sees
    block.metrics = None: it is meant for work on the block's sequences, not on its results.)"""
    from ..engine import Engine, scores_vector
    from ..multi import MultiEngine

    if isinstance(engine, Engine):
        single = engine
        engine = MultiEngine.__new__(MultiEngine)
        engine.devices, engine.engines, engine.lens = [single.device], [single], [None, None]
    same_set = ys is None
    ylist = xs if same_set else ys
    if align:
        engine.set_scores(scores_vector(dict(scores)) if scores is not None else None)
    engine.load([s.seq for s in xs], 0)
    if not same_set:
        engine.load([s.seq for s in ylist], 1)
    ny = len(ylist)
    if max_pairs is None:
        max_pairs = MAX_BLOCK_PAIRS
    if align and want_strings:
        max_pairs = min(max_pairs, 1 << 18)   # ~1.3 KB of gapped strings per barcode pair, twice
    rows = max(1, max_pairs // max(ny, 1))
    if (same_set and align and not want_strings and ny >= 2 and 32 * ny * ny <= SYMMETRIC_MAX_BYTES
            and all(hasattr(e, "align_rect_both") for e in engine.engines)):
        # versus_all.py:746 aligns (x, y) and (y, x): one alignment per unordered pair serves both, the few
        # orientation-sensitive pairs are re-aligned (MultiEngine.iter_symmetric_rows); same blocks, same order
        on_diagonal = None
        if extra is not None:
            on_diagonal = lambda eng, x0, nx: extra(eng, PairBlock(x0, nx, None, None))   # noqa: E731
        for x0, nx, res in engine.iter_symmetric_rows(("metrics",), block=rows, max_cols=max(rows, max_pairs // rows),
                                                      on_diagonal=on_diagonal):
            yield PairBlock(x0, nx, res["metrics"][x0:x0 + nx], None, None, res["_extra"].pop(x0, None))
        return
    tiles = engine.row_tiles(rows)

    def compute(eng, tile, slot) -> PairBlock:
        x0, nx = tile.x0, tile.nx
        aligned = aligned_raw = None
        if align and want_strings:
            px, py = np.divmod(np.arange(nx * ny, dtype=np.int64), ny)
            px, py = (px + x0).astype(np.int32), py.astype(np.int32)
            # raw_strings consumers are done with a block's strings before they ask for the next block, so the
            # arrays can be the ones the third block before it used (no first touch of fresh memory per block);
            # the per-pair consumers get Python strings decoded from them right here
            ox, oy, start, off, _, res = eng.align_strings_raw(px, py, want=("metrics",), slot=slot)
            metrics = res["metrics"].reshape(nx, ny, 4)
            if raw_strings:
                aligned_raw = (ox, oy, start, off)
            else:
                bx, by = ox.tobytes(), oy.tobytes()
                aligned = [(bx[int(start[k]): int(off[k + 1])].decode("latin-1"), by[int(start[k]): int(off[k + 1])].decode("latin-1"))
                           for k in range(nx * ny)]
        elif align:
            metrics = eng.align_rect(x0, nx, 0, ny, want=("metrics",))["metrics"]
        else:
            metrics = eng.count_rect(x0, nx, 0, ny, want=("metrics",))["metrics"]
        block = PairBlock(x0, nx, metrics, aligned, aligned_raw)
        if extra is not None:
            block = block._replace(extra=extra(eng, block))
        return block

    for _, block in engine.run_tiles(tiles, compute, depth=2, ordered=True):
        yield block
