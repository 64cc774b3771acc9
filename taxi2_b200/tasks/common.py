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


def iter_pair_blocks(engine, xs: list[Sequence], ys: list[Sequence] | None, align: bool, want_strings: bool,
                     scores, max_pairs: int = 1 << 20, raw_strings: bool = False):
    """Drive the device over the row-major product xs x ys (ys=None: xs x xs) in blocks of whole
    rows, yielding PairBlock in reference order.  One launch per block (plus one for strings).
    raw_strings: hand the gapped strings over as the library's arrays (for the native pair writer)
    instead of one Python string per sequence."""
    from ..engine import scores_vector

    same_set = ys is None
    ylist = xs if same_set else ys
    if align:
        engine.set_scores(scores_vector(dict(scores)) if scores is not None else None)
    engine.load([s.seq for s in xs], 0)
    if not same_set:
        engine.load([s.seq for s in ylist], 1)
    ny = len(ylist)
    if align and want_strings:
        max_pairs = min(max_pairs, 1 << 18)   # ~1.3 KB of gapped strings per barcode pair, twice
    rows = max(1, max_pairs // max(ny, 1))
    for x0 in range(0, len(xs), rows):
        nx = min(rows, len(xs) - x0)
        aligned = aligned_raw = None
        if align and want_strings:
            # one launch feeds both the distance files and aligned_pairs.txt (the reference aligns
            # each pair once, versus_all.py:527-552)
            px, py = np.divmod(np.arange(nx * ny, dtype=np.int64), ny)
            px, py = (px + x0).astype(np.int32), py.astype(np.int32)
            ox, oy, start, off, _, res = engine.align_strings_raw(px, py, want=("metrics",))
            metrics = res["metrics"].reshape(nx, ny, 4)
            if raw_strings:
                aligned_raw = (ox, oy, start, off)
            else:
                bx, by = ox.tobytes(), oy.tobytes()
                aligned = [(bx[int(start[k]): int(off[k + 1])].decode("latin-1"), by[int(start[k]): int(off[k + 1])].decode("latin-1"))
                           for k in range(nx * ny)]
        elif align:
            metrics = engine.align_rect(x0, nx, 0, ny, want=("metrics",))["metrics"]
        else:
            metrics = engine.count_rect(x0, nx, 0, ny, want=("metrics",))["metrics"]
        yield PairBlock(x0, nx, metrics, aligned, aligned_raw)
