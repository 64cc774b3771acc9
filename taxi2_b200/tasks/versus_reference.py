"""VersusReference: every data sequence against every reference sequence; the closest reference
per query (first minimum of the main metric) with the extra metrics of that winning pair.

Mirrors /root/reference/src/itaxotools/taxi2/tasks/versus_reference.py (surface :33-62,
pipeline :213-247).  The main metric of all pairs comes from tile-wise device launches; the
per-query first minimum follows the reference's groupby/min semantics (:184-188) including the
ValueError when a query has no defined distance.
"""
from __future__ import annotations

from itertools import groupby
from pathlib import Path
from time import perf_counter
from typing import Callable

import numpy as np

from .. import fastwrite
from ..distances import Distance, DistanceHandler, DistanceMetric
from ..pairs import SequencePair, SequencePairHandler
from ..sequences import Sequence, Sequences
from ..types import AttrDict
from .common import Results, console_report, create_parents, iter_pair_blocks, metric_columns, number_or_none, task_engine


class VersusReference:
    def __init__(self):
        self.work_dir: Path = None
        self.paths = AttrDict()
        self.progress_handler: Callable = console_report
        self.progress_interval: float = 0.015
        self.device: int = 0
        self.devices = None   # list of CUDA device indices or "all": shard the pair product over several GPUs
        self.native_writers: bool = True   # batch formatter for plain float formats (same bytes as the handlers)

        self.input = AttrDict()
        self.input.data: Sequences = None
        self.input.reference: Sequences = None

        self.params = AttrDict()
        self.params.pairs = AttrDict(align=True, write=True, scores=None)
        self.params.distances = AttrDict(metric=None, extra_metrics=None, write_linear=True, write_matricial=True)
        self.params.format = AttrDict(float="{:.4f}", percentage="{:.2f}", missing="NA", percentage_multiply=False)

    def generate_paths(self):
        assert self.work_dir
        w = Path(self.work_dir)
        create_parents(w)
        metric = str(self.params.distances.metric)
        self.paths.closest = w / "closest.tsv"
        self.paths.aligned_pairs = w / "aligned_pairs.txt"
        self.paths.distances_linear = w / "distances" / f"{metric}.linear.tsv"
        self.paths.distances_matricial = w / "distances" / f"{metric}.matricial.tsv"

    def check_metrics(self):
        d = self.params.distances
        d.metric = d.metric or DistanceMetric.Uncorrected()
        d.extra_metrics = d.extra_metrics or [DistanceMetric.UncorrectedWithGaps(), DistanceMetric.JukesCantor(), DistanceMetric.Kimura2P()]
        if d.metric in d.extra_metrics:
            d.extra_metrics.remove(d.metric)

    def start(self) -> Results:
        ts = perf_counter()
        self.check_metrics()
        self.generate_paths()
        p = self.params
        main, extras = p.distances.metric, p.distances.extra_metrics
        (main_col,), extra_cols = metric_columns([main]), metric_columns(extras)
        fmt, missing = p.format.float, p.format.missing
        scale = 100.0 if p.format.percentage_multiply else 1.0

        data = list(self.input.data.normalize() if p.pairs.align else self.input.data)
        reference = list(self.input.reference.normalize() if p.pairs.align else self.input.reference)
        nref = len(reference)
        engine = task_engine(self)

        writers = []
        pairs_file = linear_file = matrix_file = None
        if p.pairs.align and p.pairs.write:
            create_parents(self.paths.aligned_pairs)
            pairs_file = SequencePairHandler.Formatted(self.paths.aligned_pairs, "w")
            writers.append(pairs_file)
        if p.distances.write_linear:
            create_parents(self.paths.distances_linear)
            linear_file = DistanceHandler.Linear.WithExtras(self.paths.distances_linear, "w", missing=missing, formatter=fmt)
            writers.append(linear_file)
        if p.distances.write_matricial:
            create_parents(self.paths.distances_matricial)
            matrix_file = DistanceHandler.Matrix(self.paths.distances_matricial, "w", missing=missing, formatter=fmt)
            writers.append(matrix_file)
        create_parents(self.paths.closest)
        closest_file = DistanceHandler.Linear.WithExtras(self.paths.closest, "w", missing=missing, formatter=fmt)
        writers.append(closest_file)

        total = len(data) * nref
        state = dict(done=0, last=perf_counter())

        fmtc = fastwrite.printf_format(fmt) if self.native_writers else None
        fill = lambda values: [missing if v is None else v for v in values]  # noqa: E731
        if fmtc:
            t_x = fastwrite.StringTable(["\t".join([s.id, *fill(s.extras.values())]) for s in data])
            t_y = fastwrite.StringTable(["\t".join([s.id, *fill(s.extras.values())]) for s in reference])
            t_xid = fastwrite.StringTable([s.id for s in data])
        native_started = False

        def native_block(block):
            """Append the block's rows of the linear / matrix files through the batch formatter."""
            nonlocal native_started
            if not native_started:
                native_started = True
                for handle, path, header in (
                        (linear_file, self.paths.distances_linear,
                         ("seqid (query)", *(k + " (query)" for k in data[0].extras), "seqid (reference)",
                          *(k + " (reference)" for k in reference[0].extras), str(main))),
                        (matrix_file, self.paths.distances_matricial, ("", *(s.id for s in reference)))):
                    if handle is not None:
                        handle.close()
                        path.write_text("\t".join(header) + "\n")
            if linear_file is not None:
                fastwrite.format_pairs(self.paths.distances_linear, [fastwrite.SEG_X[0], fastwrite.SEG_Y[0], fastwrite.SEG_SCORES],
                                       [t_x], [t_y], block.x0, block.nx, nref, block.metrics, None, [main_col], scale, fmtc, missing)
            if matrix_file is not None:
                fastwrite.format_matrix(self.paths.distances_matricial, t_xid, block.x0, block.nx, nref, block.metrics, None, main_col,
                                        scale, fmtc, missing)

        def main_distances():
            """(Distance of the main metric, all four raw metrics of the pair) in product order."""
            for block in iter_pair_blocks(engine, data, reference, p.pairs.align, pairs_file is not None, p.pairs.scores):
                if fmtc and data and reference:
                    native_block(block)
                for bx in range(block.nx):
                    x = data[block.x0 + bx]
                    for j, y in enumerate(reference):
                        if block.aligned is not None:
                            ax, ay = block.aligned[bx * nref + j]
                            pair = SequencePair(Sequence(x.id, ax, x.extras), Sequence(y.id, ay, y.extras))
                            pairs_file.write(pair)
                        else:
                            pair = SequencePair(x, y)
                        d = number_or_none(block.metrics[bx, j, main_col])
                        state["done"] += 1
                        now = perf_counter()
                        if now - state["last"] >= self.progress_interval:
                            self.progress_handler("distance.x.id", state["done"], total)
                            state["last"] = now
                        if d is not None:
                            d *= scale
                        distance = Distance(main, pair.x, pair.y, d)
                        if not fmtc:
                            if linear_file:
                                linear_file.write(distance)
                            if matrix_file:
                                matrix_file.write(distance)
                        yield distance, block.metrics[bx, j]
            self.progress_handler("Finalizing...", total, total)

        def write_closest(best: Distance, raw) -> None:
            closest_file.write(best)
            for metric, col in zip(extras, extra_cols):
                d = number_or_none(raw[col])
                if d is not None:
                    d *= scale   # adjust_extra_distances: only the non-main metrics are scaled here
                closest_file.write(Distance(metric, best.x, best.y, d))

        def run_blocks_natively() -> None:
            """No per-pair Python: linear / matrix rows and aligned pairs through the batch writers,
            the closest reference of a query as the first minimum of its row (NaN = undefined)."""
            native_pairs = None
            if pairs_file is not None:
                pairs_file.close()
                Path(self.paths.aligned_pairs).write_bytes(b"")
                native_pairs = (t_xid, fastwrite.StringTable([s.id for s in reference]))
            first = True
            for block in iter_pair_blocks(engine, data, reference, p.pairs.align, native_pairs is not None, p.pairs.scores,
                                          raw_strings=True):
                native_block(block)
                if native_pairs is not None:
                    fastwrite.format_aligned_pairs(self.paths.aligned_pairs, first, *native_pairs, block.x0, block.nx, nref, *block.aligned_raw)
                    first = False
                column = block.metrics[:, :, main_col]
                undefined = ~np.isfinite(column)
                for bx in range(block.nx):
                    if undefined[bx].all():
                        raise ValueError("min() arg is an empty sequence")   # the reference's min() over no defined distance
                    j = int(np.argmin(np.where(undefined[bx], np.inf, column[bx])))
                    x, y = data[block.x0 + bx], reference[j]
                    if block.aligned_raw is not None:
                        ox, oy, start, off = block.aligned_raw
                        k = bx * nref + j
                        x = Sequence(x.id, ox[start[k]:off[k + 1]].tobytes().decode("latin-1"), x.extras)
                        y = Sequence(y.id, oy[start[k]:off[k + 1]].tobytes().decode("latin-1"), y.extras)
                    write_closest(Distance(main, x, y, float(column[bx, j]) * scale), block.metrics[bx, j])
                state["done"] += block.nx * nref
                now = perf_counter()
                if now - state["last"] >= self.progress_interval:
                    self.progress_handler("distance.x.id", state["done"], total)
                    state["last"] = now
            self.progress_handler("Finalizing...", total, total)

        try:
            # groupby(x.id) merges consecutive queries that share an id into one group (reference quirk):
            # such inputs take the per-pair path, which reproduces it
            distinct = all(a.id != b.id for a, b in zip(data, data[1:]))
            if fmtc and data and reference and distinct:
                run_blocks_natively()
                return Results(self.work_dir, perf_counter() - ts)
            for _, group in groupby(main_distances(), lambda item: item[0].x.id):
                defined = [item for item in group if item[0].d is not None]
                best, raw = min(defined, key=lambda item: item[0].d)   # ValueError on an empty group, like the reference
                write_closest(best, raw)
        finally:
            for w in writers:
                w.close()
        return Results(self.work_dir, perf_counter() - ts)
