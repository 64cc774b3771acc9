"""Decontaminate / Decontaminate2: flag query sequences that are too close to an outgroup
(or closer to the outgroup than to an ingroup).

Mirrors /root/reference/src/itaxotools/taxi2/tasks/decontaminate.py (:95-371) and
decontaminate2.py (:99-434): data x outgroup (and x ingroup) alignment, one metric, per-query
minimum with undefined distances treated as +inf (:258-264), threshold `d <= similarity`
(:273-276) or weighted `out < in` (decontaminate2.py:314-330; ingroup distances are never
percentage-scaled, :417-419).
"""
from __future__ import annotations

from itertools import groupby
from math import inf
from pathlib import Path
from time import perf_counter
from typing import Callable

import numpy as np

from .. import fastwrite
from ..distances import Distance, DistanceHandler, DistanceMetric
from ..files import FileFormat, identify_format
from ..handlers import FileHandler
from ..pairs import SequencePair, SequencePairHandler
from ..sequences import Sequence, Sequences
from ..types import AttrDict
from .common import Results, console_report, create_parents, iter_pair_blocks, metric_columns, number_or_none, task_engine
from .dereplicate import output_handler


class _DecontaminateBase:
    def __init__(self):
        self.work_dir: Path = None
        self.paths = AttrDict()
        self.progress_handler: Callable = console_report
        self.progress_interval: float = 0.015
        self.device: int = 0
        self.devices = None   # list of CUDA device indices or "all": shard the pair product over several GPUs
        self.native_writers: bool = True   # batch writers and row minima without per-pair Python (same bytes)
        self.input: Sequences = None
        self.outgroup: Sequences = None
        self.output_format: FileFormat = None
        self.params = AttrDict()
        self.params.pairs = AttrDict(align=True, write=True, scores=None)
        self.params.distances = AttrDict(metric=None, write_linear=True, write_matricial=True)
        self.params.format = AttrDict(float="{:.4f}", missing="NA", percentage_multiply=False)

    def set_output_format_from_path(self, path: Path):
        self.output_format = identify_format(path)

    def check_params(self):
        self.output_format = self.output_format or FileFormat.Tabfile
        self.params.distances.metric = self.params.distances.metric or DistanceMetric.Uncorrected()

    def _minimums(self, data, group, pairs_path, linear_path, matrix_path, scale: float):
        """Per-query minimum Distance of data x group, writing the pair / distance files on the way
        (generator over queries, in order)."""
        p = self.params
        metric = p.distances.metric
        (col,) = metric_columns([metric])
        fmt, missing = p.format.float, p.format.missing
        xs = list(data.normalize() if p.pairs.align else data)
        ys = list(group.normalize() if p.pairs.align else group)
        engine = task_engine(self)
        writers = []
        pairs_file = linear_file = matrix_file = None
        if p.pairs.align and p.pairs.write:
            create_parents(pairs_path)
            pairs_file = SequencePairHandler.Formatted(pairs_path, "w")
            writers.append(pairs_file)
        if p.distances.write_linear:
            create_parents(linear_path)
            linear_file = DistanceHandler.Linear.WithExtras(linear_path, "w", missing=missing, formatter=fmt)
            writers.append(linear_file)
        if p.distances.write_matricial:
            create_parents(matrix_path)
            matrix_file = DistanceHandler.Matrix(matrix_path, "w", missing=missing, formatter=fmt)
            writers.append(matrix_file)

        def distances():
            for block in iter_pair_blocks(engine, xs, ys, p.pairs.align, pairs_file is not None, p.pairs.scores):
                for bx in range(block.nx):
                    x = xs[block.x0 + bx]
                    for j, y in enumerate(ys):
                        if block.aligned is not None:
                            ax, ay = block.aligned[bx * len(ys) + j]
                            pair = SequencePair(Sequence(x.id, ax, x.extras), Sequence(y.id, ay, y.extras))
                            pairs_file.write(pair)
                        else:
                            pair = SequencePair(x, y)
                        d = number_or_none(block.metrics[bx, j, col])
                        if d is not None:
                            d *= scale
                        distance = Distance(metric, pair.x, pair.y, d)
                        if linear_file:
                            linear_file.write(distance)
                        if matrix_file:
                            matrix_file.write(distance)
                        yield distance

        def minimums_by_block():
            """The same stream without per-pair Python: rows of the linear / matrix files and the aligned
            pairs through the batch writers, the minimum of a query as the first minimum of its row
            (undefined counts as +inf; a row with nothing defined yields its first pair, like min())."""
            fill = lambda values: [missing if v is None else v for v in values]  # noqa: E731
            t_x = fastwrite.StringTable(["\t".join([s.id, *fill(s.extras.values())]) for s in xs])
            t_y = fastwrite.StringTable(["\t".join([s.id, *fill(s.extras.values())]) for s in ys])
            t_xid = fastwrite.StringTable([s.id for s in xs])
            t_yid = fastwrite.StringTable([s.id for s in ys])
            for handle, path, header in (
                    (linear_file, linear_path, ("seqid (query)", *(k + " (query)" for k in xs[0].extras), "seqid (reference)",
                                                *(k + " (reference)" for k in ys[0].extras), str(metric))),
                    (matrix_file, matrix_path, ("", *(s.id for s in ys))),
                    (pairs_file, pairs_path, None)):
                if handle is not None:
                    handle.close()
                    Path(path).write_text("" if header is None else "\t".join(header) + "\n")
            first = True
            for block in iter_pair_blocks(engine, xs, ys, p.pairs.align, pairs_file is not None, p.pairs.scores, raw_strings=True):
                if linear_file is not None:
                    fastwrite.format_pairs(linear_path, [fastwrite.SEG_X[0], fastwrite.SEG_Y[0], fastwrite.SEG_SCORES], [t_x], [t_y],
                                           block.x0, block.nx, len(ys), block.metrics, None, [col], scale, fmtc, missing)
                if matrix_file is not None:
                    fastwrite.format_matrix(matrix_path, t_xid, block.x0, block.nx, len(ys), block.metrics, None, col, scale, fmtc, missing)
                if pairs_file is not None:
                    fastwrite.format_aligned_pairs(pairs_path, first, t_xid, t_yid, block.x0, block.nx, len(ys), *block.aligned_raw)
                    first = False
                column = block.metrics[:, :, col]
                undefined = ~np.isfinite(column)
                for bx in range(block.nx):
                    j = int(np.argmin(np.where(undefined[bx], np.inf, column[bx])))
                    d = None if undefined[bx, j] else float(column[bx, j]) * scale
                    yield Distance(metric, xs[block.x0 + bx], ys[j], d)

        fmtc = fastwrite.printf_format(fmt) if self.native_writers else None
        distinct = all(a.id != b.id for a, b in zip(xs, xs[1:]))   # groupby(x.id) would merge equal neighbours
        try:
            if fmtc and xs and ys and distinct:
                yield from minimums_by_block()
                return
            for _, grp in groupby(distances(), lambda d: d.x.id):
                yield min(grp, key=lambda d: d.d if d.d is not None else inf)
        finally:
            for w in writers:
                w.close()

    def _finish(self, verdict_lines, header):
        """Write summary + the two sequence files from (sequence, contaminant, summary row) triples."""
        p = self.params
        total = len(self.input)
        last = perf_counter()
        with FileHandler.Tabfile(self.paths.summary, "w", columns=header) as summary, \
                output_handler(self.output_format, self.paths.decontaminated) as clean, \
                output_handler(self.output_format, self.paths.contaminants) as dirty:
            for index, (sequence, contaminant, row) in enumerate(verdict_lines, 1):
                (dirty if contaminant else clean).write(sequence)
                summary.write(row)
                now = perf_counter()
                if now - last >= self.progress_interval:
                    self.progress_handler("verdict.x.id", index, total)
                    last = now
        self.progress_handler("Finalizing...", total, total)


class Decontaminate(_DecontaminateBase):
    def __init__(self):
        super().__init__()
        self.params.thresholds = AttrDict(similarity=0.07)

    def generate_paths(self):
        assert self.work_dir
        w = Path(self.work_dir)
        create_parents(w)
        metric, ext = str(self.params.distances.metric), self.output_format.extension
        self.paths.summary = w / "summary.tsv"
        self.paths.decontaminated = w / f"decontaminated{ext}"
        self.paths.contaminants = w / f"contaminants{ext}"
        self.paths.aligned_pairs = w / "aligned_pairs.txt"
        self.paths.distances_linear = w / "distances" / f"{metric}.linear.tsv"
        self.paths.distances_matrix = w / "distances" / f"{metric}.matricial.tsv"

    def start(self) -> Results:
        ts = perf_counter()
        self.check_params()
        self.generate_paths()
        p = self.params
        scale = 100.0 if p.format.percentage_multiply else 1.0
        text = lambda d: p.format.missing if d is None else p.format.float.format(d)  # noqa: E731
        threshold = p.thresholds.similarity
        minimums = self._minimums(self.input, self.outgroup, self.paths.aligned_pairs, self.paths.distances_linear,
                                  self.paths.distances_matrix, scale)

        def verdicts():
            for sequence, best in zip(self.input, minimums):
                contaminant = False if best.d is None else bool(best.d <= threshold)
                yield sequence, contaminant, (sequence.id, best.y.id, text(best.d), "Yes" if contaminant else "No")

        self._finish(verdicts(), ("query_id", "outgroup_id", "outgroup_distance", "contaminant"))
        return Results(self.work_dir, perf_counter() - ts)


class Decontaminate2(_DecontaminateBase):
    def __init__(self):
        super().__init__()
        self.ingroup: Sequences = None
        self.params.weights = AttrDict(outgroup=1.0, ingroup=1.0)

    def generate_paths(self):
        assert self.work_dir
        w = Path(self.work_dir)
        create_parents(w)
        metric, ext = str(self.params.distances.metric), self.output_format.extension
        self.paths.summary = w / "summary.tsv"
        self.paths.decontaminated = w / f"decontaminated{ext}"
        self.paths.contaminants = w / f"contaminants{ext}"
        self.paths.outgroup_aligned_pairs = w / "aligned_pairs" / "outgroup.txt"
        self.paths.ingroup_aligned_pairs = w / "aligned_pairs" / "ingroup.txt"
        self.paths.outgroup_linear = w / "distances" / f"outgroup.{metric}.linear.tsv"
        self.paths.outgroup_matrix = w / "distances" / f"outgroup.{metric}.matricial.tsv"
        self.paths.ingroup_linear = w / "distances" / f"ingroup.{metric}.linear.tsv"
        self.paths.ingroup_matrix = w / "distances" / f"ingroup.{metric}.matricial.tsv"

    def start(self) -> Results:
        ts = perf_counter()
        self.check_params()
        self.generate_paths()
        p = self.params
        scale = 100.0 if p.format.percentage_multiply else 1.0
        text = lambda d: p.format.missing if d is None else p.format.float.format(d)  # noqa: E731
        out_min = self._minimums(self.input, self.outgroup, self.paths.outgroup_aligned_pairs, self.paths.outgroup_linear,
                                 self.paths.outgroup_matrix, scale)
        # the two products cannot share the device context's loaded sets lazily: materialise the
        # outgroup minima first (the files they write are independent of the ingroup files)
        out_min = list(out_min)
        in_min = self._minimums(self.input, self.ingroup, self.paths.ingroup_aligned_pairs, self.paths.ingroup_linear,
                                self.paths.ingroup_matrix, 1.0)
        w_out, w_in = p.weights.outgroup, p.weights.ingroup

        def verdicts():
            for sequence, best_out, best_in in zip(self.input, out_min, in_min):
                d_out = None if best_out.d is None else best_out.d * w_out
                d_in = None if best_in.d is None else best_in.d * w_in
                if d_out is None:
                    contaminant = False
                elif d_in is None:
                    contaminant = True
                else:
                    contaminant = bool(d_out < d_in)
                yield sequence, contaminant, (sequence.id, best_out.y.id, text(d_out), best_in.y.id, text(d_in),
                                              "Yes" if contaminant else "No")

        self._finish(verdicts(), ("query_id", "outgroup_id", "outgroup_distance", "ingroup_id", "ingroup_distance", "contaminant"))
        return Results(self.work_dir, perf_counter() - ts)
