"""Dereplicate: greedy "keep the longest of every similar pair".

Mirrors /root/reference/src/itaxotools/taxi2/tasks/dereplicate.py (surface :107-138, pipeline
:393-440, greedy walk :306-351).  The reference filters its pair stream by a set that the walk
itself mutates downstream (:190-198, :336), so which pairs are ever aligned depends on earlier
results.  Here the distances come from the device in blocks of rows (a superset: every column not
yet excluded when the block starts) and the walk is REPLAYED on the host in the reference's order;
writer side effects (aligned pairs, distance files, summary) happen only for the pairs the
reference would have pulled.
"""
from __future__ import annotations

from itertools import groupby
from pathlib import Path
from time import perf_counter
from typing import Callable, NamedTuple

import numpy as np

from ..distances import Distance, DistanceHandler, DistanceMetric
from ..files import FileFormat, identify_format
from ..handlers import FileHandler
from ..pairs import SequencePair, SequencePairHandler
from ..sequences import Sequence, SequenceHandler, Sequences
from ..types import AttrDict
from .common import Results, console_report, create_parents, metric_columns, number_or_none, task_engine


class AllInfo(NamedTuple):
    query: Sequence
    id_x: str
    id_y: str
    len_x: int
    len_y: int
    distance: float
    similar: bool


class SummaryLine(NamedTuple):
    query_id: str
    query_length: str
    included_id: int
    included_length: int
    included_distance: float
    excluded_id: int
    excluded_length: int
    excluded_distance: float


def output_handler(fmt: FileFormat, path: Path):
    if fmt == FileFormat.Fasta:
        return SequenceHandler.Fasta(path, "w", write_organism=True)
    if fmt == FileFormat.Tabfile:
        return SequenceHandler.Tabfile(path, "w", idHeader="seqid", seqHeader="sequence")
    raise Exception("Unknown file format")


class Dereplicate:
    rows_per_block = 64

    def __init__(self):
        self.work_dir: Path = None
        self.paths = AttrDict()
        self.progress_handler: Callable = console_report
        self.progress_interval: float = 0.015
        self.device: int = 0
        self.devices = None   # list of CUDA device indices or "all": shard the pair product over several GPUs

        self.input: Sequences = None
        self.output_format: FileFormat = None
        self.excluded: set = set()

        self.params = AttrDict()
        self.params.thresholds = AttrDict(similarity=0.07, length=10)
        self.params.pairs = AttrDict(align=True, write=True, scores=None)
        self.params.distances = AttrDict(metric=None, write_linear=True, write_matricial=True)
        self.params.format = AttrDict(float="{:.4f}", missing="NA", percentage_multiply=False)

    def set_output_format_from_path(self, path: Path):
        self.output_format = identify_format(path)

    def check_params(self):
        self.output_format = self.output_format or FileFormat.Tabfile
        self.params.distances.metric = self.params.distances.metric or DistanceMetric.Uncorrected()

    def generate_paths(self):
        assert self.work_dir
        w = Path(self.work_dir)
        create_parents(w)
        metric, ext = str(self.params.distances.metric), self.output_format.extension
        self.paths.summary = w / "summary.tsv"
        self.paths.dereplicated = w / f"dereplicated{ext}"
        self.paths.excluded = w / f"excluded{ext}"
        self.paths.aligned_pairs = w / "aligned_pairs.txt"
        self.paths.distances_linear = w / "distances" / f"{metric}.linear.tsv"
        self.paths.distances_matricial = w / "distances" / f"{metric}.matricial.tsv"

    def start(self) -> Results:
        from ..engine import scores_vector

        ts = perf_counter()
        self.excluded = set()
        self.check_params()
        self.generate_paths()
        p = self.params
        metric = p.distances.metric
        (col,) = metric_columns([metric])
        fmt, missing = p.format.float, p.format.missing
        scale = 100.0 if p.format.percentage_multiply else 1.0
        similarity = p.thresholds.similarity

        data = [s for s in self.input if len(s.seq) >= p.thresholds.length]   # raw strings, gaps included
        work = [s.normalize() for s in data] if p.pairs.align else data
        n = len(data)
        engine = task_engine(self).engines[0]
        if p.pairs.align:
            engine.set_scores(scores_vector(dict(p.pairs.scores)) if p.pairs.scores is not None else None)
        engine.load([s.seq for s in work], 0)

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
        summary = FileHandler.Tabfile(self.paths.summary, "w", columns=SummaryLine._fields)
        writers.append(summary)

        total = n * n
        state = dict(done=0, last=perf_counter())

        def infos():
            """The reference's stream of visited pairs, in order, with the exclusion set consulted
            at the moment each pair would have been pulled."""
            for x0 in range(0, n, self.rows_per_block):
                rows = [i for i in range(x0, min(n, x0 + self.rows_per_block)) if data[i].id not in self.excluded]
                cols = [j for j in range(n) if data[j].id not in self.excluded]
                if not rows or not cols:
                    continue
                px = np.repeat(np.asarray(rows, dtype=np.int32), len(cols))
                py = np.tile(np.asarray(cols, dtype=np.int32), len(rows))
                if p.pairs.align:
                    values = engine.align_pairs(px, py, want=("metrics",))["metrics"][:, col]
                else:
                    values = engine.count_pairs(px, py, want=("metrics",))["metrics"][:, col]
                col_pos = {j: k for k, j in enumerate(cols)}
                for r, i in enumerate(rows):
                    x = data[i]
                    if x.id in self.excluded:
                        continue
                    row_strings = None
                    if pairs_file is not None:   # one launch for the whole row's candidate columns
                        ax, ay, _ = engine.align_strings(np.full(len(cols), i, dtype=np.int32), np.asarray(cols, dtype=np.int32))
                        row_strings = (ax, ay)
                    for j in range(n):
                        y = data[j]
                        if x.id == y.id or x.id in self.excluded or y.id in self.excluded:
                            continue
                        k = r * len(cols) + col_pos[j]
                        if row_strings is not None:
                            pair = SequencePair(Sequence(x.id, row_strings[0][col_pos[j]].decode("latin-1"), work[i].extras),
                                                Sequence(y.id, row_strings[1][col_pos[j]].decode("latin-1"), work[j].extras))
                            pairs_file.write(pair)
                        else:
                            pair = SequencePair(work[i], work[j])
                        d = number_or_none(values[k])
                        state["done"] += 1
                        now = perf_counter()
                        if now - state["last"] >= self.progress_interval:
                            self.progress_handler("distance.x.id", state["done"], total - len(self.excluded) * n)
                            state["last"] = now
                        if d is not None:
                            d *= scale
                        distance = Distance(metric, pair.x, pair.y, d)
                        if linear_file:
                            linear_file.write(distance)
                        if matrix_file:
                            matrix_file.write(distance)
                        similar = False if d is None else bool(d <= similarity)
                        yield AllInfo(x, x.id, y.id, len(x.seq), len(y.seq), d, similar)
            self.progress_handler("Finalizing...", total, total)

        text = lambda d: missing if d is None else fmt.format(d)  # noqa: E731
        try:
            for _, group in groupby(infos(), lambda info: info.id_x):
                first = None
                for info in group:
                    if first is None:
                        first = info
                        query_id, query_length = first.id_x, first.len_x
                        max_id, max_length, max_distance = first.id_x, first.len_x, first.distance
                    _, _, id_y, _, len_y, distance, similar = info
                    if not similar:
                        continue
                    if len_y > max_length:
                        inc, exc = (id_y, len_y, distance), (max_id, max_length, max_distance)
                    else:
                        inc, exc = (max_id, max_length, max_distance), (id_y, len_y, distance)
                    self.excluded.add(exc[0])
                    summary.write((query_id, str(query_length), inc[0], str(inc[1]), text(inc[2]), exc[0], str(exc[1]), text(exc[2])))
                    if len_y > max_length:
                        max_id, max_length, max_distance = id_y, len_y, distance
        finally:
            for w in writers:
                w.close()

        with output_handler(self.output_format, self.paths.dereplicated) as kept, \
                output_handler(self.output_format, self.paths.excluded) as dropped:
            for sequence in data:
                (dropped if sequence.id in self.excluded else kept).write(sequence)
        return Results(self.work_dir, perf_counter() - ts)
