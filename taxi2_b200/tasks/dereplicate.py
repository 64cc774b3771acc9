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
    rows_per_block = None   # rows per device block (None: about 1M pairs in a block's first chunk of columns)
    first_columns = None    # columns past its own end a block is aligned against before a row asks for more (None: 2048)

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
        """Block replay of the greedy walk (dereplicate.py:306-351).

        The device is a distance oracle for (block of rows) x (every sequence still alive when the
        block was scheduled) -- a superset of what the reference would align -- and the walk is
        replayed on the host row by row with numpy over the columns alive AT THAT MOMENT:

          a row visits its live columns in order (x.id != y.id, neither excluded); every similar
          y that is not longer than x is excluded; the first similar y that IS longer excludes x,
          after which the reference's pair filter drops the rest of the row.

        Python objects are only built for pairs that are written (aligned pairs / distance files
        enabled) and for the summary lines.  Distances are computed LAZILY along the columns: a
        block of rows is aligned against the columns up to a couple of thousand past its own end,
        and a row that survives all of those is extended on its own (eight times wider each time)
        -- most rows are excluded by the first longer similar sequence among the next few thousand,
        and the reference would not have aligned the rest of their rows either.  When 30 % of the
        loaded sequences have been excluded the survivors are re-loaded, so the device stops
        computing dead columns.  With several GPUs (task.devices) the rows of a chunk are split over
        them; the walk stays sequential.  Inputs with repeated ids keep the per-pair path (the exclusion set is keyed by
        id, and the reference's groupby merges neighbouring rows of equal id)."""
        ids = [s.id for s in self.input]
        if len(set(ids)) != len(ids):
            return self._start_per_pair()
        from ..engine import Engine, scores_vector

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
        raw_len = np.array([len(s.seq) for s in data], dtype=np.int64)
        multi = task_engine(self)
        scores = scores_vector(dict(p.pairs.scores)) if p.pairs.scores is not None else None
        if p.pairs.align:
            multi.set_scores(scores)

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
        writes_pairs = pairs_file is not None or linear_file is not None or matrix_file is not None
        side = None   # a context of its own for the consumer's thread (gapped strings of the visited pairs)
        text = lambda d: missing if d is None else fmt.format(d)  # noqa: E731

        alive = np.ones(n, dtype=bool)
        total = n * n
        done, last = 0, perf_counter()
        self.stats = dict(pairs_computed=0, pairs_visited=0, reloads=0)
        pointer = 0
        try:
            while pointer < n:
                loaded = np.nonzero(alive)[0]                      # original indices on the device this epoch
                first_pos = int(np.searchsorted(loaded, pointer))
                if first_pos >= len(loaded):
                    break
                multi.load([work[k].seq for k in loaded], 0)
                self.stats["reloads"] += 1
                if pairs_file is not None:
                    if side is None:
                        side = Engine(multi.devices[0], scores)
                    side.load([work[k].seq for k in loaded], 0)
                m = len(loaded)
                len_loaded = raw_len[loaded]
                first_chunk = min(m, self.first_columns or 2048)
                rows_per_block = self.rows_per_block or max(8, min(256, (1 << 20) // max(first_chunk, 1)))

                def compute_rect(r0, r1, c0, c1):
                    """Main-metric distances of rows [r0, r1) x columns [c0, c1) of the loaded set; rows are
                    split over the GPUs when the rectangle is large enough to keep several busy."""
                    ngpu = len(multi.engines)
                    parts = ngpu if (r1 - r0) >= ngpu and (r1 - r0) * (c1 - c0) >= ngpu * (1 << 17) else 1
                    cuts = [r0 + (r1 - r0) * k // parts for k in range(parts + 1)]
                    out = [None] * parts

                    def one(k):
                        eng = multi.engines[k]
                        call = eng.align_rect if p.pairs.align else eng.count_rect
                        out[k] = call(cuts[k], cuts[k + 1] - cuts[k], c0, c1 - c0, want=("metrics",))["metrics"][..., col]

                    if parts == 1:
                        one(0)
                    else:
                        import threading
                        threads = [threading.Thread(target=one, args=(k,), daemon=True) for k in range(parts)]
                        for t in threads:
                            t.start()
                        for t in threads:
                            t.join()
                        if any(o is None for o in out):
                            raise RuntimeError("a device failed while computing a block of distances")
                    self.stats["pairs_computed"] += (r1 - r0) * (c1 - c0)
                    return np.concatenate(out, axis=0) if parts > 1 else out[0]

                pos = first_pos
                reload = False
                while pos < m and not reload:
                    block_end = min(m, pos + rows_per_block)
                    values = np.empty((block_end - pos, m), dtype=np.float64)
                    # A row is excluded by the first LONGER similar sequence it meets.  The columns before
                    # the block are rows already walked (after a reload: the survivors so far, few and
                    # rarely similar to anything left); the likely partners are the next few thousand.
                    # So a block is aligned against the columns up to `first_chunk` past its own end, and
                    # a row that survives those is extended on its own, eight times wider each time.
                    first_end = min(m, block_end + first_chunk)
                    values[:, :first_end] = compute_rect(pos, block_end, 0, first_end)
                    for r in range(pos, block_end):
                        i = int(loaded[r])
                        pointer = i + 1
                        if not alive[i]:
                            continue
                        x = data[i]
                        first_d, have_first = None, False
                        c0 = 0
                        have = first_end      # columns [0, have) of this row are computed
                        row_over = False
                        while c0 < m and not row_over:
                            if c0 >= have:
                                nxt = min(m, have + 8 * (have - pos))
                                values[r - pos, have:nxt] = compute_rect(r, r + 1, have, nxt)[0]
                                have = nxt
                            c1 = have
                            live = alive[loaded[c0:c1]]
                            if c0 <= r < c1:
                                live[r - c0] = False                   # x.id != y.id
                            idx = np.nonzero(live)[0] + c0
                            c0 = c1
                            if idx.size == 0:
                                continue
                            d = values[r - pos, idx] * scale
                            with np.errstate(invalid="ignore"):
                                similar = np.isfinite(d) & (d <= similarity)
                            longer = similar & (len_loaded[idx] > raw_len[i])
                            if longer.any():
                                stop = int(np.argmax(longer))
                                row_over = True                        # x is excluded there: the rest of the row is never pulled
                            else:
                                stop = idx.size - 1
                            visited = idx[: stop + 1]
                            dv, sv = d[: stop + 1], similar[: stop + 1]
                            if not have_first:
                                first_d, have_first = (float(dv[0]) if np.isfinite(dv[0]) else None), True
                            if writes_pairs:
                                strings = None
                                if pairs_file is not None:   # one launch for the visited columns of this stretch of the row
                                    ax, ay, _ = side.align_strings(np.full(len(visited), r, dtype=np.int32), visited.astype(np.int32))
                                    strings = (ax, ay)
                                for q, vpos in enumerate(visited):
                                    j = int(loaded[vpos])
                                    if strings is not None:
                                        pair = SequencePair(Sequence(x.id, strings[0][q].decode("latin-1"), work[i].extras),
                                                            Sequence(data[j].id, strings[1][q].decode("latin-1"), work[j].extras))
                                        pairs_file.write(pair)
                                    else:
                                        pair = SequencePair(work[i], work[j])
                                    distance = Distance(metric, pair.x, pair.y, float(dv[q]) if np.isfinite(dv[q]) else None)
                                    if linear_file:
                                        linear_file.write(distance)
                                    if matrix_file:
                                        matrix_file.write(distance)
                            for q in np.nonzero(sv)[0]:
                                j = int(loaded[visited[q]])
                                y_info = (data[j].id, int(raw_len[j]), float(dv[q]))
                                x_info = (x.id, int(raw_len[i]), first_d)
                                inc, exc = (y_info, x_info) if longer[q] else (x_info, y_info)
                                gone = i if longer[q] else j
                                alive[gone] = False
                                self.excluded.add(data[gone].id)
                                summary.write((x.id, str(int(raw_len[i])), inc[0], str(inc[1]), text(inc[2]), exc[0], str(exc[1]), text(exc[2])))
                            done += len(visited)
                            self.stats["pairs_visited"] += len(visited)
                        now = perf_counter()
                        if now - last >= self.progress_interval:
                            self.progress_handler("distance.x.id", done, total - len(self.excluded) * n)
                            last = now
                    pos = block_end
                    # the device keeps computing the columns that died since the load: start over
                    # with the survivors once 30 % of the loaded set is gone
                    reload = int(alive[loaded].sum()) * 10 < 7 * m and pointer < n
            self.progress_handler("Finalizing...", total, total)
        finally:
            for w in writers:
                w.close()
            if side is not None:
                side.close()

        with output_handler(self.output_format, self.paths.dereplicated) as kept, \
                output_handler(self.output_format, self.paths.excluded) as dropped:
            for sequence in data:
                (dropped if sequence.id in self.excluded else kept).write(sequence)
        return Results(self.work_dir, perf_counter() - ts)

    def _start_per_pair(self) -> Results:
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
            per = self.rows_per_block or 64
            for x0 in range(0, n, per):
                rows = [i for i in range(x0, min(n, x0 + per)) if data[i].id not in self.excluded]
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
