"""VersusAll: every sequence against every sequence (full ordered N x N product, diagonal
included), alignment + the selected metrics, with the reference's output files.

Mirrors /root/reference/src/itaxotools/taxi2/tasks/versus_all.py (attribute surface :374-415,
pipeline :732-773, writers :98-350, statistics files :448-520).  Difference, out of the hot-path
scope: histogram plots are not produced.
"""
from __future__ import annotations

from itertools import chain, product
from math import inf
from pathlib import Path
from time import perf_counter
from typing import Callable, Iterator, NamedTuple

import numpy as np

from .. import fastwrite
from ..align import Scores
from ..distances import Distance, DistanceHandler, DistanceMetric
from ..handlers import FileHandler
from ..pairs import SequencePair, SequencePairHandler
from ..sequences import Sequence, Sequences
from ..statistics import StatisticsCalculator, StatisticsHandler
from ..types import AttrDict
from .common import ComparisonType, Results, console_report, create_parents, iter_pair_blocks, metric_columns, number_or_none, task_engine


class SimpleAggregator:
    """Running sum / min / max / n of the defined distances (versus_all.py:57-77);
    max starts at 0.0 and min at +inf exactly like the reference."""

    def __init__(self):
        self.sum, self.min, self.max, self.n = 0.0, inf, 0.0, 0

    def add(self, value: float | None) -> None:
        if value is None:
            return
        self.sum += value
        self.min = min(self.min, value)
        self.max = max(self.max, value)
        self.n += 1

    def calculate(self):
        if not self.n:
            return (None, None, None, 0)
        return (self.min, self.max, self.sum / self.n, self.n)


class DistanceStatistics(NamedTuple):
    metric: DistanceMetric
    idx: str
    idy: str
    min: float
    max: float
    mean: float
    count: int


class _FixedStatistics:
    """Aggregate that was computed elsewhere (native aggregator)."""

    def __init__(self, mn, mx, mean, n):
        self.stats = (mn, mx, mean, n)

    def calculate(self):
        return self.stats


class DistanceAggregator:
    def __init__(self, metric: DistanceMetric):
        self.metric = metric
        self.aggs: dict = {}

    def set(self, idx, idy, mn, mx, mean, n) -> None:
        self.aggs[(idx, idy)] = _FixedStatistics(mn, mx, mean, n)

    def add(self, idx, idy, d) -> None:
        agg = self.aggs.get((idx, idy))
        if agg is None:
            agg = self.aggs[(idx, idy)] = SimpleAggregator()
        agg.add(d)

    def __iter__(self) -> Iterator[DistanceStatistics]:
        for (idx, idy), agg in self.aggs.items():
            mn, mx, mean, n = agg.calculate()
            yield DistanceStatistics(self.metric, idx, idy, mn, mx, mean, n)


COMPARISON = {
    (None, None): ComparisonType.Unknown,
    (None, True): ComparisonType.IntraSpecies,
    (None, False): ComparisonType.InterSpecies,
    (False, None): ComparisonType.InterGenus,
    (False, True): ComparisonType.InterGenus,
    (False, False): ComparisonType.InterGenus,
    (True, None): ComparisonType.IntraGenus,
    (True, True): ComparisonType.IntraSpecies,
    (True, False): ComparisonType.InterSpecies,
}


class VersusAll:
    def __init__(self):
        self.work_dir: Path = None
        self.paths = AttrDict()
        self.progress_handler: Callable = console_report
        self.progress_interval: float = 0.015
        self.device: int = 0
        self.devices = None   # list of CUDA device indices or "all": shard the pair product over several GPUs
        self.native_writers: bool = True   # batch formatter for plain float formats (same bytes as the handlers)

        self.input = AttrDict()
        self.input.sequences: Sequences = None
        self.input.species = None
        self.input.genera = None

        self.params = AttrDict()
        self.params.pairs = AttrDict(align=True, write=True, scores=None)
        self.params.distances = AttrDict(metrics=None, write_linear=True, write_matricial=True)
        self.params.plot = AttrDict(histograms=True, binwidth=0.05, formats=None, palette=None)
        self.params.format = AttrDict(float="{:.4f}", percentage="{:.2f}", missing="NA",
                                      stats_template="{mean} ({min}-{max})", percentage_multiply=False)
        self.params.stats = AttrDict(all=True, species=True, genera=True)

    # -- paths / parameters (versus_all.py:417-446) -----------------------------------------------
    def generate_paths(self):
        assert self.work_dir
        w = Path(self.work_dir)
        self.paths.summary = w / "summary.tsv"
        self.paths.stats_all = w / "stats" / "all.tsv"
        self.paths.stats_species = w / "stats" / "species.tsv"
        self.paths.stats_genera = w / "stats" / "genera.tsv"
        self.paths.aligned_pairs = w / "align" / "aligned_pairs.txt"
        self.paths.distances_linear = w / "distances" / "linear.tsv"
        self.paths.distances_matricial = w / "distances" / "matricial"
        self.paths.subsets = w / "subsets"
        create_parents(self.paths.summary)

    def check_metrics(self):
        self.params.distances.metrics = self.params.distances.metrics or [
            DistanceMetric.Uncorrected(), DistanceMetric.UncorrectedWithGaps(),
            DistanceMetric.JukesCantor(), DistanceMetric.Kimura2P()]

    # -- stats/*.tsv (versus_all.py:448-520): one pass over the (normalized) sequences ---------------
    def write_statistics(self, sequences: list[Sequence]) -> None:
        p = self.params
        options = dict(float_formatter=p.format.float, percentage_formatter=p.format.percentage,
                       percentage_multiply=p.format.percentage_multiply)
        if p.stats.all:
            create_parents(self.paths.stats_all)
            with StatisticsHandler.Single(self.paths.stats_all, "w", **options) as file:
                file.write(StatisticsCalculator(s.seq.upper() for s in sequences).calculate())
        for partition, enabled, name, path in ((self.input.species, p.stats.species, "species", self.paths.stats_species),
                                               (self.input.genera, p.stats.genera, "genera", self.paths.stats_genera)):
            if not partition or not enabled:
                continue
            calculators = {}
            for subset in partition.values():          # groups in the partition's own order
                if subset not in calculators:
                    calculators[subset] = StatisticsCalculator(group=subset)
            for s in sequences:
                subset = partition.get(s.id, None)
                if subset is not None:
                    calculators[subset].add(s.seq.upper())
            create_parents(path)
            with StatisticsHandler.Groups(path, "w", group_name=name, **options) as file:
                for calc in calculators.values():
                    file.write(calc.calculate())

    # -- the run -----------------------------------------------------------------------------------
    def start(self) -> Results:
        ts = perf_counter()
        self.generate_paths()
        self.check_metrics()
        p = self.params
        metrics = p.distances.metrics
        columns = metric_columns(metrics)
        fmt, missing = p.format.float, p.format.missing
        scale = 100.0 if p.format.percentage_multiply else 1.0

        sequences = list(self.input.sequences.normalize() if p.pairs.align else self.input.sequences)
        n = len(sequences)
        engine = task_engine(self)
        self.write_statistics(sequences)

        writers = []
        pairs_file = linear_file = None
        matrix_files = []
        native_pairs = None   # [ids of all sequences, nothing written yet]: the native pair writer appends block by block
        if p.pairs.align and p.pairs.write:
            create_parents(self.paths.aligned_pairs)
            if self.native_writers:
                Path(self.paths.aligned_pairs).write_bytes(b"")
                native_pairs = [fastwrite.StringTable([s.id for s in sequences]), True]
            else:
                pairs_file = SequencePairHandler.Formatted(self.paths.aligned_pairs, "w")
                writers.append(pairs_file)
        if p.distances.write_linear:
            create_parents(self.paths.distances_linear)
            linear_file = DistanceHandler.Linear.WithExtras(self.paths.distances_linear, "w", missing=missing, formatter=fmt)
            writers.append(linear_file)
        if p.distances.write_matricial:
            create_parents(self.paths.distances_matricial / "x.tsv")
            for metric in metrics:
                f = DistanceHandler.Matrix(self.paths.distances_matricial / f"{metric}.tsv", "w", missing=missing, formatter=fmt)
                matrix_files.append(f)
                writers.append(f)
        summary = SummaryWriter(self.paths.summary, missing, fmt)
        agg_species = SubsetAggregation(self.input.species, metrics) if self.input.species else None
        agg_genera = SubsetAggregation(self.input.genera, metrics) if self.input.genera else None

        total = len(metrics) * n * n
        done, last_time = 0, perf_counter()
        fmtc = fastwrite.printf_format(fmt) if self.native_writers else None
        native = NativeBlockWriter(self, sequences, metrics, columns, scale, fmtc, missing, linear_file, matrix_files, summary,
                                   agg_genera, agg_species) if fmtc else None
        same_key = duplicate_groups(sequences)
        try:
            for block in iter_pair_blocks(engine, sequences, None, p.pairs.align, pairs_file is not None or native_pairs is not None,
                                          p.pairs.scores, raw_strings=native_pairs is not None,
                                          extra=lambda eng, blk: self._undefined_mask(eng, blk, sequences, same_key, p.pairs.align)):
                undefined = block.extra
                if native_pairs is not None:
                    ids, first = native_pairs
                    fastwrite.format_aligned_pairs(self.paths.aligned_pairs, first, ids, ids, block.x0, block.nx, n, *block.aligned_raw)
                    native_pairs[1] = False
                if pairs_file is not None:
                    for bx in range(block.nx):
                        x = sequences[block.x0 + bx]
                        for j, y in enumerate(sequences):
                            ax, ay = block.aligned[bx * n + j]
                            pairs_file.write(SequencePair(Sequence(x.id, ax, x.extras), Sequence(y.id, ay, y.extras)))
                if native is not None:
                    native.write_block(block, undefined)
                    done += len(metrics) * block.nx * n
                    now = perf_counter()
                    if now - last_time >= self.progress_interval:
                        self.progress_handler("distance.x.id", done, total)
                        last_time = now
                    continue
                for bx in range(block.nx):
                    x = sequences[block.x0 + bx]
                    for j in range(n):
                        y = sequences[j]
                        row = []
                        for metric, col in zip(metrics, columns):
                            d = None if undefined[bx, j] else number_or_none(block.metrics[bx, j, col])
                            if d is not None:
                                d *= scale
                            row.append(Distance(metric, x, y, d))
                        for k, distance in enumerate(row):
                            if linear_file:
                                linear_file.write(distance)
                            if matrix_files:
                                matrix_files[k].write(distance)
                        gen = agg_genera.add(row) if agg_genera else None
                        spe = agg_species.add(row) if agg_species else None
                        summary.write(row, gen, spe)
                        done += len(row)
                        now = perf_counter()
                        if now - last_time >= self.progress_interval:
                            self.progress_handler("distance.x.id", done, total)
                            last_time = now
            self.progress_handler("Finalizing...", total, total)
        finally:
            for w in writers:
                w.close()
            summary.close()
        if native is not None:
            native.finish()
        if agg_genera:
            agg_genera.write(self.paths.subsets / "genera", p.format)
        if agg_species:
            agg_species.write(self.paths.subsets / "species", p.format)
        return Results(self.work_dir, perf_counter() - ts)

    @staticmethod
    def _undefined_mask(engine, block, sequences, same_key, aligned: bool) -> np.ndarray:
        """versus_all.py:549-552: distances are None when the two (aligned) records compare equal as
        tuples.  Equal aligned strings imply equal raw strings, so only pairs of records with
        identical id / sequence / extras are candidates; when aligned they are settled by the
        alignment itself (strings of this block if they were produced, else one small extra launch)."""
        n = len(sequences)
        mask = np.zeros((block.nx, n), dtype=np.uint8)
        cand = [(bx, j) for bx in range(block.nx) for j in same_key[block.x0 + bx]]
        if not cand:
            return mask
        if not aligned:
            for bx, j in cand:
                mask[bx, j] = 1
            return mask
        if block.aligned is not None:
            for bx, j in cand:
                ax, ay = block.aligned[bx * n + j]
                mask[bx, j] = ax == ay
            return mask
        if block.aligned_raw is not None:
            ox, oy, start, off = block.aligned_raw
            for bx, j in cand:
                k = bx * n + j
                mask[bx, j] = np.array_equal(ox[start[k]:off[k + 1]], oy[start[k]:off[k + 1]])
            return mask
        px = np.array([block.x0 + bx for bx, _ in cand], dtype=np.int32)
        py = np.array([j for _, j in cand], dtype=np.int32)
        ax, ay, _ = engine.align_strings(px, py)   # indices into the loaded set (it serves both sides)
        for k, (bx, j) in enumerate(cand):
            mask[bx, j] = ax[k] == ay[k]
        return mask


def duplicate_groups(sequences) -> list[list[int]]:
    """For every record, the indices of all records that equal it as a tuple (itself included)."""
    groups: dict = {}
    for k, s in enumerate(sequences):
        groups.setdefault((s.id, s.seq, tuple(s.extras.items())), []).append(k)
    return [groups[(s.id, s.seq, tuple(s.extras.items()))] for s in sequences]


class NativeBlockWriter:
    """Feeds whole result blocks to the native formatter / aggregator (fastwrite) instead of
    sending every value through the Python handlers.  Same files, same bytes."""

    def __init__(self, task, sequences, metrics, columns, scale, fmtc, missing, linear_file, matrix_files, summary, agg_genera, agg_species):
        self.task, self.sequences, self.metrics, self.columns = task, sequences, metrics, columns
        self.scale, self.fmtc, self.missing = scale, fmtc, missing
        self.n = len(sequences)
        self.linear_path = task.paths.distances_linear if linear_file is not None else None
        self.matrix_paths = [task.paths.distances_matricial / f"{m}.tsv" for m in metrics] if matrix_files else []
        self.summary_path = task.paths.summary
        self.agg = {"genera": agg_genera, "species": agg_species}
        self.header_done = False
        self.pool = None
        # the Python handlers stay open (they create / truncate the files) but never receive a row;
        # they are closed before the first native append so the two never interleave
        self.python_handles = [h for h in [linear_file, *matrix_files] if h is not None]
        self.summary = summary
        fill = lambda values: [missing if v is None else v for v in values]  # noqa: E731
        ids = [s.id for s in sequences]
        self.t_id = fastwrite.StringTable(ids)
        self.t_rec = fastwrite.StringTable(["\t".join([s.id, *fill(s.extras.values())]) for s in sequences])
        self.has_extras = bool(sequences) and bool(sequences[0].extras)
        self.t_extras = fastwrite.StringTable(["\t".join(fill(s.extras.values())) for s in sequences])
        genera, species = task.input.genera, task.input.species
        taxon = lambda part, s: ((part.get(s.id, None) if part else None) or "-")  # noqa: E731
        self.t_taxon = fastwrite.StringTable([taxon(genera, s) + "\t" + taxon(species, s) for s in sequences])
        self.ids = {}
        self.states = {}
        for name, part in (("genera", genera), ("species", species)):
            if not part:
                self.ids[name] = None
                continue
            names: dict = {}
            raw = [part.get(s.id, None) for s in sequences]
            for v in raw:
                if v is not None:
                    names.setdefault(v, len(names))
            comp = np.array([names[v] if v is not None else -1 for v in raw], dtype=np.int32)
            agg = np.array([names[v] if v is not None else len(names) for v in raw], dtype=np.int32)
            self.ids[name] = (comp, agg, [*names, None])
            self.states[name] = [fastwrite.NativeSubsetState(len(names) + 1) for _ in metrics]

    def _headers(self):
        for h in self.python_handles:
            h.close()
        self.summary.close()
        if not self.n:
            return
        first = self.sequences[0]
        labels = [str(m) for m in self.metrics]
        if self.linear_path is not None:
            row = ("seqid (query)", *(k + " (query)" for k in first.extras), "seqid (reference)",
                   *(k + " (reference)" for k in first.extras), *labels)
            self.linear_path.write_text("\t".join(row) + "\n")
        for path in self.matrix_paths:
            path.write_text("\t".join(("", *(s.id for s in self.sequences))) + "\n")
        tx, ty = " (query 1)", " (query 2)"
        row = ("seqid" + tx, "seqid" + ty, *labels, *(k + tx for k in first.extras), *(k + ty for k in first.extras),
               "genus" + tx, "species" + tx, "genus" + ty, "species" + ty, "comparison_type")
        self.summary_path.write_text("\t".join(row) + "\n")

    def write_block(self, block, undefined) -> None:
        if not self.header_done:
            self._headers()
            self.header_done = True
        fw = fastwrite
        n, m = self.n, block.metrics
        # every file (and every subset aggregate) is independent of the others: the native calls run
        # side by side on a few host threads (ctypes drops the GIL), each appending to its own file
        jobs = []
        if self.linear_path is not None:
            jobs.append(lambda: fw.format_pairs(self.linear_path, [fw.SEG_X[0], fw.SEG_Y[0], fw.SEG_SCORES], [self.t_rec], [self.t_rec],
                                                block.x0, block.nx, n, m, undefined, self.columns, self.scale, self.fmtc, self.missing))
        for path, col in zip(self.matrix_paths, self.columns):
            jobs.append(lambda path=path, col=col: fw.format_matrix(path, self.t_id, block.x0, block.nx, n, m, undefined, col, self.scale,
                                                                    self.fmtc, self.missing))
        seg = [fw.SEG_X[0], fw.SEG_Y[0], fw.SEG_SCORES]
        if self.has_extras:
            seg += [fw.SEG_X[1], fw.SEG_Y[1]]
        seg += [fw.SEG_X[2], fw.SEG_Y[2], fw.SEG_COMPARISON]
        g, s = self.ids["genera"], self.ids["species"]
        jobs.append(lambda: fw.format_pairs(self.summary_path, seg, [self.t_id, self.t_extras, self.t_taxon],
                                            [self.t_id, self.t_extras, self.t_taxon], block.x0, block.nx, n, m, undefined, self.columns,
                                            self.scale, self.fmtc, self.missing, xgenus=g[0] if g else None, xspecies=s[0] if s else None,
                                            ygenus=g[0] if g else None, yspecies=s[0] if s else None))
        for name in ("genera", "species"):
            if self.ids[name] is None:
                continue
            agg_ids = self.ids[name][1]
            for state, col in zip(self.states[name], self.columns):
                jobs.append(lambda state=state, col=col, agg_ids=agg_ids: state.add_block(m, undefined, block.x0, block.nx, n, col,
                                                                                         self.scale, agg_ids, agg_ids))
        if self.pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self.pool = ThreadPoolExecutor(max_workers=6, thread_name_prefix="taxi-writer")
        for future in [self.pool.submit(job) for job in jobs]:
            future.result()

    def finish(self) -> None:
        if self.pool is not None:
            self.pool.shutdown()
            self.pool = None
        if not self.header_done:   # no sequences at all: the handlers produced the (empty) files
            for h in self.python_handles:
                h.close()
            self.summary.close()
        for name in ("genera", "species"):
            agg = self.agg[name]
            if agg is None:
                continue
            # the subset files are written from the native states' arrays in one vectorised pass
            # (SubsetAggregation.write_arrays); with S subsets they hold S^2 rows / cells, which the
            # per-key Python objects of the handler path made the slowest part of a large task
            agg.native = (self.ids[name][2], [str(m) for m in self.metrics], self.states[name])


class SubsetAggregation:
    """Per (subset_x, subset_y) statistics of every metric (versus_all.py:623-684)."""

    def __init__(self, partition, metrics):
        self.partition = partition
        self.aggregators = {str(m): DistanceAggregator(m) for m in metrics}

    def add(self, row: list[Distance]):
        sx = self.partition.get(row[0].x.id, None)
        sy = self.partition.get(row[0].y.id, None)
        for d in row:
            self.aggregators[str(d.metric)].add(sx, sy, d.d)
        return (sx, sy)

    native = None   # (subset names, metric labels, NativeSubsetState per metric) when the native aggregator ran

    def write_arrays(self, path: Path, fmt) -> None:
        """The same files as write(), from the arrays of the native aggregator (versus_all.py:143-249,
        647-684): keys in first-seen order, mean / min / max per metric, "NA" where nothing was
        defined; matrix rows are the runs of equal first subset in that order."""
        names, labels, states = self.native
        nsub = states[0].nsub
        keys = np.nonzero(states[0].first_seen >= 0)[0]
        keys = keys[np.argsort(states[0].first_seen[keys], kind="stable")]
        kx, ky = keys // nsub, keys % nsub
        label_of = ["?" if v is None else v for v in names]
        tx = [label_of[k] for k in kx.tolist()]
        ty = [label_of[k] for k in ky.tolist()]
        form = fmt.float.format
        fmtc = fastwrite.printf_format(fmt.float)
        import re
        plain = re.match(r"^([^{}]*)\{mean\}([^{}]*)\{min\}([^{}]*)\{max\}([^{}]*)$", fmt.stats_template)
        heads = [f"{label} {stat}" for label in labels for stat in ("mean", "min", "max")]
        if fmtc and plain:
            # plain float spec and a plain statistics template: the library writes every row (a partition of a
            # thousand species has a million keys; a dozen Python strings per key took most of a large run's tail)
            table = fastwrite.StringTable(label_of)
            kx32, ky32 = kx.astype(np.int32), ky.astype(np.int32)
            cnts = [st.count[keys] for st in states]
            with np.errstate(invalid="ignore", divide="ignore"):
                means = [st.sum[keys] / c for st, c in zip(states, cnts)]
            mins, maxs = [st.min[keys] for st in states], [st.max[keys] for st in states]
            linear = path / "linear"
            linear.mkdir(parents=True, exist_ok=True)
            for name, select, lead in (("pairs.tsv", kx != ky, ("target", "query")), ("identity.tsv", kx == ky, ("target",))):
                with open(linear / name, "w", newline="") as f:
                    if select.any():
                        f.write("\t".join((*lead, *heads)) + "\n")
                if select.any():
                    fastwrite.format_subset_rows(linear / name, table, kx32, ky32, select, len(lead) == 2, means, mins, maxs, cnts, fmtc)
            matricial = path / "matricial"
            matricial.mkdir(parents=True, exist_ok=True)
            for label, mean, mn, mx, cnt in zip(labels, means, mins, maxs, cnts):
                fastwrite.format_subset_matrix(matricial / f"{label}.tsv", table, kx32, ky32, mean, mn, mx, cnt, fmtc, plain.groups())
            return

        def column(values, counts):
            if fmtc:   # plain float spec: the library formats the whole column in one call
                return fastwrite.format_values(np.where(counts > 0, values, 0.0), counts == 0, fmtc, "NA")
            return [form(v) if c else "NA" for v, c in zip(values.tolist(), counts.tolist())]

        stats = []          # per metric: (mean, min, max) text columns over the keys
        for st in states:
            cnt = st.count[keys]
            with np.errstate(invalid="ignore", divide="ignore"):
                mean = st.sum[keys] / cnt
            stats.append((column(mean, cnt), column(st.min[keys], cnt), column(st.max[keys], cnt), cnt))
        linear = path / "linear"
        linear.mkdir(parents=True, exist_ok=True)
        same = (kx == ky).tolist()
        rows_pairs, rows_identity = [], []
        columns = [col for st in stats for col in st[:3]]
        for k, values in enumerate(zip(*columns)):
            if same[k]:
                rows_identity.append("\t".join((tx[k], *values)))
            else:
                rows_pairs.append("\t".join((tx[k], ty[k], *values)))
        with open(linear / "pairs.tsv", "w", newline="") as f:
            if rows_pairs:
                f.write("\t".join(("target", "query", *heads)) + "\n" + "\n".join(rows_pairs) + "\n")
        with open(linear / "identity.tsv", "w", newline="") as f:
            if rows_identity:
                f.write("\t".join(("target", *heads)) + "\n" + "\n".join(rows_identity) + "\n")
        matricial = path / "matricial"
        matricial.mkdir(parents=True, exist_ok=True)
        # runs of equal first subset (the reference flushes a matrix row whenever idx changes)
        starts = [0] + [k for k in range(1, len(keys)) if kx[k] != kx[k - 1]] + [len(keys)] if len(keys) else [0]
        template = fmt.stats_template
        for label, (mean, mn, mx, cnt) in zip(labels, stats):
            has = cnt.tolist()
            if plain:   # "{mean} ({min}-{max})" and the like: concatenation instead of 1e6 str.format calls
                p0, p1, p2, p3 = plain.groups()
                cells = [p0 + a + p1 + b + p2 + c + p3 if h else "NA" for a, b, c, h in zip(mean, mn, mx, has)]
            else:
                cells = [template.format(mean=a, min=b, max=c) if h else "NA" for a, b, c, h in zip(mean, mn, mx, has)]
            with open(matricial / f"{label}.tsv", "w", newline="") as f:
                for r in range(len(starts) - 1):
                    lo, hi = starts[r], starts[r + 1]
                    if r == 0:
                        f.write("\t".join(("", *ty[lo:hi])) + "\n")
                    f.write("\t".join((tx[lo], *cells[lo:hi])) + "\n")

    def write(self, path: Path, fmt) -> None:
        if self.native is not None:
            return self.write_arrays(path, fmt)
        # the subset writers of the reference keep their own default missing marker "NA"
        to_text = lambda v: "NA" if v is None else fmt.float.format(v)  # noqa: E731
        linear = path / "linear"
        linear.mkdir(parents=True, exist_ok=True)
        with FileHandler.Tabfile(linear / "pairs.tsv", "w") as pairs_file, \
                FileHandler.Tabfile(linear / "identity.tsv", "w") as identity_file:
            wrote = {"pairs": False, "identity": False}
            for bunch in zip(*(iter(a) for a in self.aggregators.values())):
                names = [f"{s.metric} {stat}" for s, stat in product(bunch, ["mean", "min", "max"])]
                values = [to_text(v) for v in chain(*((s.mean, s.min, s.max) for s in bunch))]
                idx = "?" if bunch[0].idx is None else bunch[0].idx
                idy = "?" if bunch[0].idy is None else bunch[0].idy
                if bunch[0].idx == bunch[0].idy:
                    if not wrote["identity"]:
                        identity_file.write(("target", *names))
                        wrote["identity"] = True
                    identity_file.write((idx, *values))
                else:
                    if not wrote["pairs"]:
                        pairs_file.write(("target", "query", *names))
                        wrote["pairs"] = True
                    pairs_file.write((idx, idy, *values))
        matricial = path / "matricial"
        matricial.mkdir(parents=True, exist_ok=True)
        for label, aggregator in self.aggregators.items():
            with FileHandler.Tabfile(matricial / f"{label}.tsv", "w") as file:
                line: list[DistanceStatistics] = []
                wrote_header = False

                def flush():
                    nonlocal wrote_header
                    if not wrote_header:
                        file.write(("", *("?" if s.idy is None else s.idy for s in line)))
                        wrote_header = True
                    cells = []
                    for s in line:
                        if not s.count:
                            cells.append("NA")
                        else:
                            cells.append(fmt.stats_template.format(mean=to_text(s.mean), min=to_text(s.min), max=to_text(s.max)))
                    file.write(("?" if line[0].idx is None else line[0].idx, *cells))

                for stats in aggregator:
                    if line and line[0].idx != stats.idx:
                        flush()
                        line = []
                    line.append(stats)
                if line:
                    flush()


class SummaryWriter:
    """summary.tsv (versus_all.py:278-350): ids, metrics, extras of both records, genus / species
    of both, and the comparison type."""

    def __init__(self, path: Path, missing: str, formatter: str):
        self.file = FileHandler.Tabfile(path, "w")
        self.missing, self.formatter = missing, formatter
        self.wrote_headers = False
        self.tagX, self.tagY = " (query 1)", " (query 2)"

    def write(self, row: list[Distance], genera, species) -> None:
        x, y = row[0].x, row[0].y
        if not self.wrote_headers:
            self.file.write(("seqid" + self.tagX, "seqid" + self.tagY, *(str(d.metric) for d in row),
                             *(k + self.tagX for k in x.extras), *(k + self.tagY for k in y.extras),
                             "genus" + self.tagX, "species" + self.tagX, "genus" + self.tagY, "species" + self.tagY,
                             "comparison_type"))
            self.wrote_headers = True
        same_genera = bool(genera[0] == genera[1]) if genera else None
        same_species = bool(species[0] == species[1]) if species else None
        fill = lambda values: [self.missing if v is None else v for v in values]  # noqa: E731
        self.file.write((x.id, y.id,
                         *(self.missing if d.d is None else self.formatter.format(d.d) for d in row),
                         *fill(x.extras.values()), *fill(y.extras.values()),
                         (genera[0] if genera else "-") or "-", (species[0] if species else "-") or "-",
                         (genera[1] if genera else "-") or "-", (species[1] if species else "-") or "-",
                         COMPARISON[(same_genera, same_species)].label))

    def close(self) -> None:
        self.file.close()
