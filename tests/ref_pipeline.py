"""Per-pair restatement of the reference's task pipelines for the parity tests (test
infrastructure): the same generator order as /root/reference/src/itaxotools/taxi2/tasks/
versus_all.py:732-773 and versus_reference.py:213-247, with the two native calls answered by the
CPU oracle.  Deliberately written pair-at-a-time, independent of taxi2_b200.tasks AND of the
package's file handlers: every output file is written by the small plain-text writers below,
restated from the reference's writer code, so a "want" tree never passes
through the code under test.  (The package's handlers are pinned separately, against the
reference's own fixture files, in tests/test_host_api.py.)"""
from __future__ import annotations

from itertools import groupby
from math import inf, isnan
from pathlib import Path

import oracle
from taxi2_b200.distances import Distance, DistanceMetric
from taxi2_b200.pairs import SequencePair
from taxi2_b200.sequences import Sequence


class PlainTab:
    """handlers.py:219-227: "\\t".join(row) + "\\n", LF endings, no quoting; optional header row."""

    def __init__(self, path, mode="w", columns=None):
        self.file = open(path, "w", newline="")
        if columns:
            self.write(columns)

    def write(self, row):
        self.file.write("\t".join(str(v) for v in row) + "\n")

    def close(self):
        self.file.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class PlainLinear(PlainTab):
    """distances.py:75-130, 244-279: one row per run of distances with the same (x.id, y.id);
    header from the first row: seqid (query), <extras> (query), seqid (reference), <extras>
    (reference), metric labels; None -> missing."""

    def __init__(self, path, mode="w", missing="NA", formatter="{:.4f}", tag_x=" (query)", tag_y=" (reference)"):
        super().__init__(path)
        self.missing, self.formatter, self.tags = missing, formatter, (tag_x, tag_y)
        self.run, self.header_done = [], False

    def write(self, distance):
        if self.run and (self.run[0].x.id, self.run[0].y.id) != (distance.x.id, distance.y.id):
            self.flush()
        self.run.append(distance)

    def flush(self):
        if not self.run:
            return
        x, y = self.run[0].x, self.run[0].y
        if not self.header_done:
            PlainTab.write(self, ["seqid" + self.tags[0], *(k + self.tags[0] for k in x.extras), "seqid" + self.tags[1],
                                  *(k + self.tags[1] for k in y.extras), *(str(d.metric) for d in self.run)])
            self.header_done = True
        text = lambda v: self.missing if v is None else self.formatter.format(v)  # noqa: E731
        fill = lambda v: self.missing if v is None else v  # noqa: E731
        PlainTab.write(self, [x.id, *map(fill, x.extras.values()), y.id, *map(fill, y.extras.values()), *(text(d.d) for d in self.run)])
        self.run = []

    def close(self):
        self.flush()
        super().close()


class PlainMatrix(PlainTab):
    """distances.py:143-186: one row per run of distances with the same x.id: x.id, scores; the
    header (an empty cell, then the y ids of the first row) precedes the first row."""

    def __init__(self, path, mode="w", missing="NA", formatter="{:.4f}"):
        super().__init__(path)
        self.missing, self.formatter = missing, formatter
        self.run, self.header_done = [], False

    def write(self, distance):
        if self.run and self.run[0].x.id != distance.x.id:
            self.flush()
        self.run.append(distance)

    def flush(self):
        if not self.run:
            return
        if not self.header_done:
            PlainTab.write(self, ["", *(d.y.id for d in self.run)])
            self.header_done = True
        PlainTab.write(self, [self.run[0].x.id, *(self.missing if d.d is None else self.formatter.format(d.d) for d in self.run)])
        self.run = []

    def close(self):
        self.flush()
        super().close()


class PlainPairs:
    """pairs.py:51-97: 'idx / idy', aligned x, pattern ('-' if either is a gap, '|' if equal, '.'
    otherwise), aligned y; one empty line between records."""

    def __init__(self, path, mode="w"):
        self.file = open(path, "w", newline="")
        self.first = True

    def write(self, pair):
        pattern = "".join("-" if "-" in (a, b) else ("|" if a == b else ".") for a, b in zip(pair.x.seq, pair.y.seq))
        self.file.write(("" if self.first else "\n") + f"{pair.x.id} / {pair.y.id}\n{pair.x.seq}\n{pattern}\n{pair.y.seq}\n")
        self.first = False

    def close(self):
        self.file.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class _Plain:
    """Stand-ins with the handler names the pipelines below were written against."""

    class FileHandler:
        Tabfile = PlainTab

    class DistanceHandler:
        Matrix = PlainMatrix

        class Linear:
            WithExtras = PlainLinear

    class SequencePairHandler:
        Formatted = PlainPairs


FileHandler, DistanceHandler, SequencePairHandler = _Plain.FileHandler, _Plain.DistanceHandler, _Plain.SequencePairHandler

LABELS = ["p", "p-gaps", "jc", "k2p"]


def oracle_metric(metric, x: Sequence, y: Sequence):
    c = oracle.count(x.seq, y.seq)
    if c is None:
        return None
    v = oracle.metrics(c)[LABELS.index(str(metric))]
    return None if isnan(v) else v


def oracle_align(pair: SequencePair, scores=None) -> SequencePair:
    ax, ay, _ = oracle.align(pair.x.seq, pair.y.seq, scores)
    return SequencePair(Sequence(pair.x.id, ax, pair.x.extras), Sequence(pair.y.id, ay, pair.y.extras))


STAT_LABELS = ["Total number of sequences", "Total length of all sequences ", "Number of sequences with 0 bp",
               "Number of sequences with less than 100 bp", "Number of sequences between 101-300 bp",
               "Number of sequences between 301-1000 bp", "Number of sequences with more than 1000 bp", "Minimum sequence length",
               "Maximum sequence length ", "Mean sequence length  ", "Median sequence length  ", "Standard deviation of sequence length",
               "Percentage of base A", "Percentage of base C", "Percentage of base G", "Percentage of base T", "GC content",
               "Percentage of ambiguity codes", "Percentage of missing data ", "Percentage of missing data including gaps",
               "Percentage of gaps", "N50 statistic", "L50 statistic", "N90 statistic", "L90 statistic"]


def stat_values(seqs: list[str], ffmt: str, pfmt: str, multiply: bool) -> list[str]:
    """The 25 statistics of a list of upper-case strings as text, computed with numpy straight
    from the definitions in /root/reference/src/itaxotools/taxi2/statistics.py:45-224."""
    import numpy as np
    joined = "".join(seqs)
    cnt = {ch: joined.count(ch) for ch in "-NACGT"}
    lens = np.array([len(s) - s.count("-") for s in seqs], dtype=np.int64)
    n, nuc, tot = len(seqs), int(lens.sum()), len(joined)

    def nl(arg):
        if not lens.any():
            return 0, 0
        d = np.sort(lens)[::-1]
        pos = int(np.argmax(np.cumsum(d) >= d.sum() * arg / 100))
        return int(d[pos]), pos + 1
    ints = [n, nuc, int((lens == 0).sum()), int(((lens > 0) & (lens <= 100)).sum()), int(((lens > 100) & (lens <= 300)).sum()),
            int(((lens > 300) & (lens <= 1000)).sum()), int((lens > 1000).sum()), int(lens.min()) if n else 0, int(lens.max()) if n else 0]
    from statistics import median, pstdev
    floats = [nuc / n if n else 0.0, float(median(lens.tolist())) if n else 0.0, float(pstdev(lens.tolist())) if n > 1 else 0.0]
    frac = lambda v, d: (v / d if d else 0.0) * (100 if multiply else 1)  # noqa: E731
    amb = nuc - cnt["N"] - cnt["A"] - cnt["C"] - cnt["G"] - cnt["T"]
    pct = [frac(cnt["A"], nuc), frac(cnt["C"], nuc), frac(cnt["G"], nuc), frac(cnt["T"], nuc), frac(cnt["C"] + cnt["G"], nuc),
           frac(amb, nuc), frac(cnt["N"], nuc), frac(cnt["N"] + cnt["-"], tot), frac(cnt["-"], tot)]
    return [*map(str, ints), *(ffmt.format(v) for v in floats), *(pfmt.format(v) for v in pct), *map(str, nl(50) + nl(90))]


def write_stats(seqs, work: Path, species, genera, ffmt: str, pfmt: str = "{:.2f}", multiply: bool = False) -> None:
    (work / "stats").mkdir(parents=True, exist_ok=True)
    with FileHandler.Tabfile(work / "stats" / "all.tsv", "w") as f:
        for row in zip(STAT_LABELS, stat_values([s.seq.upper() for s in seqs], ffmt, pfmt, multiply)):
            f.write(row)
    for name, part in (("species", species), ("genera", genera)):
        if not part:
            continue
        with FileHandler.Tabfile(work / "stats" / f"{name}.tsv", "w") as f:
            f.write((name, *STAT_LABELS))
            for group in dict.fromkeys(part.values()):
                members = [s.seq.upper() for s in seqs if part.get(s.id, None) == group]
                f.write((group, *stat_values(members, ffmt, pfmt, multiply)))


COMPARISON = {(None, None): "no info", (None, True): "intra-species", (None, False): "inter-species",
              (False, None): "inter-genus", (False, True): "inter-genus", (False, False): "inter-genus",
              (True, None): "intra-genus", (True, True): "intra-species", (True, False): "inter-species"}


def versus_all(sequences, work: Path, species=None, genera=None, align=True, metrics=None, fmt="{:.4f}", missing="NA",
               multiply=False):
    metrics = metrics or [DistanceMetric.fromLabel(label) for label in LABELS]
    work = Path(work)
    (work / "align").mkdir(parents=True, exist_ok=True)
    (work / "distances" / "matricial").mkdir(parents=True, exist_ok=True)
    seqs = [s.normalize() for s in sequences] if align else list(sequences)
    write_stats(seqs, work, species, genera, fmt, multiply=multiply)
    text = lambda v: missing if v is None else fmt.format(v)  # noqa: E731
    aggs = {name: {str(m): {} for m in metrics} for name in ("genera", "species")}
    with SequencePairHandler.Formatted(work / "align" / "aligned_pairs.txt", "w") as pairs_file, \
            DistanceHandler.Linear.WithExtras(work / "distances" / "linear.tsv", "w", missing=missing, formatter=fmt) as linear, \
            FileHandler.Tabfile(work / "summary.tsv", "w") as summary:
        matrices = [DistanceHandler.Matrix(work / "distances" / "matricial" / f"{m}.tsv", "w", missing=missing, formatter=fmt) for m in metrics]
        wrote_summary_header = False
        for x in seqs:
            for y in seqs:
                pair = SequencePair(x, y)
                if align:
                    pair = oracle_align(pair)
                    pairs_file.write(pair)
                row = []
                for metric in metrics:
                    d = oracle_metric(metric, pair.x, pair.y) if pair.x != pair.y else None
                    if d is not None and multiply:
                        d *= 100
                    row.append(Distance(metric, pair.x, pair.y, d))
                for k, d in enumerate(row):
                    linear.write(d)
                    matrices[k].write(d)
                subset = {}
                for name, part in (("genera", genera), ("species", species)):
                    if part is None:
                        subset[name] = None
                        continue
                    sx, sy = part.get(x.id, None), part.get(y.id, None)
                    subset[name] = (sx, sy)
                    for d in row:
                        a = aggs[name][str(d.metric)].setdefault((sx, sy), [0.0, inf, 0.0, 0])
                        if d.d is not None:
                            a[0] += d.d; a[1] = min(a[1], d.d); a[2] = max(a[2], d.d); a[3] += 1
                if not wrote_summary_header:
                    summary.write(("seqid (query 1)", "seqid (query 2)", *(str(m) for m in metrics),
                                   *(k + " (query 1)" for k in pair.x.extras), *(k + " (query 2)" for k in pair.y.extras),
                                   "genus (query 1)", "species (query 1)", "genus (query 2)", "species (query 2)", "comparison_type"))
                    wrote_summary_header = True
                g, s = subset["genera"], subset["species"]
                same_g = bool(g[0] == g[1]) if g else None
                same_s = bool(s[0] == s[1]) if s else None
                summary.write((pair.x.id, pair.y.id, *(text(d.d) for d in row),
                               *(missing if v is None else v for v in pair.x.extras.values()),
                               *(missing if v is None else v for v in pair.y.extras.values()),
                               (g[0] if g else "-") or "-", (s[0] if s else "-") or "-",
                               (g[1] if g else "-") or "-", (s[1] if s else "-") or "-", COMPARISON[(same_g, same_s)]))
        for m in matrices:
            m.close()
    if not align:
        (work / "align" / "aligned_pairs.txt").unlink()
        (work / "align").rmdir()
    for name, part in (("genera", genera), ("species", species)):
        if part is None:
            continue
        base = work / "subsets" / name
        (base / "linear").mkdir(parents=True, exist_ok=True)
        (base / "matricial").mkdir(parents=True, exist_ok=True)
        stat = lambda a: (None, None, None) if not a[3] else (a[0] / a[3], a[1], a[2])  # noqa: E731  mean, min, max
        ftext = lambda v: "NA" if v is None else fmt.format(v)  # noqa: E731
        keys = list(next(iter(aggs[name].values())).keys())
        with FileHandler.Tabfile(base / "linear" / "pairs.tsv", "w") as pf, FileHandler.Tabfile(base / "linear" / "identity.tsv", "w") as idf:
            heads = [f"{m} {s}" for m in metrics for s in ("mean", "min", "max")]
            wrote = [False, False]
            for key in keys:
                vals = [ftext(v) for m in metrics for v in stat(aggs[name][str(m)][key])]
                q = lambda v: "?" if v is None else v  # noqa: E731
                if key[0] == key[1]:
                    if not wrote[0]:
                        idf.write(("target", *heads)); wrote[0] = True
                    idf.write((q(key[0]), *vals))
                else:
                    if not wrote[1]:
                        pf.write(("target", "query", *heads)); wrote[1] = True
                    pf.write((q(key[0]), q(key[1]), *vals))
        for m in metrics:
            with FileHandler.Tabfile(base / "matricial" / f"{m}.tsv", "w") as f:
                rows = [(k, list(grp)) for k, grp in groupby(keys, key=lambda kk: kk[0])]
                header_done = False
                for idx, grp in rows:
                    if not header_done:
                        f.write(("", *("?" if k[1] is None else k[1] for k in grp))); header_done = True
                    cells = []
                    for k in grp:
                        a = aggs[name][str(m)][k]
                        if not a[3]:
                            cells.append("NA")
                        else:
                            mean, mn, mx = stat(a)
                            cells.append(f"{ftext(mean)} ({ftext(mn)}-{ftext(mx)})")
                    f.write(("?" if idx is None else idx, *cells))


def versus_reference(data, reference, work: Path, align=True, metric=None, extra=None, fmt="{:.4f}", missing="NA", multiply=False):
    metric = metric or DistanceMetric.Uncorrected()
    extra = extra if extra is not None else [DistanceMetric.UncorrectedWithGaps(), DistanceMetric.JukesCantor(), DistanceMetric.Kimura2P()]
    extra = [m for m in extra if m != metric]
    work = Path(work)
    (work / "distances").mkdir(parents=True, exist_ok=True)
    data = [s.normalize() for s in data] if align else list(data)
    reference = [s.normalize() for s in reference] if align else list(reference)

    def distances():
        with SequencePairHandler.Formatted(work / "aligned_pairs.txt", "w") as pairs_file, \
                DistanceHandler.Linear.WithExtras(work / "distances" / f"{metric}.linear.tsv", "w", missing=missing, formatter=fmt) as linear, \
                DistanceHandler.Matrix(work / "distances" / f"{metric}.matricial.tsv", "w", missing=missing, formatter=fmt) as matrix:
            for x in data:
                for y in reference:
                    pair = SequencePair(x, y)
                    if align:
                        pair = oracle_align(pair)
                        pairs_file.write(pair)
                    d = oracle_metric(metric, pair.x, pair.y)
                    if d is not None and multiply:
                        d *= 100
                    dist = Distance(metric, pair.x, pair.y, d)
                    linear.write(dist)
                    matrix.write(dist)
                    yield dist

    with DistanceHandler.Linear.WithExtras(work / "closest.tsv", "w", missing=missing, formatter=fmt) as closest:
        for _, group in groupby(distances(), lambda d: d.x.id):
            best = min((d for d in group if d.d is not None), key=lambda d: d.d)
            closest.write(best)
            for m in extra:
                d = oracle_metric(m, best.x, best.y)
                if d is not None and multiply:
                    d *= 100
                closest.write(Distance(m, best.x, best.y, d))
    if not align:
        (work / "aligned_pairs.txt").unlink()


# ---- dereplicate / decontaminate (dereplicate.py:393-440, decontaminate.py:336-371, decontaminate2.py) ----
class PlainSequenceTab(PlainTab):
    """sequences.py:211-234 with idHeader="seqid", seqHeader="sequence": header seqid, <extras>,
    sequence from the first record (seqid, sequence alone for an empty file), then one row each."""

    def __init__(self, path):
        super().__init__(path)
        self.header_done = False

    def write(self, sequence):
        if not self.header_done:
            PlainTab.write(self, ["seqid", *sequence.extras.keys(), "sequence"])
            self.header_done = True
        PlainTab.write(self, [sequence.id, *sequence.extras.values(), sequence.seq])

    def close(self):
        if not self.header_done:
            PlainTab.write(self, ["seqid", "sequence"])
        super().close()


class PlainFasta:
    """sequences.py:97-114 with write_organism=True: ">id|organism" (the organism only when the
    record has one), the sequence in 60-column lines, one empty line after every record."""

    def __init__(self, path):
        self.file = open(path, "w", newline="")

    def write(self, sequence):
        title = sequence.id
        if organism := sequence.extras.get("organism", None):
            title += "|" + organism
        self.file.write(">" + title + "\n")
        for k in range(0, len(sequence.seq), 60):
            self.file.write(sequence.seq[k:k + 60] + "\n")
        self.file.write("\n")

    def close(self):
        self.file.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _out_handler(path: Path, fasta: bool):
    return PlainFasta(path) if fasta else PlainSequenceTab(path)


def dereplicate(sequences, work: Path, similarity=0.07, length=10, align=True, metric=None, fmt="{:.4f}", missing="NA",
                multiply=False, fasta=False):
    metric = metric or DistanceMetric.Uncorrected()
    work = Path(work)
    (work / "distances").mkdir(parents=True, exist_ok=True)
    ext = ".fas" if fasta else ".tsv"
    data = [s for s in sequences if len(s.seq) >= length]
    excluded = set()
    text = lambda v: missing if v is None else fmt.format(v)  # noqa: E731

    def infos(pairs_file, linear, matrix):
        for x in data:
            for y in data:
                if x.id == y.id:
                    continue
                if x.id in excluded or y.id in excluded:
                    continue
                pair = SequencePair(x, y)
                if align:
                    pair = oracle_align(SequencePair(x.normalize(), y.normalize()))
                    pairs_file.write(pair)
                d = oracle_metric(metric, pair.x, pair.y)
                if d is not None and multiply:
                    d *= 100
                dist = Distance(metric, pair.x, pair.y, d)
                linear.write(dist)
                matrix.write(dist)
                yield (x.id, y.id, len(x.seq), len(y.seq), d, False if d is None else bool(d <= similarity))

    with SequencePairHandler.Formatted(work / "aligned_pairs.txt", "w") as pairs_file, \
            DistanceHandler.Linear.WithExtras(work / "distances" / f"{metric}.linear.tsv", "w", missing=missing, formatter=fmt) as linear, \
            DistanceHandler.Matrix(work / "distances" / f"{metric}.matricial.tsv", "w", missing=missing, formatter=fmt) as matrix, \
            FileHandler.Tabfile(work / "summary.tsv", "w", columns=("query_id", "query_length", "included_id", "included_length",
                                                                    "included_distance", "excluded_id", "excluded_length",
                                                                    "excluded_distance")) as summary:
        for _, group in groupby(infos(pairs_file, linear, matrix), lambda t: t[0]):
            first = next(group)
            query_id, query_length = first[0], first[2]
            max_id, max_length, max_distance = first[0], first[2], first[4]
            from itertools import chain
            for _, id_y, _, len_y, distance, similar in chain([first], group):
                if not similar:
                    continue
                if len_y > max_length:
                    inc, exc = (id_y, len_y, distance), (max_id, max_length, max_distance)
                else:
                    inc, exc = (max_id, max_length, max_distance), (id_y, len_y, distance)
                excluded.add(exc[0])
                summary.write((query_id, str(query_length), inc[0], str(inc[1]), text(inc[2]), exc[0], str(exc[1]), text(exc[2])))
                if len_y > max_length:
                    max_id, max_length, max_distance = id_y, len_y, distance
    if not align:
        (work / "aligned_pairs.txt").unlink()
    with _out_handler(work / f"dereplicated{ext}", fasta) as kept, _out_handler(work / f"excluded{ext}", fasta) as dropped:
        for s in data:
            (dropped if s.id in excluded else kept).write(s)
    return excluded


def _group_minimums(data, group, align, metric, multiply, pairs_path, linear_path, matrix_path, fmt, missing):
    xs = [s.normalize() for s in data] if align else list(data)
    ys = [s.normalize() for s in group] if align else list(group)
    pairs_path.parent.mkdir(parents=True, exist_ok=True)
    linear_path.parent.mkdir(parents=True, exist_ok=True)

    def distances():
        with SequencePairHandler.Formatted(pairs_path, "w") as pairs_file, \
                DistanceHandler.Linear.WithExtras(linear_path, "w", missing=missing, formatter=fmt) as linear, \
                DistanceHandler.Matrix(matrix_path, "w", missing=missing, formatter=fmt) as matrix:
            for x in xs:
                for y in ys:
                    pair = SequencePair(x, y)
                    if align:
                        pair = oracle_align(pair)
                        pairs_file.write(pair)
                    d = oracle_metric(metric, pair.x, pair.y)
                    if d is not None and multiply:
                        d *= 100
                    dist = Distance(metric, pair.x, pair.y, d)
                    linear.write(dist)
                    matrix.write(dist)
                    yield dist

    out = [min(grp, key=lambda d: d.d if d.d is not None else inf) for _, grp in groupby(distances(), lambda d: d.x.id)]
    if not align:
        pairs_path.unlink()
    return out


def decontaminate(data, outgroup, work: Path, similarity=0.07, align=True, metric=None, fmt="{:.4f}", missing="NA",
                  multiply=False, fasta=False):
    metric = metric or DistanceMetric.Uncorrected()
    work = Path(work)
    ext = ".fas" if fasta else ".tsv"
    text = lambda v: missing if v is None else fmt.format(v)  # noqa: E731
    mins = _group_minimums(data, outgroup, align, metric, multiply, work / "aligned_pairs.txt",
                           work / "distances" / f"{metric}.linear.tsv", work / "distances" / f"{metric}.matricial.tsv", fmt, missing)
    with FileHandler.Tabfile(work / "summary.tsv", "w", columns=("query_id", "outgroup_id", "outgroup_distance", "contaminant")) as summary, \
            _out_handler(work / f"decontaminated{ext}", fasta) as clean, _out_handler(work / f"contaminants{ext}", fasta) as dirty:
        for s, best in zip(data, mins):
            bad = False if best.d is None else bool(best.d <= similarity)
            (dirty if bad else clean).write(s)
            summary.write((s.id, best.y.id, text(best.d), "Yes" if bad else "No"))


def decontaminate2(data, outgroup, ingroup, work: Path, w_out=1.0, w_in=1.0, align=True, metric=None, fmt="{:.4f}", missing="NA",
                   multiply=False, fasta=False):
    metric = metric or DistanceMetric.Uncorrected()
    work = Path(work)
    ext = ".fas" if fasta else ".tsv"
    text = lambda v: missing if v is None else fmt.format(v)  # noqa: E731
    (work / "aligned_pairs").mkdir(parents=True, exist_ok=True)
    outs = _group_minimums(data, outgroup, align, metric, multiply, work / "aligned_pairs" / "outgroup.txt",
                           work / "distances" / f"outgroup.{metric}.linear.tsv", work / "distances" / f"outgroup.{metric}.matricial.tsv", fmt, missing)
    ins = _group_minimums(data, ingroup, align, metric, False, work / "aligned_pairs" / "ingroup.txt",
                          work / "distances" / f"ingroup.{metric}.linear.tsv", work / "distances" / f"ingroup.{metric}.matricial.tsv", fmt, missing)
    if not align:
        (work / "aligned_pairs").rmdir()
    with FileHandler.Tabfile(work / "summary.tsv", "w", columns=("query_id", "outgroup_id", "outgroup_distance", "ingroup_id",
                                                                 "ingroup_distance", "contaminant")) as summary, \
            _out_handler(work / f"decontaminated{ext}", fasta) as clean, _out_handler(work / f"contaminants{ext}", fasta) as dirty:
        for s, bo, bi in zip(data, outs, ins):
            do = None if bo.d is None else bo.d * w_out
            di = None if bi.d is None else bi.d * w_in
            bad = False if do is None else (True if di is None else bool(do < di))
            (dirty if bad else clean).write(s)
            summary.write((s.id, bo.y.id, text(do), bi.y.id, text(di), "Yes" if bad else "No"))
