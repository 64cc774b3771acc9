"""Per-sequence-set statistics (`stats/{all,species,genera}.tsv` of VersusAll).

Host-side O(N*L) byte counting, off the O(N^2) hot path; present so that the VersusAll drop-in
writes every file the reference task writes.  Behaviour follows
/root/reference/src/itaxotools/taxi2/statistics.py: the 26 labelled statistics and their order
(:45-73, labels keep the reference's trailing blanks), counting rules (:25-37: nucleotides = length
minus '-', missing = 'N' only, upper case expected), length classes (:146-155), population standard
deviation / median of the nucleotide counts (:171-175), N50/L50/N90/L90 (:215-224) and the two
writers `StatisticsHandler.Single` / `.Groups` (:258-313).  Pinned by the reference's own vectors
and fixtures (tests/test_statistics.py:113-250, tests/test_statistics/*), see tests/test_statistics.py.
"""
from __future__ import annotations

import statistics as pystat
from enum import Enum
from typing import Iterable, NamedTuple

from .handlers import FileHandler
from .types import Percentage


class Counts(NamedTuple):
    total: int
    nucleotides: int
    missing: int
    gaps: int
    a: int
    c: int
    g: int
    t: int

    @classmethod
    def from_sequence(cls, seq: str) -> "Counts":
        gaps = seq.count("-")
        return cls(len(seq), len(seq) - gaps, seq.count("N"), gaps, seq.count("A"), seq.count("C"), seq.count("G"), seq.count("T"))


class NL(NamedTuple):
    N: int
    L: int


class Statistic(Enum):
    """Label and value type of every statistic, in output order."""

    Group = "Group", str
    SequenceCount = "Total number of sequences", int
    NucleotideCount = "Total length of all sequences ", int
    BP_0 = "Number of sequences with 0 bp", int
    BP_1_100 = "Number of sequences with less than 100 bp", int
    BP_101_300 = "Number of sequences between 101-300 bp", int
    BP_301_1000 = "Number of sequences between 301-1000 bp", int
    BP_1001_plus = "Number of sequences with more than 1000 bp", int
    Minimum = "Minimum sequence length", int
    Maximum = "Maximum sequence length ", int
    Mean = "Mean sequence length  ", float
    Median = "Median sequence length  ", float
    Stdev = "Standard deviation of sequence length", float
    PercentA = "Percentage of base A", Percentage
    PercentC = "Percentage of base C", Percentage
    PercentG = "Percentage of base G", Percentage
    PercentT = "Percentage of base T", Percentage
    PercentGC = "GC content", Percentage
    PercentAmbiguous = "Percentage of ambiguity codes", Percentage
    PercentMissing = "Percentage of missing data ", Percentage
    PercentMissingGaps = "Percentage of missing data including gaps", Percentage
    PercentGaps = "Percentage of gaps", Percentage
    N50 = "N50 statistic", int
    L50 = "L50 statistic", int
    N90 = "N90 statistic", int
    L90 = "L90 statistic", int

    def __init__(self, label, type):
        self.label = label
        self.type = type

    def __repr__(self):
        return f"<{type(self).__name__}.{self._name_}>"

    def __str__(self):
        return self.label


class Statistics(dict):
    """Statistic -> value, always in `Statistic` order, values coerced to the statistic's type."""

    def __init__(self, stats: dict):
        super().__init__({s: s.type(stats[s]) for s in Statistic if s in stats})

    @classmethod
    def from_sequences(cls, sequences: Iterable[str], group: str = None) -> "Statistics":
        return StatisticsCalculator(sequences, group).calculate()


class StatisticsCalculator:
    """Accumulates sequences; `calculate()` may be called once.  After it, `add` and `calculate`
    raise StopIteration like the reference's exhausted generator (tests/test_statistics.py:263-274)."""

    _CLASSES = (0, 100, 300, 1000)   # upper bounds of BP_0, BP_1_100, BP_101_300, BP_301_1000

    def __init__(self, sequences: Iterable[str] = (), group: str = None):
        self.group = group
        self.lengths: list[int] = []
        self.classes = [0, 0, 0, 0, 0]
        self.total = self.missing = self.gaps = self.a = self.c = self.g = self.t = 0
        self.finished = False
        for seq in sequences:
            self.add(seq)

    def add(self, seq: str) -> None:
        if self.finished:
            raise StopIteration
        k = Counts.from_sequence(seq)
        self.lengths.append(k.nucleotides)
        self.classes[next((i for i, bound in enumerate(self._CLASSES) if k.nucleotides <= bound), 4)] += 1
        self.total += k.total
        self.missing += k.missing
        self.gaps += k.gaps
        self.a += k.a
        self.c += k.c
        self.g += k.g
        self.t += k.t

    @staticmethod
    def _calculate_NL(counts: list[int], arg: int = 50) -> NL:
        """N = length of the contig at which the descending cumulative sum reaches arg % of the
        total, L = how many contigs that took."""
        if not any(counts):
            return NL(0, 0)
        ordered = sorted(counts, reverse=True)
        target = sum(ordered) * arg / 100
        running = 0
        for pos, v in enumerate(ordered):
            running += v
            if running >= target:
                return NL(v, pos + 1)
        raise AssertionError("cumulative sum never reached its own total")

    def calculate(self) -> Statistics:
        if self.finished:
            raise StopIteration
        self.finished = True
        lengths = self.lengths
        count = len(lengths)
        nucleotides = sum(lengths)
        share = (lambda v: v / nucleotides) if nucleotides else (lambda v: 0)
        share_total = (lambda v: v / self.total) if self.total else (lambda v: 0)
        n50, l50 = self._calculate_NL(lengths, 50)
        n90, l90 = self._calculate_NL(lengths, 90)
        out = {
            Statistic.SequenceCount: count,
            Statistic.NucleotideCount: nucleotides,
            Statistic.BP_0: self.classes[0],
            Statistic.BP_1_100: self.classes[1],
            Statistic.BP_101_300: self.classes[2],
            Statistic.BP_301_1000: self.classes[3],
            Statistic.BP_1001_plus: self.classes[4],
            Statistic.Minimum: min(lengths) if count else 0,
            Statistic.Maximum: max(lengths) if count else 0,
            Statistic.Mean: nucleotides / count if count else 0,
            Statistic.Median: pystat.median(lengths) if count else 0,
            Statistic.Stdev: pystat.pstdev(lengths) if count > 1 else 0,
            Statistic.PercentA: share(self.a),
            Statistic.PercentC: share(self.c),
            Statistic.PercentG: share(self.g),
            Statistic.PercentT: share(self.t),
            Statistic.PercentGC: share(self.c + self.g),
            Statistic.PercentAmbiguous: share(nucleotides - self.missing - self.a - self.t - self.c - self.g),
            Statistic.PercentMissing: share(self.missing),
            Statistic.PercentMissingGaps: share_total(self.missing + self.gaps),
            Statistic.PercentGaps: share_total(self.gaps),
            Statistic.N50: n50,
            Statistic.L50: l50,
            Statistic.N90: n90,
            Statistic.L90: l90,
        }
        if self.group:
            out[Statistic.Group] = self.group
        return Statistics(out)


class StatisticsHandler(FileHandler):
    """Write-only.  `Single`: one "label<TAB>value" line per statistic; `Groups`: one column per
    statistic, one row per group (statistics.py:227-313)."""

    def _open(self, path, mode="w", float_formatter="{:f}", percentage_formatter="{:f}", percentage_multiply=False, *args, **kwargs):
        self.formatters = {float: float_formatter, Percentage: percentage_formatter}
        self.percentage_multiply = percentage_multiply
        super()._open(path, mode, *args, **kwargs)

    def _iter_read(self):
        raise NotImplementedError()

    def statisticToText(self, value) -> str:
        if isinstance(value, Percentage) and self.percentage_multiply:
            value = Percentage(value * 100)
        return self.formatters.get(type(value), "{}").format(value)


class Single(StatisticsHandler):
    def _iter_write(self):
        with FileHandler.Tabfile(self.path, "w") as file:
            try:
                stats = yield
                for stat, value in stats.items():
                    file.write((str(stat), self.statisticToText(value)))
                yield
                raise Exception("Can only write a single statistics instance")
            except GeneratorExit:
                return


class Groups(StatisticsHandler):
    def _open(self, path, mode="w", group_name="group", *args, **kwargs):
        self.group_name = group_name
        super()._open(path, mode, *args, **kwargs)

    def _iter_write(self):
        with FileHandler.Tabfile(self.path, "w") as file:
            try:
                first = True
                while True:
                    stats = yield
                    if Statistic.Group not in stats:
                        raise Exception("Statistics must contain a group name")
                    if first:
                        file.write((self.group_name, *[str(s) for s in stats][1:]))
                        first = False
                    file.write(tuple(self.statisticToText(v) for v in stats.values()))
            except GeneratorExit:
                return
