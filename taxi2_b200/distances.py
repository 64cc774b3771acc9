"""Distance metrics and distance file handlers: drop-in for
/root/reference/src/itaxotools/taxi2/distances.py (metrics :282-348, handlers :19-279).

The four metrics on the hot path (p, p-gaps, jc, k2p) are computed on the GPU from the
same / transition / transversion / gap-column counts of the two (aligned) strings; undefined
values come back as NaN and are mapped to None exactly as `_is_number` does (:290-292).
NCD and BBC (alfpy, alignment-free compression / k-mer statistics) are out of scope.
"""
from __future__ import annotations

import re
import threading
from math import isinf, isnan
from pathlib import Path
from typing import Generator, Iterable, Literal, NamedTuple

import numpy as np

from .handlers import FileHandler, ReadHandle, WriteHandle
from .sequences import Sequence
from .types import Container, Type


class Distance(NamedTuple):
    metric: "DistanceMetric"
    x: Sequence
    y: Sequence
    d: float | None


class Distances(Container[Distance]):
    @classmethod
    def fromPath(cls, path: Path, handler: "DistanceHandler", *args, **kwargs) -> "Distances":
        return cls(handler, path, *args, **kwargs)


# ---------------------------------------------------------------------------------------------
# metrics
# ---------------------------------------------------------------------------------------------
class DistanceMetric(Type):
    """Metrics for calculating distances"""

    label: str
    column: int | None = None  # index into the engine's (p, p-gaps, jc, k2p) output

    def __str__(self):
        return self.label

    @staticmethod
    def _is_number(x) -> bool:
        return not (x is None or isnan(x) or isinf(x))

    # the four metrics of the pair seen last: every caller asks for them back to back on the same
    # two strings (versus_all.py:546-552 and siblings), and one device call yields all four.
    # Per thread: two threads computing distances never see each other's pair.
    _last = threading.local()

    def _calculate(self, x: str, y: str) -> float | None:
        if self.column is None:
            raise NotImplementedError()
        last = DistanceMetric._last
        if getattr(last, "pair", None) != (x, y):
            from .engine import default_engine

            eng = default_engine()
            eng.load([x, y], 0)
            last.values = eng.count_pairs([0], [1], want=("metrics",))["metrics"][0].copy()
            last.pair = (x, y)
        value = float(last.values[self.column])
        return value if self._is_number(value) else None

    def calculate(self, x: Sequence, y: Sequence) -> Distance:
        return Distance(self, x, y, self._calculate(x.seq, y.seq))

    @classmethod
    def calculate_batch(cls, metrics: Iterable["DistanceMetric"], pairs) -> list[Distance]:
        """All `metrics` for all `pairs` (aligned SequencePair objects) in one device launch;
        returns Distances in the reference's order: pair-major, metric-minor (versus_all.py:546-552)."""
        from .engine import default_engine

        metrics = list(metrics)
        pairs = list(pairs)
        if not pairs:
            return []
        index: dict[str, int] = {}
        px = np.array([index.setdefault(p.x.seq, len(index)) for p in pairs], dtype=np.int32)
        py = np.array([index.setdefault(p.y.seq, len(index)) for p in pairs], dtype=np.int32)
        eng = default_engine()
        eng.load(list(index), 0)
        values = eng.count_pairs(px, py, want=("metrics",))["metrics"]
        out = []
        for k, pair in enumerate(pairs):
            for metric in metrics:
                v = float(values[k, metric.column])
                out.append(Distance(metric, pair.x, pair.y, v if cls._is_number(v) else None))
        return out

    @classmethod
    def fromLabel(cls, label: str):
        label_arg = None
        res = re.search(r"(\w+)\((\d+)\)", label)
        if res:
            label = res.group(1) + "({})"
            label_arg = res.group(2)
        for child in cls:
            if label == child.label:
                return child(int(label_arg)) if label_arg else child()
        return None


class Unknown(DistanceMetric):
    label = "?"


class Uncorrected(DistanceMetric):
    label = "p"
    column = 0


class UncorrectedWithGaps(DistanceMetric):
    label = "p-gaps"
    column = 1


class JukesCantor(DistanceMetric):
    label = "jc"
    column = 2


class Kimura2P(DistanceMetric):
    label = "k2p"
    column = 3


class NCD(DistanceMetric):
    """Label carrier only: the compression distance (alfpy) is out of scope; files naming it still parse."""

    label = "ncd"

    def _calculate(self, x: str, y: str) -> float:
        raise NotImplementedError("NCD is outside the B200 hot path (SURVEY.md section 2)")


class BBC(DistanceMetric):
    """Label carrier only: base-base correlation (alfpy) is out of scope; files naming it still parse."""

    label = "bbc({})"

    def __init__(self, k=10):
        self.k = k

    def __str__(self):
        return self.label.format(self.k)

    def __eq__(self, other):
        return super().__eq__(other) and self.k == other.k

    def __hash__(self):
        return hash((type(self), self.k))

    def _calculate(self, x: str, y: str) -> float:
        raise NotImplementedError("BBC is outside the B200 hot path (SURVEY.md section 2)")


# ---------------------------------------------------------------------------------------------
# file handlers
# ---------------------------------------------------------------------------------------------
class DistanceHandler(FileHandler[Distance]):
    def _open(self, path: Path, mode: Literal["r", "w"] = "r", missing: str = "NA", formatter: str = "{:f}",
              *args, **kwargs):
        self.missing = missing
        self.formatter = formatter
        super()._open(path, mode, *args, **kwargs)

    def distanceFromText(self, text: str) -> float | None:
        return None if text == self.missing else float(text)

    def distanceToText(self, d: float | None) -> str:
        return self.missing if d is None else self.formatter.format(d)


class _LineWriter:
    """Shared writer skeleton: distances are buffered until the row key changes, the header is
    derived from the first complete row (distances.py:76-118, 158-186)."""

    def _same_row(self, a: Distance, b: Distance) -> bool:
        raise NotImplementedError()

    def _header(self, line: list[Distance]) -> tuple:
        raise NotImplementedError()

    def _row(self, line: list[Distance]) -> tuple:
        raise NotImplementedError()

    def _iter_write(self) -> WriteHandle[Distance]:
        self.buffer: list[Distance] = []
        self.wrote_headers = False
        with FileHandler.Tabfile(self.path, "w") as file:
            try:
                while True:
                    distance = yield
                    if self.buffer and not self._same_row(self.buffer[0], distance):
                        self._flush(file)
                    self.buffer.append(distance)
            except GeneratorExit:
                if self.buffer:
                    self._flush(file)
                return

    def _flush(self, file) -> None:
        if not self.wrote_headers:
            file.write(self._header(self.buffer))
            self.wrote_headers = True
        file.write(self._row(self.buffer))
        self.buffer = []


class Linear(_LineWriter, DistanceHandler):
    def _iter_read(self) -> ReadHandle[Distance]:
        with FileHandler.Tabfile(self.path, "r", has_headers=True) as file:
            if file.headers is None:
                yield self
                return
            metrics = [DistanceMetric.fromLabel(label) for label in file.headers[2:]]
            yield self
            for row in file:
                x, y = Sequence(row[0], None), Sequence(row[1], None)
                for text, metric in zip(row[2:], metrics):
                    yield Distance(metric, x, y, self.distanceFromText(text))

    def _same_row(self, a: Distance, b: Distance) -> bool:
        return a.x.id == b.x.id and a.y.id == b.y.id

    def _header(self, line):
        return ("idx", "idy", *(str(d.metric) for d in line))

    def _row(self, line):
        return (line[0].x.id, line[0].y.id, *(self.distanceToText(d.d) for d in line))


class Matrix(_LineWriter, DistanceHandler):
    def _iter_read(self, metric: DistanceMetric = None) -> ReadHandle[Distance]:
        metric = metric or DistanceMetric.Unknown()
        with FileHandler.Tabfile(self.path, "r", has_headers=True) as file:
            if file.headers is None:
                yield self
                return
            idys = file.headers[1:]
            yield self
            for row in file:
                x = Sequence(row[0], None)
                for text, idy in zip(row[1:], idys):
                    yield Distance(metric, x, Sequence(idy, None), self.distanceFromText(text))

    def _same_row(self, a: Distance, b: Distance) -> bool:
        return a.x.id == b.x.id

    def _header(self, line):
        return ("", *(d.y.id for d in line))

    def _row(self, line):
        return (line[0].x.id, *(self.distanceToText(d.d) for d in line))


class WithExtras(DistanceHandler.Linear):
    def _iter_read(self, idxHeader: str = None, idyHeader: str = None, tagX: str = " (query)",
                   tagY: str = " (reference)", idxColumn: int = 0, idyColumn: int = 1) -> ReadHandle[Distance]:
        with FileHandler.Tabfile(self.path, "r", has_headers=True) as file:
            if file.headers is None:
                yield self
                return
            headers = file.headers
            if idxHeader and idyHeader:
                idxColumn = headers.index(idxHeader + tagX)
                idyColumn = headers.index(idyHeader + tagY)
            first_metric = next((k for k, h in enumerate(headers) if DistanceMetric.fromLabel(h)), None)
            if first_metric is None:
                raise Exception("No metrics found in the header line!")
            slice_x, slice_y = slice(idxColumn + 1, idyColumn), slice(idyColumn + 1, first_metric)
            metrics = [DistanceMetric.fromLabel(h) for h in headers[first_metric:]]
            keys_x = [h.removesuffix(tagX) for h in headers[slice_x]]
            keys_y = [h.removesuffix(tagY) for h in headers[slice_y]]
            yield self
            for row in file:
                x = Sequence(row[idxColumn], None, dict(zip(keys_x, row[slice_x])))
                y = Sequence(row[idyColumn], None, dict(zip(keys_y, row[slice_y])))
                for text, metric in zip(row[first_metric:], metrics):
                    yield Distance(metric, x, y, self.distanceFromText(text))

    def _iter_write(self, idxHeader: str = "seqid", idyHeader: str = "seqid", tagX: str = " (query)",
                    tagY: str = " (reference)") -> WriteHandle[Distance]:
        self.idxHeader, self.idyHeader, self.tagX, self.tagY = idxHeader, idyHeader, tagX, tagY
        yield from super()._iter_write()

    def _header(self, line):
        x, y = line[0].x, line[0].y
        return (self.idxHeader + self.tagX, *(k + self.tagX for k in x.extras),
                self.idyHeader + self.tagY, *(k + self.tagY for k in y.extras),
                *(str(d.metric) for d in line))

    def _row(self, line):
        x, y = line[0].x, line[0].y
        fill = lambda values: [self.missing if v is None else v for v in values]  # noqa: E731
        return (x.id, *fill(x.extras.values()), y.id, *fill(y.extras.values()),
                *(self.distanceToText(d.d) for d in line))
