"""Sequence pairs: the unit of work of the hot path, and the aligned-pairs file formats.

Mirrors /root/reference/src/itaxotools/taxi2/pairs.py (SequencePair/SequencePairs :11-25,
Tabfile :32-48, Formatted :51-97).
"""
from __future__ import annotations

from pathlib import Path
from typing import Iterator, NamedTuple

import numpy as np

from .handlers import FileHandler, ReadHandle, WriteHandle
from .sequences import Sequence, Sequences
from .types import Container


class SequencePair(NamedTuple):
    x: Sequence
    y: Sequence


class SequencePairs(Container[SequencePair]):
    @classmethod
    def fromPath(cls, path: Path, handler: "SequencePairHandler", *args, **kwargs) -> "SequencePairs":
        return cls(handler, path, *args, **kwargs)

    @classmethod
    def fromProduct(cls, xs: Sequences, ys: Sequences) -> "SequencePairs":
        """Lazy row-major Cartesian product, ys re-iterated for every x (pairs.py:23-25)."""
        return cls(lambda: (SequencePair(x, y) for x in xs for y in ys))


class SequencePairHandler(FileHandler[SequencePair]):
    pass


PAIR_COLUMNS = ("idx", "idy", "seqx", "seqy")


def match_pattern(x: str, y: str) -> str:
    """The middle line of a formatted pair (pairs.py:52-58): '-' where either string has a gap,
    '|' where the two symbols are equal, '.' elsewhere.  Compared as bytes in one numpy pass
    (aligned barcodes are hundreds of columns; the reference maps a Python function over them)."""
    a = np.frombuffer(x.encode("latin-1", "replace"), dtype=np.uint8)
    b = np.frombuffer(y.encode("latin-1", "replace"), dtype=np.uint8)
    n = min(len(a), len(b))          # map() over two strings stops at the shorter one
    a, b = a[:n], b[:n]
    out = np.full(n, ord("."), dtype=np.uint8)
    out[a == b] = ord("|")
    out[(a == ord("-")) | (b == ord("-"))] = ord("-")
    return out.tobytes().decode("latin-1")


def _records(path: Path) -> Iterator[list[str]]:
    """Blocks of stripped lines of a formatted pairs file, five lines apiece (the fifth is the
    separator); the file ends at the first block that is entirely empty."""
    with open(path, "r") as file:
        while True:
            block = [file.readline().strip() for _ in range(5)]
            if not any(block):
                return
            yield block


class Tabfile(SequencePairHandler):
    """idx / idy / seqx / seqy, tab separated, one header row (pairs.py:32-48)."""

    def _iter_read(self) -> ReadHandle[SequencePair]:
        with FileHandler.Tabfile(self.path, "r", has_headers=True) as rows:
            yield self
            for row in rows:
                id_x, id_y, seq_x, seq_y = row
                yield SequencePair(x=Sequence(id_x, seq_x), y=Sequence(id_y, seq_y))

    def _iter_write(self) -> WriteHandle[SequencePair]:
        with FileHandler.Tabfile(self.path, "w", columns=list(PAIR_COLUMNS)) as rows:
            try:
                while True:
                    x, y = yield
                    rows.write((x.id, y.id, x.seq, y.seq))
            except GeneratorExit:
                pass


class Formatted(SequencePairHandler):
    """Four-line records -- 'idx / idy', aligned x, match pattern, aligned y -- separated by one
    empty line (pairs.py:51-97)."""

    _format = staticmethod(match_pattern)

    def _iter_read(self) -> ReadHandle[SequencePair]:
        yield self
        for title, seq_x, _pattern, seq_y, _blank in _records(self.path):
            id_x, id_y = title.split(" / ")
            yield SequencePair(x=Sequence(id_x, seq_x), y=Sequence(id_y, seq_y))

    def _iter_write(self) -> WriteHandle[SequencePair]:
        with open(self.path, "w") as file:
            separator = ""
            try:
                while True:
                    x, y = yield
                    file.write(f"{separator}{x.id} / {y.id}\n{x.seq}\n{match_pattern(x.seq, y.seq)}\n{y.seq}\n")
                    separator = "\n"
            except GeneratorExit:
                pass
