"""Sequence pairs: the unit of work of the hot path, and the aligned-pairs file formats.

Mirrors /root/reference/src/itaxotools/taxi2/pairs.py (SequencePair/SequencePairs :11-25,
Tabfile :32-48, Formatted :51-97).
"""
from __future__ import annotations

from pathlib import Path
from typing import NamedTuple, TextIO

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


class Tabfile(SequencePairHandler):
    def _iter_read(self) -> ReadHandle[SequencePair]:
        with FileHandler.Tabfile(self.path, "r", has_headers=True) as file:
            yield self
            for idx, idy, seqX, seqY in file:
                yield SequencePair(Sequence(idx, seqX), Sequence(idy, seqY))

    def _iter_write(self) -> WriteHandle[SequencePair]:
        with FileHandler.Tabfile(self.path, "w", columns=["idx", "idy", "seqx", "seqy"]) as file:
            try:
                while True:
                    pair = yield
                    file.write((pair.x.id, pair.y.id, pair.x.seq, pair.y.seq))
            except GeneratorExit:
                return


class Formatted(SequencePairHandler):
    """Four-line blocks: 'idx / idy', aligned x, match pattern, aligned y; blank line between."""

    @staticmethod
    def _format_char(x: str, y: str) -> str:
        if x == "-" or y == "-":
            return "-"
        return "|" if x == y else "."

    @classmethod
    def _format(cls, x: str, y: str) -> str:
        return "".join(map(cls._format_char, x, y))

    def _iter_read(self) -> ReadHandle[SequencePair]:
        with open(self.path, "r") as file:
            yield self
            while True:
                lines = [file.readline().strip() for _ in range(5)]
                if not any(lines):
                    return
                idx, idy = lines[0].split(" / ")
                yield SequencePair(Sequence(idx, lines[1]), Sequence(idy, lines[3]))

    def _iter_write(self) -> WriteHandle[SequencePair]:
        with open(self.path, "w") as file:
            try:
                first = True
                while True:
                    pair = yield
                    if not first:
                        file.write("\n")
                    first = False
                    self._write_lines(file, pair)
            except GeneratorExit:
                return

    def _write_lines(self, file: TextIO, pair: SequencePair) -> None:
        file.write(f"{pair.x.id} / {pair.y.id}\n{pair.x.seq}\n{self._format(pair.x.seq, pair.y.seq)}\n{pair.y.seq}\n")
