"""Partitions (individual id -> subset) as far as the distance tasks need them.

Mirrors /root/reference/src/itaxotools/taxi2/partitions.py:15-110 for the tabular formats and
:127-155 for FASTA titles ("individual|subset"); the Spart / Excel partition readers are out of
scope (third-party parsers, not on the compute path).
"""
from __future__ import annotations

from abc import abstractmethod
from pathlib import Path
from typing import Callable, Literal, NamedTuple

from .handlers import FileHandler, ReadHandle, WriteHandle
from .sequences import _fasta_records


class Classification(NamedTuple):
    individual: str
    subset: str


class Partition(dict):
    """Keys are individuals, values are subsets"""

    @classmethod
    def fromPath(cls, path: Path, handler: "PartitionHandler", *args, **kwargs) -> "Partition":
        return handler.as_dict(path, *args, **kwargs)


class PartitionHandler(FileHandler[Classification]):
    """Read-only handlers yielding Classification(individual, subset); an optional `filter`
    rewrites every classification or drops it by returning None (partitions.py:30-66)."""

    filter: Callable | None = None

    @classmethod
    def as_dict(cls, path: Path, *args, **kwargs) -> Partition:
        # later rows win, as in a dict built by assignment
        return Partition({c.individual: c.subset for c in map(Classification._make, cls(path, "r", *args, **kwargs))})

    def _open(self, path: Path, mode: Literal["r", "w"] = "r", filter: Callable = None, *args, **kwargs):
        self.filter = filter
        super()._open(path, mode, *args, **kwargs)

    def _iter_write(self) -> WriteHandle[Classification]:
        raise NotImplementedError("partitions are read-only here")

    def _iter_read(self, *args, **kwargs) -> ReadHandle[Classification]:
        source = self._iter_read_inner(*args, **kwargs)
        yield next(source)                     # the handler itself, once the file is open
        if self.filter is None:
            yield from source
            return
        yield from (kept for kept in map(self.filter, source) if kept is not None)

    @abstractmethod
    def _iter_read_inner(self, *args, **kwargs) -> ReadHandle[Classification]:
        yield self

    @staticmethod
    def subset_first_word(classification: Classification) -> Classification | None:
        """Genus = first word of the organism column (partitions.py:67-75)."""
        individual, subset = classification
        first, sep, _ = subset.partition(" ")
        if not sep:
            print(f"Cannot split subset {subset} for individual {individual}")
            return None
        return Classification(individual, first)


class Tabular(PartitionHandler):
    subhandler = FileHandler.Tabular

    def _iter_read_inner(self, idHeader: str = None, subHeader: str = None, hasHeader: bool = False,
                         idColumn: int = 0, subColumn: int = 1) -> ReadHandle[Classification]:
        if idHeader and subHeader:
            columns, hasHeader = (idHeader, subHeader), True
        else:
            columns = (idColumn, subColumn)
        with self.subhandler(self.path, has_headers=hasHeader, columns=columns) as rows:
            yield self
            for individual, subset in rows:
                yield Classification(individual, subset)


class Tabfile(Tabular, PartitionHandler):
    subhandler = FileHandler.Tabular.Tabfile


class Fasta(PartitionHandler):
    """Subsets from FASTA titles `>individual<separator>subset`; titles without the separator are
    reported and skipped (partitions.py:127-137)."""

    def _iter_read_inner(self, separator: str = "|") -> ReadHandle[Classification]:
        with open(self.path, "r") as handle:
            yield self
            for title, _ in _fasta_records(handle):
                individual, found, subset = title.partition(separator)
                if not found:
                    print(f"Could not extract partition info from fasta line: {title}")
                    continue
                yield Classification(individual, subset)

    @classmethod
    def has_subsets(cls, path: Path, separator: str = "|") -> bool:
        if not separator:
            return False
        with open(path, "r") as handle:
            for title, _ in _fasta_records(handle):
                return separator in title

    @classmethod
    def guess_subset_separator(cls, path: Path) -> str | None:
        with open(path, "r") as handle:
            for title, _ in _fasta_records(handle):
                for separator in "|.":
                    if separator in title:
                        return separator
            return None
