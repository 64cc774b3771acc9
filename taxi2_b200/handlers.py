"""Generator-backed file handlers: the minimal part of the reference's handler runtime that the
pairwise-distance path reads from and writes to.

Mirrors /root/reference/src/itaxotools/taxi2/handlers.py:24-227 (FileHandler, Tabular, Tabfile).
The Excel handler is out of scope (openpyxl is not part of the hot path).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from itertools import chain
from pathlib import Path
from typing import Generator, Generic, Iterator, Literal, TypeVar

from .types import Type, TypeMeta

Item = TypeVar("Item")
ReadHandle = Iterator[Item]
WriteHandle = Generator[None, Item, None]
Row = tuple


class _HandlerMeta(type(ABC), TypeMeta):
    pass


class FileHandler(ABC, Type, Generic[Item], metaclass=_HandlerMeta):
    """Read or write items through a primed generator; mimics io.IOBase.

    Readers `yield self` once they are ready (headers parsed), then yield items.
    Writers are coroutines receiving items with `send`; closing the handler closes the generator,
    which flushes whatever the writer still buffers.
    """

    def __init__(self, *args, **kwargs):
        self._open(*args, **kwargs)
        primed = next(self.it)
        if self.readable() and primed is not self:
            raise Exception("Read handler was not properly primed!")

    def _open(self, path: Path, mode: Literal["r", "w"] = "r", *args, **kwargs):
        self.path = path
        self.mode = mode
        if mode == "r":
            self.it = self._iter_read(*args, **kwargs)
        elif mode == "w":
            self.it = self._iter_write(*args, **kwargs)
        else:
            raise ValueError('Mode must be "r" or "w"')
        self.closed = False

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        self.close()

    def __iter__(self):
        assert self.readable()
        return self

    def __next__(self):
        assert self.readable()
        return next(self.it)

    @abstractmethod
    def _iter_read(self, *args, **kwargs) -> ReadHandle[Item]:
        yield self

    @abstractmethod
    def _iter_write(self, *args, **kwargs) -> WriteHandle[Item]:
        try:
            while True:
                _ = yield
        except GeneratorExit:
            return

    def close(self) -> None:
        self.it.close()
        self.closed = True

    def read(self):
        return next(self.it, None)

    def write(self, item) -> None:
        self.it.send(item)

    def readable(self) -> bool:
        return self.mode == "r"

    def writable(self) -> bool:
        return self.mode == "w"


class Tabular(FileHandler):
    """Rows of strings with optional header row and column selection (handlers.py:106-207)."""

    def _iter_read(self, columns=None, has_headers: bool = False, get_all_columns: bool = False) -> ReadHandle[Row]:
        if columns is not None:
            columns = tuple(columns)
            if not columns:
                raise ValueError("Columns argument must contain at least one item")
            if isinstance(columns[0], str):
                has_headers = True
        self.has_headers = has_headers
        self.header_row = None
        self.column_order = None

        rows = self._iter_read_rows()
        if has_headers:
            self.header_row = next(rows, None)
            if self.header_row is None:
                yield self
                return
        if columns is None:
            yield self
            yield from rows
            return

        if isinstance(columns[0], str):
            missing = set(columns) - set(self.header_row)
            if missing:
                raise ValueError(f"Column header(s) not found in file: {missing}")
            columns = tuple(self.header_row.index(name) for name in columns)
        if get_all_columns:
            if has_headers:
                width = len(self.header_row)
            else:
                first = next(rows)
                rows = chain([first], rows)
                width = len(first)
            columns = columns + tuple(set(range(width)) - set(columns))
        self.column_order = columns
        yield self
        for row in rows:
            yield tuple(row[k] for k in columns)

    def _iter_write(self, columns=None) -> WriteHandle[Row]:
        sink = self._iter_write_rows()
        next(sink)
        if columns is not None:
            columns = tuple(columns)
            if not columns:
                raise ValueError("Columns argument must contain at least one item")
            sink.send(columns)
        try:
            while True:
                row = yield
                sink.send(row)
        except GeneratorExit:
            sink.close()
            return

    @property
    def headers(self):
        assert self.readable()
        if not self.has_headers:
            return None
        if self.column_order:
            return tuple(self.header_row[k] for k in self.column_order)
        return self.header_row

    @classmethod
    def get_headers(cls, path: Path):
        with cls(path) as handler:
            return handler.read()

    @abstractmethod
    def _iter_read_rows(self) -> Iterator[Row]:
        return iter(())

    @abstractmethod
    def _iter_write_rows(self) -> Generator[None, Row, None]:
        yield


class Tabfile(Tabular, FileHandler):
    """Tab-separated text: utf-8 with surrogateescape on read, '\\n' rows, empty lines skipped."""

    def _iter_read_rows(self) -> Iterator[Row]:
        with open(self.path, "r", encoding="utf-8", errors="surrogateescape") as file:
            for line in file:
                line = line[:-1]
                if line:
                    yield tuple(line.split("\t"))

    def _iter_write_rows(self) -> Generator[None, Row, None]:
        with open(self.path, "w") as file:
            try:
                while True:
                    row = yield
                    file.write("\t".join(row) + "\n")
            except GeneratorExit:
                return
