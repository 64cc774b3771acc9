"""File handlers of the pairwise-distance path: objects that read items from, or write items to, a
file through one generator each, opened like a file (`Handler(path, "r" | "w", ...)`), usable as
context managers and iterators.

Interface of /root/reference/src/itaxotools/taxi2/handlers.py:24-227 (FileHandler, Tabular,
Tabfile) -- the subclasses in sequences.py, pairs.py, partitions.py and distances.py plug their
`_iter_read` / `_iter_write` generators into it exactly as the reference's do -- with the bodies
written for this package.  The Excel handler is out of scope (openpyxl is not on the hot path).

Protocol.  A reader generator yields the handler itself once it has opened the file and parsed
whatever header it needs (so that errors surface in the constructor), then yields items.  A writer
generator is primed up to its first `yield` and receives items through `send`; closing the
handler closes the generator, which is where writers flush.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
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
    def __init__(self, *args, **kwargs):
        self._open(*args, **kwargs)
        first = next(self.it)            # readers: runs up to `yield self`; writers: up to their first `yield`
        if self.mode == "r" and first is not self:
            raise Exception("Read handler was not properly primed!")

    def _open(self, path: Path, mode: Literal["r", "w"] = "r", *args, **kwargs):
        makers = {"r": self._iter_read, "w": self._iter_write}
        if mode not in makers:
            raise ValueError('Mode must be "r" or "w"')
        self.path, self.mode, self.closed = path, mode, False
        self.it = makers[mode](*args, **kwargs)

    # -- file-like surface -------------------------------------------------------------------------
    def readable(self) -> bool:
        return self.mode == "r"

    def writable(self) -> bool:
        return self.mode == "w"

    def read(self):
        """The next item, or None at the end."""
        return next(self.it, None)

    def write(self, item) -> None:
        self.it.send(item)

    def close(self) -> None:
        self.it.close()
        self.closed = True

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

    # -- what a concrete handler provides ----------------------------------------------------------
    @abstractmethod
    def _iter_read(self, *args, **kwargs) -> ReadHandle[Item]:
        yield self

    @abstractmethod
    def _iter_write(self, *args, **kwargs) -> WriteHandle[Item]:
        try:
            while True:
                yield
        except GeneratorExit:
            pass


def _checked_columns(columns):
    """None, or the requested columns as a non-empty tuple (names or positions)."""
    if columns is None:
        return None
    columns = tuple(columns)
    if not columns:
        raise ValueError("Columns argument must contain at least one item")
    return columns


class Tabular(FileHandler):
    """Rows as tuples of strings.  Reading can select (and reorder) columns by header name or by
    position, optionally followed by all the others; writing can start with a header row.
    Concrete formats supply the raw row reader / writer (handlers.py:106-207)."""

    def _iter_read(self, columns=None, has_headers: bool = False, get_all_columns: bool = False) -> ReadHandle[Row]:
        wanted = _checked_columns(columns)
        by_name = wanted is not None and isinstance(wanted[0], str)
        self.has_headers = has_headers or by_name
        self.header_row = None
        self.column_order = None
        rows = self._iter_read_rows()

        if self.has_headers:
            self.header_row = next(rows, None)
            if self.header_row is None:          # an empty file: nothing to select from
                yield self
                return
        if wanted is not None:
            if by_name:
                absent = set(wanted) - set(self.header_row)
                if absent:
                    raise ValueError(f"Column header(s) not found in file: {absent}")
                wanted = tuple(self.header_row.index(name) for name in wanted)
            pending = []
            if get_all_columns:
                if self.has_headers:
                    width = len(self.header_row)
                else:                            # the first data row tells how wide the table is
                    pending = [next(rows)]
                    width = len(pending[0])
                wanted += tuple(set(range(width)) - set(wanted))
            self.column_order = wanted

        yield self
        if wanted is None:
            yield from rows
            return
        for row in pending:
            yield tuple(row[k] for k in wanted)
        for row in rows:
            yield tuple(row[k] for k in wanted)

    def _iter_write(self, columns=None) -> WriteHandle[Row]:
        header = _checked_columns(columns)
        sink = self._iter_write_rows()
        next(sink)
        if header is not None:
            sink.send(header)
        try:
            while True:
                sink.send((yield))
        except GeneratorExit:
            sink.close()

    @property
    def headers(self):
        """The header row in the order the rows are delivered; None for a file without headers."""
        assert self.readable()
        if not self.has_headers:
            return None
        if not self.column_order:
            return self.header_row
        return tuple(self.header_row[k] for k in self.column_order)

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
    """Tab-separated text.  Read as utf-8 (undecodable bytes survive as surrogates), one row per
    line, empty lines skipped; written with "\\n" line ends."""

    def _iter_read_rows(self) -> Iterator[Row]:
        with open(self.path, "r", encoding="utf-8", errors="surrogateescape") as file:
            for line in file:
                text = line[:-1]             # the reference drops the last character of every line, newline or not
                if text:
                    yield tuple(text.split("\t"))

    def _iter_write_rows(self) -> Generator[None, Row, None]:
        with open(self.path, "w") as file:
            try:
                while True:
                    file.write("\t".join((yield)) + "\n")
            except GeneratorExit:
                pass
