"""Sequence records and the two sequence file formats on either side of the hot path.

Mirrors /root/reference/src/itaxotools/taxi2/sequences.py: Sequence / Sequences (:15-40), the
Fasta handler (:47-134) and the Tabular / Tabfile handlers (:172-234).  Ali, FastQ, Genbank and
Excel readers are out of scope (format zoo unrelated to the compute path).
"""
from __future__ import annotations

from pathlib import Path
from typing import Literal, NamedTuple

from .encoding import sanitize
from .handlers import FileHandler, ReadHandle, WriteHandle
from .types import Container


class Sequence(NamedTuple):
    id: str
    seq: str
    extras: dict = dict()

    _tr_normalize = str.maketrans("?", "N", "-")

    def normalize(self) -> "Sequence":
        """'?' -> 'N', gaps removed, upper case: the input contract of the aligner (sequences.py:20-25)."""
        return Sequence(self.id, self.seq.translate(self._tr_normalize).upper(), self.extras)

    def get_sanitized_id_with_extras(self) -> str:
        return sanitize("_".join([self.id] + list(self.extras.values())))


class Sequences(Container[Sequence]):
    @classmethod
    def fromPath(cls, path: Path, handler: "SequenceHandler", *args, **kwargs) -> "Sequences":
        return cls(handler, path, "r", *args, **kwargs)

    def normalize(self) -> "Sequences":
        return Sequences(lambda: (seq.normalize() for seq in self))


class SequenceHandler(FileHandler[Sequence]):
    pass


def _fasta_records(handle):
    """(title, sequence) per record; title = header line without '>', sequence lines joined and
    stripped of whitespace -- what Bio.SeqIO.FastaIO.SimpleFastaParser yields."""
    title, chunks = None, []
    for line in handle:
        if line.startswith(">"):
            if title is not None:
                yield title, "".join(chunks).replace(" ", "").replace("\r", "")
            title, chunks = line[1:].rstrip(), []
        elif title is not None:
            chunks.append(line.strip())
    if title is not None:
        yield title, "".join(chunks).replace(" ", "").replace("\r", "")


class Fasta(SequenceHandler):
    def _open(self, path: Path, mode: Literal["r", "w"] = "r", organism_separator="|", organism_tag="organism",
              *args, **kwargs):
        self.organism_separator = organism_separator
        self.organism_tag = organism_tag
        super()._open(path, mode, *args, **kwargs)

    def _iter_read(self, parse_organism: bool = False) -> ReadHandle[Sequence]:
        with open(self.path, "r") as handle:
            yield self
            for title, sequence in _fasta_records(handle):
                if not parse_organism:
                    yield Sequence(title, sequence)
                    continue
                id, sep, organism = title.partition(self.organism_separator)
                yield Sequence(id, sequence, extras={self.organism_tag: organism if sep else None})

    def _iter_write(self, write_organism: bool = False, concatenate_extras: list = [], line_width=60) -> WriteHandle[Sequence]:
        self.concatenate_extras = concatenate_extras
        with open(self.path, "w") as handle:
            try:
                while True:
                    sequence = yield
                    identifier = "_".join((sequence.id, *(sequence.extras[tag] for tag in concatenate_extras)))
                    if write_organism:
                        organism = sequence.extras.get(self.organism_tag, None)
                        if organism:
                            identifier += self.organism_separator + organism
                    handle.write(">" + identifier + "\n")
                    if line_width:
                        for k in range(0, len(sequence.seq), line_width):
                            handle.write(sequence.seq[k:k + line_width] + "\n")
                        handle.write("\n")
                    else:
                        handle.write(sequence.seq + "\n")
            except GeneratorExit:
                return


class Tabular(SequenceHandler):
    subhandler = FileHandler.Tabular

    def _iter_read(self, idHeader: str = None, seqHeader: str = None, hasHeader: bool = False,
                   idColumn: int = 0, seqColumn: int = 1) -> ReadHandle[Sequence]:
        if idHeader and seqHeader:
            columns, hasHeader = (idHeader, seqHeader), True
        else:
            columns = (idColumn, seqColumn)
        with self.subhandler(self.path, has_headers=hasHeader, columns=columns, get_all_columns=True) as rows:
            headers = rows.headers
            if headers is not None:
                headers = [sanitize(header) for header in headers]
            yield self
            for row in rows:
                extras = dict(zip(headers[2:], row[2:])) if headers is not None else dict()
                yield Sequence(row[0], row[1], extras)

    def _iter_write(self, *args, **kwargs) -> WriteHandle[Sequence]:
        raise NotImplementedError()


class Tabfile(SequenceHandler.Tabular, SequenceHandler):
    subhandler = FileHandler.Tabular.Tabfile

    def _iter_write(self, idHeader: str = None, seqHeader: str = None, hasHeader: bool = False) -> WriteHandle[Sequence]:
        if idHeader and seqHeader:
            hasHeader = True
        wrote_headers = False
        with self.subhandler(self.path, "w") as file:
            try:
                sequence = yield
                if hasHeader:
                    file.write((idHeader, *sequence.extras.keys(), seqHeader))
                    wrote_headers = True
                while True:
                    file.write((sequence.id, *sequence.extras.values(), sequence.seq))
                    sequence = yield
            except GeneratorExit:
                if hasHeader and not wrote_headers:
                    file.write((idHeader, seqHeader))
