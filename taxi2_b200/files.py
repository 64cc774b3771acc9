"""File format tags used by the dereplicate / decontaminate tasks to pick their output writer.

Mirrors the Fasta / Tabfile part of /root/reference/src/itaxotools/taxi2/file_types.py:10-25 and
files.py:24-28,54-85; the other sniffers (Ali, FastQ, Spart, Excel, Newick) are out of scope.
"""
from __future__ import annotations

from enum import Enum
from pathlib import Path
from re import fullmatch


class FileFormat(Enum):
    Fasta = "Fasta", ".fas"
    Tabfile = "Tabfile", ".tsv"
    Unknown = "Unknown", None

    def __init__(self, label, extension):
        self.label = label
        self.extension = extension

    def __repr__(self):
        return f"<{type(self).__name__}.{self._name_}>"


def is_fasta(path: Path) -> bool:
    with Path(path).open() as file:
        for line in file:
            if not line.strip() or line.startswith(";"):
                continue
            if line.startswith(">"):
                return True
    return False


def is_tabfile(path: Path) -> bool:
    with Path(path).open() as file:
        return bool(fullmatch(r"([^\t]+\t)+[^\t]+", file.readline()))


def identify_format(path: Path) -> FileFormat:
    if is_fasta(path):
        return FileFormat.Fasta
    if is_tabfile(path):
        return FileFormat.Tabfile
    return FileFormat.Unknown
