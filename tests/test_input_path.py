"""Input side of the path: plain tab files and partition files, against the
reference's own reader fixtures (tests/golden/{handlers,partitions}/, copied by make_golden.py from
/root/reference/tests/test_handlers/ and tests/test_partitions/).  Cases follow
/root/reference/tests/test_handlers.py:198-383 (Tabfile only) and tests/test_partitions.py:77-121
(Tabfile and Fasta; the Spart and Excel readers are out of scope)."""
from __future__ import annotations

import re
from pathlib import Path

import pytest

from taxi2_b200.handlers import FileHandler
from taxi2_b200.partitions import Classification, Partition, PartitionHandler

GOLDEN = Path(__file__).parent / "golden"

ROWS = [("item_1_1", "item_1_2", "item_1_3"), ("item_2_1", "item_2_2", "item_2_3"), ("item_3_1", "item_3_2", "item_3_3")]
HEADERS = ("header_1", "header_2", "header_3")


def pick(row, order):
    return tuple(row[k] for k in order)


# (file, kwargs, expected headers, expected rows)
READ_CASES = [
    ("simple.tsv", {}, None, ROWS),
    ("simple.tsv", dict(columns=[0, 2]), None, [pick(r, (0, 2)) for r in ROWS]),
    ("simple.tsv", dict(columns=[0, 2], get_all_columns=True), None, [pick(r, (0, 2, 1)) for r in ROWS]),
    ("headers.tsv", dict(has_headers=True), HEADERS, ROWS),
    ("headers.tsv", dict(columns=[0, 2], has_headers=True), pick(HEADERS, (0, 2)), [pick(r, (0, 2)) for r in ROWS]),
    ("headers.tsv", dict(columns=[0, 2], has_headers=True, get_all_columns=True), pick(HEADERS, (0, 2, 1)), [pick(r, (0, 2, 1)) for r in ROWS]),
    ("headers.tsv", dict(columns=["header_1", "header_3"]), pick(HEADERS, (0, 2)), [pick(r, (0, 2)) for r in ROWS]),
    ("headers.tsv", dict(columns=["header_1", "header_3"], get_all_columns=True), pick(HEADERS, (0, 2, 1)), [pick(r, (0, 2, 1)) for r in ROWS]),
    ("skip.tsv", {}, None, ROWS),                                   # blank lines are skipped
    ("empty.tsv", {}, None, []),
    ("empty.tsv", dict(has_headers=True), None, []),
    ("empty.tsv", dict(columns=[0, 2]), None, []),
    ("empty.tsv", dict(columns=["header_1", "header_3"]), None, []),
]


@pytest.mark.parametrize("name,kwargs,headers,rows", READ_CASES)
def test_read_tabfile(name, kwargs, headers, rows):
    path = GOLDEN / "handlers" / name
    with FileHandler.Tabfile(path, **kwargs) as file:                # iterate inside a context
        assert not file.closed
        assert file.headers == headers
        assert list(file) == rows
    assert file.closed
    with FileHandler.Tabfile(path, **kwargs) as file:                # read() until exhausted
        got = []
        while (item := file.read()) is not None:
            got.append(item)
        assert got == rows
    file = FileHandler.Tabfile(path, **kwargs)                        # plain open / close
    assert [item for item, _ in zip(file, rows)] == rows
    file.close()
    assert file.closed


def test_read_tabfile_errors_and_early_close():
    with pytest.raises(ValueError):
        FileHandler.Tabfile(GOLDEN / "handlers" / "headers.tsv", columns=["header_X"])
    with pytest.raises(ValueError):
        FileHandler.Tabfile(GOLDEN / "handlers" / "headers.tsv", columns=[])
    file = FileHandler.Tabfile(GOLDEN / "handlers" / "simple.tsv")
    file.read()
    assert not file.closed
    file.close()
    assert file.closed
    assert FileHandler.Tabfile.get_headers(GOLDEN / "handlers" / "headers.tsv") == HEADERS


def test_a_reader_must_yield_itself_first():
    class Bad(FileHandler):
        def _iter_read(self):
            yield 42
            yield self

        def _iter_write(self):
            raise NotImplementedError()

    with pytest.raises(Exception):
        Bad(Path(), "r")


@pytest.mark.parametrize("name,kwargs,rows", [
    ("simple.tsv", {}, ROWS),
    ("headers.tsv", dict(columns=list(HEADERS)), ROWS),
])
def test_write_tabfile(tmp_path, name, kwargs, rows):
    out = tmp_path / name
    with FileHandler.Tabfile(out, "w", **kwargs) as file:
        for row in rows:
            file.write(row)
    strip = lambda p: re.sub(r"\s", "", p.read_text())  # noqa: E731  (the reference's assert_eq_files, tests/utility.py)
    assert strip(out) == strip(GOLDEN / "handlers" / name)
    assert out.read_bytes() == (GOLDEN / "handlers" / name).read_bytes()   # and in fact byte for byte


SIMPLE = {"sample1": "speciesA", "sample2": "speciesA", "sample3": "speciesA", "sample4": "speciesA",
          "sample5": "speciesB", "sample6": "speciesB", "sample7": "speciesC"}
MISSING = {"sample3": "speciesA", "sample4": "speciesA", "sample6": "speciesB", "sample7": "speciesC"}
GENERA = {"sample1": "genusX", "sample2": "genusX", "sample3": "genusX", "sample4": "genusX",
          "sample5": "genusY", "sample6": "genusY", "sample7": "genusY"}


@pytest.mark.parametrize("want,name,handler,kwargs", [
    (SIMPLE, "simple.tsv", PartitionHandler.Tabfile, {}),
    (SIMPLE, "extras.tsv", PartitionHandler.Tabfile, dict(idHeader="seqid", subHeader="organism")),
    (GENERA, "genera.tsv", PartitionHandler.Tabfile, dict(filter=PartitionHandler.subset_first_word, idHeader="seqid", subHeader="organism")),
    (SIMPLE, "simple.fas", PartitionHandler.Fasta, {}),
    (SIMPLE, "simple.dot.fas", PartitionHandler.Fasta, dict(separator=".")),
    (MISSING, "missing.fas", PartitionHandler.Fasta, {}),
    (GENERA, "genera.fas", PartitionHandler.Fasta, dict(filter=PartitionHandler.subset_first_word)),
    (SIMPLE, "genera.fas", PartitionHandler.Fasta, dict(filter=lambda x: Classification(x.individual, x.subset.split(" ")[1]))),
])
def test_read_partitions(want, name, handler, kwargs):
    got = Partition.fromPath(GOLDEN / "partitions" / name, handler, **kwargs)
    assert got == want and list(got) == list(want)      # same mapping, same (file) order


def test_fasta_partition_sniffing():
    fasta = PartitionHandler.Fasta
    assert fasta.guess_subset_separator(GOLDEN / "partitions" / "simple.fas") == "|"
    assert fasta.guess_subset_separator(GOLDEN / "partitions" / "simple.dot.fas") == "."
    assert fasta.has_subsets(GOLDEN / "partitions" / "simple.fas") is True
    assert fasta.has_subsets(GOLDEN / "partitions" / "missing.fas") is False     # decided by the first record
    assert fasta.has_subsets(GOLDEN / "partitions" / "simple.fas", "") is False
