"""Task-level parity on the GPU: every output file of VersusAll / VersusReference byte-for-byte
against the per-pair restatement of the reference pipeline (tests/ref_pipeline.py, oracle-backed).
The reference itself pins no task output (its only task test is an xfail stub), so this is the
strongest check available: same order, same None/NA quirks, same aggregates, same formatting."""
from __future__ import annotations

from pathlib import Path

import pytest

import ref_pipeline
from conftest import GOLDEN
from taxi2_b200.distances import DistanceMetric
from taxi2_b200.partitions import Partition, PartitionHandler
from taxi2_b200.sequences import Sequence, SequenceHandler, Sequences
from taxi2_b200.tasks import VersusAll, VersusReference

pytestmark = pytest.mark.gpu

SILENT = lambda *a: None  # noqa: E731


def load(name):
    path = GOLDEN / name
    seqs = Sequences.fromPath(path, SequenceHandler.Tabfile, idHeader="seqid", seqHeader="sequence")
    species = Partition.fromPath(path, PartitionHandler.Tabfile, idHeader="seqid", subHeader="organism")
    genera = Partition.fromPath(path, PartitionHandler.Tabfile, idHeader="seqid", subHeader="organism",
                                filter=PartitionHandler.subset_first_word)
    return seqs, species, genera


def assert_same_tree(got: Path, want: Path):
    files_got = sorted(p.relative_to(got) for p in got.rglob("*") if p.is_file())
    files_want = sorted(p.relative_to(want) for p in want.rglob("*") if p.is_file())
    assert files_got == files_want
    for rel in files_want:
        assert (got / rel).read_bytes() == (want / rel).read_bytes(), rel


@pytest.mark.parametrize("align,write,multiply", [(True, True, False), (True, False, True), (False, False, False)])
def test_versus_all_outputs(tmp_path, align, write, multiply):
    seqs, species, genera = load("Taxi2test1_10.tab")
    # a duplicated record and a pair of different records with equal sequences exercise the x != y quirk
    records = list(seqs)
    records.append(records[3])
    records.append(Sequence("twin", records[5].seq, records[5].extras))
    task = VersusAll()
    task.work_dir = tmp_path / "got"
    task.progress_handler = SILENT
    task.input.sequences = Sequences(records)
    task.input.species, task.input.genera = species, genera
    task.params.pairs.align, task.params.pairs.write = align, write
    task.params.format.percentage_multiply = multiply
    results = task.start()
    assert results.output_directory == task.work_dir and results.seconds_taken > 0
    want = tmp_path / "want"
    ref_pipeline.versus_all(records, want, species, genera, align=align, multiply=multiply)
    if align and not write:
        (want / "align" / "aligned_pairs.txt").unlink()
        (want / "align").rmdir()
    assert_same_tree(task.work_dir, want)


def test_versus_all_without_partitions_and_metric_subset(tmp_path):
    seqs, _, _ = load("Taxi2test1_10.tab")
    task = VersusAll()
    task.work_dir = tmp_path / "got"
    task.progress_handler = SILENT
    task.input.sequences = seqs
    task.params.distances.metrics = [DistanceMetric.Kimura2P(), DistanceMetric.Uncorrected()]
    task.start()
    ref_pipeline.versus_all(list(seqs), tmp_path / "want", metrics=[DistanceMetric.Kimura2P(), DistanceMetric.Uncorrected()])
    assert_same_tree(task.work_dir, tmp_path / "want")


@pytest.mark.parametrize("align,multiply,main", [(True, False, "p"), (True, True, "k2p"), (False, False, "p-gaps")])
def test_versus_reference_outputs(tmp_path, align, multiply, main):
    seqs, _, _ = load("Taxi2test1_50.tab")
    records = list(seqs)
    data, reference = records[:12], records[12:40]
    task = VersusReference()
    task.work_dir = tmp_path / "got"
    task.progress_handler = SILENT
    task.input.data, task.input.reference = Sequences(data), Sequences(reference)
    task.params.pairs.align = align
    task.params.distances.metric = DistanceMetric.fromLabel(main)
    task.params.format.percentage_multiply = multiply
    task.start()
    ref_pipeline.versus_reference(data, reference, tmp_path / "want", align=align, metric=DistanceMetric.fromLabel(main), multiply=multiply)
    assert_same_tree(task.work_dir, tmp_path / "want")


def test_versus_reference_raises_without_any_defined_distance(tmp_path):
    task = VersusReference()
    task.work_dir = tmp_path / "got"
    task.progress_handler = SILENT
    task.input.data = Sequences([Sequence("q", "NNNN")])
    task.input.reference = Sequences([Sequence("r", "ACGT")])
    with pytest.raises(ValueError):
        task.start()
