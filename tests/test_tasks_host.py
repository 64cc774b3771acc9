"""Task host logic on CPU: the block iterator over a MultiEngine (row tiles dealt to several
engines, blocks consumed in reference order while the engines run ahead) under VersusAll and
VersusReference, with the oracle standing in for the device.  Output trees must equal the per-pair
restatement of the reference pipelines byte for byte, for any number of engines and block sizes."""
from __future__ import annotations

import pytest

import ref_pipeline
from conftest import GOLDEN
from fake_engine import oracle_multi
from taxi2_b200.partitions import Partition, PartitionHandler
from taxi2_b200.sequences import Sequence, SequenceHandler, Sequences
from taxi2_b200.tasks import VersusAll, VersusReference, common, versus_all, versus_reference


def tree(path):
    return {str(p.relative_to(path)): p.read_bytes() for p in sorted(path.rglob("*")) if p.is_file()}


def load(name):
    path = GOLDEN / name
    seqs = Sequences.fromPath(path, SequenceHandler.Tabfile, idHeader="seqid", seqHeader="sequence")
    species = Partition.fromPath(path, PartitionHandler.Tabfile, idHeader="seqid", subHeader="organism")
    genera = Partition.fromPath(path, PartitionHandler.Tabfile, idHeader="seqid", subHeader="organism",
                                filter=PartitionHandler.subset_first_word)
    return seqs, species, genera


def short(records, n=90):
    """The sample's records cut to their first n bases: the oracle aligns them in milliseconds."""
    return [Sequence(s.id, s.seq[:n], s.extras) for s in records]


@pytest.mark.parametrize("engines,block_pairs,align,write", [(1, 40, True, True), (3, 25, True, True), (2, 13, True, False), (4, 30, False, False)])
def test_versus_all_blocks_over_several_engines(tmp_path, monkeypatch, engines, block_pairs, align, write):
    seqs, species, genera = load("Taxi2test1_10.tab")
    records = short(list(seqs))
    records.append(records[3])                                    # the x != y quirk: an identical record
    multi = oracle_multi(engines)
    monkeypatch.setattr(versus_all, "task_engine", lambda task: multi)
    monkeypatch.setattr(common, "MAX_BLOCK_PAIRS", block_pairs)
    task = VersusAll()
    task.work_dir = tmp_path / "got"
    task.progress_handler = lambda *a: None
    task.input.sequences = Sequences(records)
    task.input.species, task.input.genera = species, genera
    task.params.pairs.align, task.params.pairs.write = align, write
    task.start()
    ref_pipeline.versus_all(records, tmp_path / "want", species, genera, align=align)
    want = tree(tmp_path / "want")
    if align and not write:
        want.pop("align/aligned_pairs.txt")
    assert tree(task.work_dir) == want
    assert sum(e.calls > 0 for e in multi.engines) == min(engines, -(-len(records) // max(1, block_pairs // len(records))))


@pytest.mark.parametrize("engines,block_pairs", [(1, 50), (3, 20)])
def test_versus_reference_blocks_over_several_engines(tmp_path, monkeypatch, engines, block_pairs):
    seqs, _, _ = load("Taxi2test1_50.tab")
    records = short(list(seqs))
    data, reference = records[:9], records[9:25]
    multi = oracle_multi(engines)
    monkeypatch.setattr(versus_reference, "task_engine", lambda task: multi)
    monkeypatch.setattr(common, "MAX_BLOCK_PAIRS", block_pairs)
    task = VersusReference()
    task.work_dir = tmp_path / "got"
    task.progress_handler = lambda *a: None
    task.input.data, task.input.reference = Sequences(data), Sequences(reference)
    task.start()
    ref_pipeline.versus_reference(data, reference, tmp_path / "want")
    assert tree(task.work_dir) == tree(tmp_path / "want")


@pytest.mark.parametrize("align,multiply,fasta,native", [(True, False, False, True), (True, True, True, False), (False, False, True, True)])
def test_decontaminate_blocks_over_several_engines(tmp_path, monkeypatch, align, multiply, fasta, native):
    from taxi2_b200.files import FileFormat
    from taxi2_b200.tasks import Decontaminate, Decontaminate2, decontaminate

    seqs, _, _ = load("Taxi2test1_50.tab")
    records = short(list(seqs))
    data, outgroup, ingroup = records[:10], records[10:22], records[22:34]
    data.append(Sequence("allN", "nnnnnnnn", records[0].extras))
    multi = oracle_multi(3)
    monkeypatch.setattr(decontaminate, "task_engine", lambda task: multi)
    monkeypatch.setattr(common, "MAX_BLOCK_PAIRS", 30)
    task = Decontaminate()
    task.work_dir = tmp_path / "got1"
    task.native_writers = native
    task.progress_handler = lambda *a: None
    task.input, task.outgroup = Sequences(data), Sequences(outgroup)
    task.output_format = FileFormat.Fasta if fasta else FileFormat.Tabfile
    task.params.pairs.align = align
    task.params.format.percentage_multiply = multiply
    task.params.thresholds.similarity = 12.0 if multiply else 0.12
    task.start()
    ref_pipeline.decontaminate(data, outgroup, tmp_path / "want1", similarity=task.params.thresholds.similarity, align=align,
                               multiply=multiply, fasta=fasta)
    assert tree(task.work_dir) == tree(tmp_path / "want1")

    task2 = Decontaminate2()
    task2.work_dir = tmp_path / "got2"
    task2.native_writers = native
    task2.progress_handler = lambda *a: None
    task2.input, task2.outgroup, task2.ingroup = Sequences(data), Sequences(outgroup), Sequences(ingroup)
    task2.output_format = FileFormat.Fasta if fasta else FileFormat.Tabfile
    task2.params.pairs.align = align
    task2.params.format.percentage_multiply = multiply
    task2.params.weights.outgroup, task2.params.weights.ingroup = 1.0, 1.5
    task2.start()
    ref_pipeline.decontaminate2(data, outgroup, ingroup, tmp_path / "want2", w_out=1.0, w_in=1.5, align=align, multiply=multiply, fasta=fasta)
    assert tree(task2.work_dir) == tree(tmp_path / "want2")
