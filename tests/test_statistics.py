"""Sequence-set statistics against the reference's own vectors and writer fixtures
(tests/golden/statistics_cases.json and tests/golden/statistics/*, extracted by make_golden.py from
/root/reference/tests/test_statistics.py:113-250 and tests/test_statistics/*)."""
from __future__ import annotations

import json
import re
from pathlib import Path

import pytest

from taxi2_b200.statistics import Counts, Statistic, Statistics, StatisticsCalculator, StatisticsHandler

GOLDEN = Path(__file__).parent / "golden"
CASES = json.loads((GOLDEN / "statistics_cases.json").read_text())


def same_modulo_whitespace(a: Path, b: Path) -> bool:
    """the reference's `assert_eq_files` ignores every whitespace character (tests/utility.py:5-19)"""
    strip = lambda p: re.sub(r"\s", "", p.read_text())  # noqa: E731
    return strip(a) == strip(b)


@pytest.mark.parametrize("case", CASES["counts"], ids=lambda c: f"{c['count']}-{c['sequence'] or 'empty'}")
def test_counts(case):
    assert getattr(Counts.from_sequence(case["sequence"]), case["count"]) == case["fixed"]


@pytest.mark.parametrize("k", range(len(CASES["statistics"])))
def test_statistics(k):
    case = CASES["statistics"][k]
    got = Statistics.from_sequences(case["sequences"])[Statistic[case["stat"]]]
    if isinstance(got, float):
        assert abs(got - case["fixed"]) <= 0.00051       # the reference's own tolerance
    else:
        assert got == case["fixed"]


def test_calculator_is_single_use():
    calc = StatisticsCalculator()
    calc.calculate()
    with pytest.raises(StopIteration):
        calc.add("ACTG")
    with pytest.raises(StopIteration):
        calc.calculate()


def tiny(single: bool):
    a = {Statistic.SequenceCount: 1, Statistic.NucleotideCount: 42}
    b = {Statistic.SequenceCount: 2, Statistic.NucleotideCount: 43}
    if single:
        return [Statistics(a)]
    return [Statistics({Statistic.Group: "A", **a}), Statistics({Statistic.Group: "B", **b})]


def simple(single: bool):
    if single:
        return [Statistics.from_sequences(["ACTG"])]
    return [Statistics.from_sequences(["ACTG"], "A"), Statistics.from_sequences(["AC", "TG"], "B")]


@pytest.mark.parametrize("fixture,output,handler,kwargs", [
    (tiny, "tiny.single", "Single", {}),
    (simple, "simple.single", "Single", dict(float_formatter="{:.1f}", percentage_formatter="{:.4f}")),
    (simple, "percent.single", "Single", dict(float_formatter="{:.1f}", percentage_formatter="{}")),
    (tiny, "tiny.groups", "Groups", {}),
    (simple, "simple.groups", "Groups", dict(group_name="genus", float_formatter="{:.1f}", percentage_formatter="{:.4f}")),
    (simple, "percent.groups", "Groups", dict(group_name="genus", float_formatter="{:.1f}", percentage_formatter="{}")),
])
def test_writers_match_the_reference_fixtures(tmp_path, fixture, output, handler, kwargs):
    out = tmp_path / output
    with getattr(StatisticsHandler, handler)(out, "w", **kwargs) as file:
        for stats in fixture(handler == "Single"):
            file.write(stats)
    assert same_modulo_whitespace(out, GOLDEN / "statistics" / output)


def test_single_rejects_a_second_record(tmp_path):
    with StatisticsHandler.Single(tmp_path / "bad.single") as file:
        file.write(Statistics.from_sequences("ACGT"))
        with pytest.raises(Exception, match="single"):
            file.write(Statistics.from_sequences("ACGT"))


def test_groups_need_a_group_name(tmp_path):
    with StatisticsHandler.Groups(tmp_path / "bad.groups") as file:
        with pytest.raises(Exception, match="name"):
            file.write(Statistics.from_sequences("ACGT"))


def test_percentage_multiply_and_order(tmp_path):
    stats = Statistics.from_sequences(["AACG--", "NNTT"], "g")
    assert list(stats)[0] is Statistic.Group and list(stats)[1] is Statistic.SequenceCount
    out = tmp_path / "m.groups"
    with StatisticsHandler.Groups(out, "w", percentage_formatter="{:.2f}", percentage_multiply=True) as file:
        file.write(stats)
    header, row = [line.split("\t") for line in out.read_text().splitlines()]
    col = header.index(str(Statistic.PercentA))
    assert row[col] == "25.00"                       # 2 A of 8 nucleotides, x100
    assert row[header.index(str(Statistic.PercentGaps))] == "20.00"   # 2 gaps of 10 characters


@pytest.mark.parametrize("align,multiply", [(True, False), (False, True)])
def test_versus_all_statistics_files(tmp_path, align, multiply):
    """stats/{all,species,genera}.tsv of the VersusAll task (host only, no GPU) against a numpy
    restatement of the definitions (tests/ref_pipeline.py)."""
    import ref_pipeline
    from taxi2_b200.partitions import Partition, PartitionHandler
    from taxi2_b200.sequences import Sequence, SequenceHandler, Sequences
    from taxi2_b200.tasks import VersusAll

    path = GOLDEN / "Taxi2test1_50.tab"
    records = list(Sequences.fromPath(path, SequenceHandler.Tabfile, idHeader="seqid", seqHeader="sequence"))
    records.append(Sequence("gappy", "ac-gt--nnryk" * 9, records[0].extras))       # gaps, missing and ambiguity codes
    records.append(Sequence("unplaced", "a" * 1200, {}))                            # not in any partition; > 1000 bp
    species = Partition.fromPath(path, PartitionHandler.Tabfile, idHeader="seqid", subHeader="organism")
    genera = Partition.fromPath(path, PartitionHandler.Tabfile, idHeader="seqid", subHeader="organism",
                                filter=PartitionHandler.subset_first_word)
    task = VersusAll()
    task.work_dir = tmp_path / "got"
    task.input.sequences = Sequences(records)
    task.input.species, task.input.genera = species, genera
    task.params.format.percentage_multiply = multiply
    task.generate_paths()
    seqs = [s.normalize() for s in records] if align else records
    task.write_statistics(seqs)
    ref_pipeline.write_stats(seqs, tmp_path / "want", species, genera, "{:.4f}", multiply=multiply)
    for name in ("all", "species", "genera"):
        got, want = tmp_path / "got" / "stats" / f"{name}.tsv", tmp_path / "want" / "stats" / f"{name}.tsv"
        assert got.read_bytes() == want.read_bytes(), name
    assert len((tmp_path / "got" / "stats" / "genera.tsv").read_text().splitlines()) == 1 + len(set(genera.values()))
