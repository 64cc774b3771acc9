"""Dereplicate's block replay of the greedy walk (host logic) against the per-pair restatement of
the reference pipeline (tests/ref_pipeline.py), on CPU: distances come from the oracle through a
stand-in engine, so what is under test is the replay itself -- which pairs are visited, in what
order exclusions happen, when the survivors are re-loaded, and every output file."""
from __future__ import annotations

import numpy as np
import pytest

import ref_pipeline
from fake_engine import OracleEngine, oracle_multi
from taxi2_b200.files import FileFormat
from taxi2_b200.sequences import Sequence, Sequences
from taxi2_b200.tasks import Dereplicate, common, dereplicate as dereplicate_module


def tree(path):
    return {str(p.relative_to(path)): p.read_bytes() for p in sorted(path.rglob("*")) if p.is_file()}


def clusters(rng, nclusters, per, length, sub=0.02):
    """Clusters of near-identical sequences of varying length (prefixes), shuffled: plenty of
    similar pairs, longer and shorter partners on both sides of every query."""
    al = np.frombuffer(b"acgt", dtype=np.uint8)
    records = []
    for c in range(nclusters):
        root = al[rng.integers(0, 4, length)]
        for k in range(per):
            s = root.copy()
            hit = rng.random(length) < sub
            s[hit] = al[rng.integers(0, 4, int(hit.sum()))]
            cut = int(rng.integers(length // 2, length + 1))
            text = s[:cut].tobytes().decode()
            if rng.random() < 0.2:
                text = text[: cut // 2] + "-" + text[cut // 2:]          # raw length counts the gap
            records.append(Sequence(f"c{c}_{k}", text, {"organism": f"G{c} s{k % 3}"}))
    order = rng.permutation(len(records))
    return [records[k] for k in order]


@pytest.fixture
def cpu_engines(monkeypatch):
    multi = oracle_multi(2)
    monkeypatch.setattr(dereplicate_module, "task_engine", lambda task: multi)
    import taxi2_b200.engine as engine_module
    monkeypatch.setattr(engine_module, "Engine", OracleEngine)
    return multi


@pytest.mark.parametrize("seed,align,write,multiply,rows,first", [
    (1, True, True, False, 3, 4), (2, True, False, True, 5, 2), (3, False, False, False, 2, 7), (4, True, False, False, None, None),
    (5, False, True, False, 1, 1), (6, True, True, False, 4, 5), (7, False, False, True, 6, 3)])
def test_block_replay_matches_the_per_pair_pipeline(tmp_path, cpu_engines, seed, align, write, multiply, rows, first):
    rng = np.random.default_rng(seed)
    records = clusters(rng, nclusters=5, per=6, length=60)
    records.append(Sequence("tiny", "acg", {"organism": "x y"}))                # dropped by the length threshold
    task = Dereplicate()
    task.work_dir = tmp_path / "got"
    task.progress_handler = lambda *a: None
    task.input = Sequences(records)
    task.output_format = FileFormat.Tabfile
    task.rows_per_block = rows
    task.first_columns = first     # columns aligned before a row asks for more: small values force the lazy extension
    task.params.pairs.align, task.params.pairs.write = align, write
    task.params.distances.write_linear = task.params.distances.write_matricial = write or seed == 2
    task.params.format.percentage_multiply = multiply
    task.params.thresholds.similarity = 7.0 if multiply else 0.07
    task.start()
    excluded = ref_pipeline.dereplicate(records, tmp_path / "want", similarity=task.params.thresholds.similarity, align=align, multiply=multiply)
    want = tree(tmp_path / "want")
    if not (align and write):
        want.pop("aligned_pairs.txt", None)
    if not task.params.distances.write_linear:
        want = {k: v for k, v in want.items() if not k.startswith("distances/")}
    assert task.excluded == excluded and 10 <= len(excluded) < len(records)
    got = tree(task.work_dir)
    assert got.keys() == want.keys()
    for rel in want:
        assert got[rel] == want[rel], rel
    # the survivors were re-loaded at least once, and far fewer pairs were visited than n^2
    assert task.stats["reloads"] >= (2 if rows else 1)
    assert task.stats["pairs_visited"] < len(records) ** 2 / 2
    if first:   # the lazy evaluation computed less than the full rows of every walked block
        assert task.stats["pairs_computed"] < task.stats["reloads"] * len(records) ** 2


def test_repeated_ids_take_the_per_pair_path(tmp_path, cpu_engines):
    rng = np.random.default_rng(9)
    records = clusters(rng, nclusters=3, per=4, length=40)
    records.insert(4, Sequence(records[3].id, records[3].seq, records[3].extras))   # neighbouring rows of equal id
    task = Dereplicate()
    task.work_dir = tmp_path / "got"
    task.progress_handler = lambda *a: None
    task.input = Sequences(records)
    task.output_format = FileFormat.Tabfile
    task.start()
    excluded = ref_pipeline.dereplicate(records, tmp_path / "want")
    assert task.excluded == excluded
    got, want = tree(task.work_dir), tree(tmp_path / "want")
    assert got == want


def test_block_replay_on_a_larger_set_with_lazy_columns(tmp_path, cpu_engines):
    rng = np.random.default_rng(42)
    records = clusters(rng, nclusters=8, per=12, length=50, sub=0.03)
    task = Dereplicate()
    task.work_dir = tmp_path / "got"
    task.progress_handler = lambda *a: None
    task.input = Sequences(records)
    task.output_format = FileFormat.Fasta
    task.rows_per_block = 5
    task.first_columns = 6
    task.params.pairs.write = False
    task.start()
    excluded = ref_pipeline.dereplicate(records, tmp_path / "want", fasta=True)
    (tmp_path / "want" / "aligned_pairs.txt").unlink()
    assert task.excluded == excluded
    assert tree(task.work_dir) == tree(tmp_path / "want")
    assert task.stats["pairs_computed"] < 0.8 * len(records) ** 2
