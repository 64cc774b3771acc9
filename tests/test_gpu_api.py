"""The reference-facing Python API on the GPU; written to read like the reference's own tests
(tests/test_align.py AlignTest.check, tests/test_distances.py MetricTest.check)."""
from __future__ import annotations

import json

import pytest

import oracle
from conftest import GOLDEN
from taxi2_b200.align import PairwiseAligner, Scores
from taxi2_b200.distances import DistanceMetric
from taxi2_b200.pairs import SequencePair, SequencePairs
from taxi2_b200.sequences import Sequence, SequenceHandler, Sequences

pytestmark = pytest.mark.gpu

ALIGN = json.loads((GOLDEN / "align_cases.json").read_text())
METRICS = json.loads((GOLDEN / "metrics_cases.json").read_text())


def scores_from_tuple(scores):
    return Scores(**{k: v for k, v in zip(Scores.defaults, scores)})


@pytest.mark.parametrize("case", ALIGN["align_tests"] + ALIGN["align_tests_failing"], ids=lambda c: f"{c['x']}-{c['y']}-{c['scores']}")
def test_align(case):
    aligner = PairwiseAligner.Biopython(scores_from_tuple(case["scores"]))
    x, y = Sequence("idx", case["x"]), Sequence("idy", case["y"])
    ax, ay = aligner.align(SequencePair(x, y))
    assert ax.id == x.id and ay.id == y.id
    assert len(ax.seq) == len(ay.seq)
    assert any(solution == [ax.seq, ay.seq] for solution in case["solutions"])


def test_align_pairs_is_lazy_ordered_and_reiterable():
    seqs = Sequences.fromPath(GOLDEN / "Taxi2test1_10.tab", SequenceHandler.Tabfile, idHeader="seqid", seqHeader="sequence").normalize()
    pairs = SequencePairs.fromProduct(seqs, seqs)
    aligner = PairwiseAligner.Biopython()
    aligner.batch_size = 32  # force several device batches
    aligned = aligner.align_pairs(pairs)
    first, second = list(aligned), list(aligned)
    assert first == second and len(first) == 100
    for pair, got in zip(pairs, first):
        assert (got.x.id, got.y.id) == (pair.x.id, pair.y.id)
        assert got.x.extras == pair.x.extras
        ox, oy, _ = oracle.align(pair.x.seq, pair.y.seq)
        assert (got.x.seq, got.y.seq) == (ox, oy)


def test_zero_length_raises_value_error():
    with pytest.raises(ValueError):
        PairwiseAligner.Biopython().align(SequencePair(Sequence("a", ""), Sequence("b", "ACGT")))


def test_non_integer_scores_are_rejected():
    with pytest.raises(ValueError):
        PairwiseAligner.Biopython(Scores(match_score=0.5))


@pytest.mark.parametrize("row", METRICS["rows"], ids=lambda r: f"{r['x']}-{r['y']}")
def test_metrics_from_files(row):
    for label, expected in zip(METRICS["labels"], row["expected"]):
        metric = DistanceMetric.fromLabel(label)
        r = metric.calculate(Sequence("idx", row["x"]), Sequence("idy", row["y"]))
        assert r.metric == metric and r.x.id == "idx" and r.y.id == "idy"
        if expected is None:
            assert r.d is None
        else:
            assert abs(r.d - expected) <= METRICS["tolerance"]


def test_metrics_exact():
    for case in METRICS["exact"]:
        r = DistanceMetric.fromLabel(case["metric"]).calculate(Sequence("idx", case["x"]), Sequence("idy", case["y"]))
        assert r.d == case["expected"]


def test_calculate_batch_matches_single_calls():
    aligner = PairwiseAligner.Biopython()
    seqs = [Sequence(f"s{k}", s) for k, s in enumerate(["ACGTACGTTGCA", "ACGTTCGTGCA", "TTGTACGAAGCA", "ACGNACG"])]
    pairs = list(aligner.align_pairs(SequencePairs.fromProduct(Sequences(seqs), Sequences(seqs))))
    metrics = [DistanceMetric.Uncorrected(), DistanceMetric.UncorrectedWithGaps(), DistanceMetric.JukesCantor(), DistanceMetric.Kimura2P()]
    batch = DistanceMetric.calculate_batch(metrics, pairs)
    single = [m.calculate(p.x, p.y) for p in pairs for m in metrics]
    assert batch == single


def test_best_matches_first_minimum():
    """Engine.best_matches = per-query first minimum (versus_reference.py:184-188), on device."""
    import numpy as np

    from synth import coi_like
    from taxi2_b200.engine import default_engine

    seqs = coi_like(90, seed=11)
    queries, refs = seqs[:30], seqs[30:] + [seqs[31], seqs[31]]   # duplicated references: ties -> earliest
    eng = default_engine()
    eng.set_scores(None)
    eng.load(queries, 0)
    eng.load(refs, 1)
    full = eng.align_rect(0, len(queries), 0, len(refs), want=("metrics", "counts"))
    got = eng.best_matches(metric=0, rows_per_tile=7)
    for q in range(len(queries)):
        col = full["metrics"][q, :, 0]
        want = int(np.nanargmin(col))     # numpy's argmin returns the first minimum too
        assert got["index"][q] == want
        assert np.array_equal(got["metrics"][q], full["metrics"][q, want], equal_nan=True)
        assert np.array_equal(got["counts"][q], full["counts"][q, want])
    eng.load(["NNNN"], 0)
    eng.load(["ACGT"], 1)
    assert eng.best_matches()["index"][0] == -1
