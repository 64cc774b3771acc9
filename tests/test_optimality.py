"""The oracle's alignments are OPTIMAL, established without its traceback: every alignment it
emits, re-scored column by column under the six scores, equals the optimum of an independent
score-only dynamic programme -- on the reference's 53 vectors (where the re-scored reference
solutions also pin the scorer itself) and on random tie-rich pairs under many score sets."""
from __future__ import annotations

import json

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from rescore import best_score, rescore
from synth import random_pairs

ALIGN = json.loads((GOLDEN / "align_cases.json").read_text())
SCORE_SETS = [(1, -1, -8, -1, -1, -1), (1, 0, 0, 0, 0, 0), (2, -3, -5, -2, -4, -1), (1, -1, -1, -1, -2, -2), (10, 0, -10, -6, 0, 0),
              (1, -1, -8, -1, -3, -2), (1, 0, 0, -2, -1, 0), (0, 1, -1, 0, 0, 0), (3, -4, -90, -1, -1, -1), (5, -4, -10, -4, -4, -4)]


@pytest.mark.parametrize("case", ALIGN["align_tests"] + ALIGN["align_tests_failing"], ids=lambda c: f"{c['x']}-{c['y']}-{c['scores']}")
def test_reference_solutions_are_optimal_and_equally_scored(case):
    """Every alignment the reference's tests accept for a case has the same score, and it is the
    optimum: this checks rescore() and best_score() against the reference's own expectations."""
    best = best_score(case["x"], case["y"], case["scores"])
    for ax, ay in case["solutions"]:
        assert ax.replace("-", "") == case["x"] and ay.replace("-", "") == case["y"]
        assert rescore(ax, ay, case["scores"]) == best
    ox, oy, score = oracle.align(case["x"], case["y"], case["scores"])
    assert score == best and rescore(ox, oy, case["scores"]) == best
    if "biopython" in case:
        assert case["biopython"]["score"] == best


@pytest.mark.parametrize("scores", SCORE_SETS)
def test_oracle_alignments_are_optimal_on_random_pairs(scores):
    rng = np.random.default_rng(abs(hash(scores)) % (2**32))
    for alphabet, n in ((b"ACGTN", 60), (b"AT", 60)):
        xs, ys = random_pairs(rng, n, 1, 45, sub=0.3, indel=0.1, alphabet=alphabet)
        for x, y in zip(xs, ys):
            x, y = x.decode(), y.decode()
            ax, ay, score = oracle.align(x, y, scores)
            assert ax.replace("-", "") == x and ay.replace("-", "") == y
            assert score == rescore(ax, ay, scores) == best_score(x, y, scores), (x, y, scores)
