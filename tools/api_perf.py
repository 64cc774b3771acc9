"""Throughput of the reference-facing Python API itself (not the bench contract): the literal
per-pair calls a TaxI2 task makes -- `aligner.align_pairs(SequencePairs.fromProduct(xs, ys))`
followed by `metric.calculate(x, y)` for the four metrics -- and the batched form the tasks here
use (`DistanceMetric.calculate_batch`).  Usage: python tools/api_perf.py [n]"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from synth import coi_like  # noqa: E402
from taxi2_b200.align import PairwiseAligner  # noqa: E402
from taxi2_b200.distances import DistanceMetric  # noqa: E402
from taxi2_b200.pairs import SequencePairs  # noqa: E402
from taxi2_b200.sequences import Sequence, Sequences  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
seqs = Sequences([Sequence(f"s{k}", s.decode(), {}) for k, s in enumerate(coi_like(n, seed=650))])
metrics = [DistanceMetric.Uncorrected(), DistanceMetric.UncorrectedWithGaps(), DistanceMetric.JukesCantor(), DistanceMetric.Kimura2P()]
aligner = PairwiseAligner.Biopython()
list(aligner.align_pairs(SequencePairs.fromProduct(Sequences(list(seqs)[:8]), Sequences(list(seqs)[:8]))))   # warm-up

t0 = time.perf_counter()
aligned = list(aligner.align_pairs(SequencePairs.fromProduct(seqs, seqs)))
t1 = time.perf_counter()
per_pair = [metric.calculate(x, y) for x, y in aligned[: 4 * n] for metric in metrics]
t2 = time.perf_counter()
batch = DistanceMetric.calculate_batch(metrics, aligned)
t3 = time.perf_counter()
assert [d.d for d in per_pair] == [d.d for d in batch[: len(per_pair)]]
print(json.dumps(dict(pairs=len(aligned), align_pairs_per_s=round(len(aligned) / (t1 - t0)),
                      calculate_per_pair_per_s=round(4 * n / (t2 - t1)), calculate_batch_pairs_per_s=round(len(aligned) / (t3 - t2)))))
