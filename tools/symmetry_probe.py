"""How often does the orientation of a pair matter?  The reference aligns (x, y) AND (y, x)
(versus_all.py:746); the second is the transpose of the first unless a tie on the traced path is
broken differently (Biopython prefers Ix over Iy, which swap roles under transposition).  This
probe runs the CPU oracle on both orientations of sampled C3 pairs and reports the fraction whose
alignments are NOT transposes of each other, and the fraction whose distance counts differ.
Test infrastructure (uses oracle/): python tools/symmetry_probe.py [pairs]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from synth import coi_like  # noqa: E402
from taxi2_b200.engine import pack_strings  # noqa: E402

npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
seqs = coi_like(4096, seed=650)
rng = np.random.default_rng(12)
px = rng.integers(0, len(seqs), npairs).astype(np.int32)
py = rng.integers(0, len(seqs), npairs).astype(np.int32)
data, off = pack_strings(seqs)
fwd = oracle.align_count_pairs(data, off, px, py)
rev = oracle.align_count_pairs(data, off, py, px)
counts_differ = np.nonzero((fwd["counts"] != rev["counts"]).any(axis=1))[0]
assert np.array_equal(fwd["score"], rev["score"])
paths_differ = 0
sample = min(npairs, 3000)
for k in range(sample):
    ax, ay, _ = oracle.align(seqs[px[k]], seqs[py[k]])
    bx, by, _ = oracle.align(seqs[py[k]], seqs[px[k]])
    paths_differ += (ax, ay) != (by, bx)
print(json.dumps(dict(pairs=npairs, counts_differ=int(len(counts_differ)), counts_differ_fraction=len(counts_differ) / npairs,
                      alignments_compared=sample, alignments_not_transposes=paths_differ,
                      alignments_not_transposes_fraction=paths_differ / sample)))
