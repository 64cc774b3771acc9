"""Throughput probes for the other BASELINE configs at reduced pair counts (same sequence
geometry): C4 = versusReference best match (queries x references, device-side first minimum),
C5 = mixed 300-1500 bp all-vs-all (rows grouped by length: one packed-kernel launch per stripe
geometry, the multi-stripe packed kernel above 1023 bp).
Prints one JSON line per config."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from synth import ALPHA, COMPOSITION, coi_like  # noqa: E402
from taxi2_b200.engine import Engine  # noqa: E402

eng = Engine(0)

# ---- C4: 4096 queries x 2048 references of the C3 generator (seed 200020) ----------------------
seqs = coi_like(6144, seed=200020)
q, r = seqs[:4096], seqs[4096:]
eng.load(q, 0)
eng.load(r, 1)
eng.best_matches(rows_per_tile=1024)          # warm-up (arena allocation)
t0 = time.perf_counter()
out = eng.best_matches(rows_per_tile=2048)
dt = time.perf_counter() - t0
cells = sum(map(len, q)) * sum(map(len, r))
print(json.dumps(dict(config="C4 (reduced): 4096 queries x 2048 references, best match + 4 metrics of the winner",
                      pairs=len(q) * len(r), seconds=round(dt, 4), pairs_per_s=len(q) * len(r) / dt, gcups=cells / dt / 1e9,
                      kernel=eng.last_kernel, winners_defined=int((out["index"] >= 0).sum()))))

# ---- C5: 1536 sequences, lengths uniform 300-1500, all-vs-all -----------------------------------
rng = np.random.default_rng(5)
base = coi_like(1536, length=1500, seed=5)
mixed = [s[: int(rng.integers(300, 1501))] for s in base]
n = len(mixed)
eng.load(mixed, 0)
counts = torch.empty((n * n, 4), dtype=torch.int32, device="cuda")
metrics = torch.empty((n * n, 4), dtype=torch.float64, device="cuda")
for it in range(2):
    t0 = time.perf_counter()
    eng.align_rect_device(0, n, 0, n, 0, counts.data_ptr(), metrics.data_ptr())
    eng.sync()
    dt = time.perf_counter() - t0
cells = sum(map(len, mixed)) ** 2
short = sum(len(s) <= 1023 for s in mixed)
print(json.dumps(dict(config="C5 (reduced): 1536 sequences of 300-1500 bp, all ordered pairs", pairs=n * n, seconds=round(dt, 4),
                      pairs_per_s=n * n / dt, gcups=cells / dt / 1e9, kernel=eng.last_kernel, rows_single_stripe=short, rows_multi_stripe=n - short)))
