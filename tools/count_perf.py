"""Throughput probe of the alignment-free path (BASELINE config 2): 9000 pre-aligned sequences of
618 columns = a seeded resample of Taxi2test1_120.tab padded with '-' on both sides (bench.make_prealigned) (the real
Taxi2test1_ca9000.tab is missing from the reference checkout).  Prints one JSON line."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from taxi2_b200.engine import Engine  # noqa: E402

from bench import make_prealigned  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 9000
data, off = make_prealigned(n)
width = int(off[1] - off[0])
seqs = (data, off)
import os  # noqa: E402

eng = Engine(0)
eng.set_option("count_kernel", int(os.environ.get("TAXI_COUNT_KERNEL", "0")))   # 1 = popcount, 2 = tensor cores
eng.set_option("tc_tile_x", int(os.environ.get("TAXI_TC_TILE_X", "128")))        # x rows per tile: 64 (two CTAs per SM) or 128
eng.set_option("tc_persistent", int(os.environ.get("TAXI_TC_PERSISTENT", "0")))   # 1: persistent form of the tensor-core kernel
eng.set_option("metric_tables", int(os.environ.get("TAXI_METRIC_TABLES", "1")))   # 0: floating-point JC / K2P in the count kernels
eng.load(seqs, 0)
counts = torch.empty((n * n, 4), dtype=torch.int32, device="cuda")
metrics = torch.empty((n * n, 4), dtype=torch.float64, device="cuda")
W = (width + 31) // 32
best = None
prev = 0.0
for it in range(5):
    eng.count_rect_device(0, n, 0, n, counts.data_ptr(), metrics.data_ptr())
    eng.sync()
    ms = eng.stats()["kernel_ms"] - prev
    prev = eng.stats()["kernel_ms"]
    best = ms if best is None else min(best, ms)
best_counts = None
for it in range(3):      # counts only: what the fp64 metric epilogue costs
    eng.count_rect_device(0, n, 0, n, counts.data_ptr(), 0)
    eng.sync()
    ms = eng.stats()["kernel_ms"] - prev
    prev = eng.stats()["kernel_ms"]
    best_counts = ms if best_counts is None else min(best_counts, ms)
pairs = n * n
bytes_alg = n * W * 16 + pairs * 48          # planes read once + 16 B counts + 32 B metrics per pair
print(json.dumps(dict(persistent=int(os.environ.get("TAXI_TC_PERSISTENT", "0")), metric_tables=int(os.environ.get("TAXI_METRIC_TABLES", "1")), tc_tile_x=int(os.environ.get("TAXI_TC_TILE_X", "128")), kernel={8: "count_rect_kernel (popcount)", 9: "count_tc_kernel (tcgen05 int8)"}.get(eng.last_kernel), workload=f"{n} x {n} pre-aligned pairs, {width} columns", pairs=pairs, kernel_ms=round(best, 3), counts_only_ms=round(best_counts, 3),
                      pairs_per_s=pairs / best * 1e3, hbm_gbs=bytes_alg / best / 1e6, checksum=int(counts.sum().item()))))
