"""Quick device-side throughput probe (not the bench contract): GCUPS of one rect launch."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
import torch
from synth import coi_like
from taxi2_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
seqs = coi_like(n, seed=650)
eng = Engine(0)
eng.load(seqs, 0)
counts = torch.empty((n * n, 4), dtype=torch.int32, device="cuda")
metrics = torch.empty((n * n, 4), dtype=torch.float64, device="cuda")
for it in range(3):
    t0 = time.perf_counter()
    eng.align_rect_device(0, n, 0, n, 0, counts.data_ptr(), metrics.data_ptr())
    eng.sync()
    dt = time.perf_counter() - t0
    st = eng.stats()
    print(json.dumps(dict(n=n, pairs=n * n, wall_s=round(dt, 4), kernel_ms=round(st["kernel_ms"], 3), cells=st["cells"],
                          gcups=round(st["cells"] / max(st["kernel_ms"], 1e-9) / 1e6, 1))))
    # stats accumulate across device-variant calls: reset by differencing
    eng._lib.taxi_last_stats  # noqa
    break_ = False
print("sum counts", counts.sum(dim=0).tolist())
