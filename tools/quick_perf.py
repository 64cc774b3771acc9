"""Quick device-side throughput probe (not the bench contract): GCUPS of one rect launch.
usage: quick_perf.py [n] [option=value ...]   e.g.  quick_perf.py 1024 force_top=1 length=1200"""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch  # noqa: E402
from synth import coi_like  # noqa: E402
from taxi2_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
options = dict(opt.split("=") for opt in sys.argv[2:])
seqs = coi_like(n, length=int(options.pop("length", 650)), seed=650)
eng = Engine(0)
for k, v in options.items():
    eng.set_option(k, int(v))
eng.load(seqs, 0)
counts = torch.empty((n * n, 4), dtype=torch.int32, device="cuda")
metrics = torch.empty((n * n, 4), dtype=torch.float64, device="cuda")
prev = 0.0
for it in range(3):
    t0 = time.perf_counter()
    eng.align_rect_device(0, n, 0, n, 0, counts.data_ptr(), metrics.data_ptr())
    eng.sync()
    dt = time.perf_counter() - t0
    st = eng.stats()
    ms = st["kernel_ms"] - prev
    prev = st["kernel_ms"]
    cells = st["cells"] // (it + 1)
    print(json.dumps(dict(n=n, kernel=eng.last_kernel, wall_s=round(dt, 4), kernel_ms=round(ms, 3), gcups=round(cells / max(ms, 1e-9) / 1e6, 1))))
print("sum counts", counts.sum(dim=0).tolist())
