"""One launch of every hot kernel, for a single ncu capture:
packed bottom-aligned (650 bp), packed multi-stripe (1 200 bp), general int32 (650 bp),
alignment-free rectangle (618 columns).  Usage: python tools/profile_kernels.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from synth import coi_like  # noqa: E402
from taxi2_b200.engine import Engine  # noqa: E402

eng = Engine(0)


def rect(n, length, **options):
    for k, v in options.items():
        eng.set_option(k, v)
    eng.load(coi_like(n, length=length, seed=650), 0)
    counts = torch.empty((n * n, 4), dtype=torch.int32, device="cuda")
    metrics = torch.empty((n * n, 4), dtype=torch.float64, device="cuda")
    eng.align_rect_device(0, n, 0, n, 0, counts.data_ptr(), metrics.data_ptr())
    eng.sync()
    print(length, options, "kernel", eng.last_kernel, "ms", round(eng.stats()["kernel_ms"], 2))
    for k in options:
        eng.set_option(k, 0)


rect(384, 650)
rect(256, 1200)
rect(256, 650, force_general=1)
# alignment-free rectangle at BASELINE C2 size: 9000 pre-aligned rows x 618 columns, one launch
sys.path.insert(0, str(ROOT))
from bench import make_prealigned  # noqa: E402

data, off = make_prealigned(9000)
n = len(off) - 1
eng.load((data, off), 0)
c = torch.empty((n * n, 4), dtype=torch.int32, device="cuda")
m = torch.empty((n * n, 4), dtype=torch.float64, device="cuda")
for kernel in (1, 2, 2):   # popcount kernel, then the tensor-core kernel twice (the first launch builds its operands)
    eng.set_option("count_kernel", kernel)
    before = eng.stats()["kernel_ms"]
    eng.count_rect_device(0, n, 0, n, c.data_ptr(), m.data_ptr())
    eng.sync()
    print("count", {1: "popcount", 2: "tensor cores"}[kernel], n, "x", n, "ms", round(eng.stats()["kernel_ms"] - before, 3))
