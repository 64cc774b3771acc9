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
# both orientations of a 384 x 384 rectangle of two different row ranges: the SYM variant (tie bit per cell) + the re-alignment launch
eng.load(coi_like(768, length=650, seed=650), 0)
tc_ = torch.empty((384 * 384, 4), dtype=torch.int32, device="cuda"); tm_ = torch.empty((384 * 384, 4), dtype=torch.float64, device="cuda")
uc_ = torch.empty_like(tc_); um_ = torch.empty_like(tm_)
eng.align_rect_both_device(0, 384, 384, 384, tc_.data_ptr(), tm_.data_ptr(), uc_.data_ptr(), um_.data_ptr())
print("both orientations 384 x 384: kernel", eng.last_kernel, "re-aligned", eng.last_redo)
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
for kernel, tile_x in ((1, 128), (2, 128), (2, 128), (2, 64)):   # popcount kernel, then the tensor-core kernel (the first launch builds its operands); last: two CTAs per SM
    eng.set_option("count_kernel", kernel)
    eng.set_option("tc_tile_x", tile_x)
    before = eng.stats()["kernel_ms"]
    eng.count_rect_device(0, n, 0, n, c.data_ptr(), m.data_ptr())
    eng.sync()
    print("count", {1: "popcount", 2: f"tensor cores (x tile {tile_x})"}[kernel], n, "x", n, "ms", round(eng.stats()["kernel_ms"] - before, 3))
