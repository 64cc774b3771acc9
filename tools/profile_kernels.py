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
rng = np.random.default_rng(9)
al = np.frombuffer(b"ACGT-N", dtype=np.uint8)
eng.load([al[rng.choice(6, 618, p=[.24, .24, .24, .24, .03, .01])].tobytes() for _ in range(4096)], 0)
c = torch.empty((4096 * 4096, 4), dtype=torch.int32, device="cuda")
m = torch.empty((4096 * 4096, 4), dtype=torch.float64, device="cuda")
eng.count_rect_device(0, 4096, 0, 4096, c.data_ptr(), m.data_ptr())
eng.sync()
print("count_rect ms", round(eng.stats()["kernel_ms"], 2))
