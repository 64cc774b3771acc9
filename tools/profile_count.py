"""The tensor-core counting kernel alone at BASELINE C2 size, for one ncu capture with source counters:
ncu --set full --import-source on -k regex:count_tc_kernel -s 1 -c 1 python tools/profile_count.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402
from bench import make_prealigned  # noqa: E402
from taxi2_b200.engine import Engine  # noqa: E402

data, off = make_prealigned(9000)
n = len(off) - 1
eng = Engine(0)
eng.set_option("count_kernel", 2)
eng.load((data, off), 0)
c = torch.empty((n * n, 4), dtype=torch.int32, device="cuda")
m = torch.empty((n * n, 4), dtype=torch.float64, device="cuda")
for _ in range(2):
    eng.count_rect_device(0, n, 0, n, c.data_ptr(), m.data_ptr())
    eng.sync()
print("kernel", eng.last_kernel, "ms", eng.stats()["kernel_ms"])
