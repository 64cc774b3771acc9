"""Where the end-to-end step (bench.py `e2e`) spends its time: load x, load y, align_rect with
host buffers (kernel + result download).  Usage: python tools/e2e_breakdown.py [steps]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
from taxi2_b200.engine import Engine  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
data, off = bench.make_sequences(8192)
eng = Engine(0)
n = len(off) - 1
for k in range(steps):
    x0, y0 = bench.tile_of(k, 0, 1, n)
    xs = (data[off[x0]:off[x0 + bench.TILE_X]], off[x0:x0 + bench.TILE_X + 1] - off[x0])
    ys = (data[off[y0]:off[y0 + bench.TILE_Y]], off[y0:y0 + bench.TILE_Y + 1] - off[y0])
    t0 = time.perf_counter(); eng.load(xs, 0)
    t1 = time.perf_counter(); eng.load(ys, 1)
    t2 = time.perf_counter(); out = eng.align_rect(0, bench.TILE_X, 0, bench.TILE_Y, want=("counts", "metrics"), pinned=True)
    t3 = time.perf_counter()
    st = eng.stats()
    print(f"step {k}: load x {1e3*(t1-t0):.1f} ms, load y {1e3*(t2-t1):.1f} ms, align_rect {1e3*(t3-t2):.1f} ms "
          f"(kernel {st['kernel_ms']:.1f} ms, {st['launches']} launch(es)), checksum {int(out['counts'].sum())}", flush=True)
