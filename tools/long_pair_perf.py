"""One very long pair (default 12 000 x 9 000 bp, beyond the packed kernel's 16-bit window) on the
general kernel: one warp running its 18 stripes one after the other (no_coop) vs the intra-task
kernel (the stripes pipelined over the 8 warps of a CTA).  Prints one JSON line.
Usage: python tools/long_pair_perf.py [len_x] [len_y] [pairs]"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
from synth import random_pairs  # noqa: E402
from taxi2_b200.engine import Engine  # noqa: E402

la = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
lb = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
npairs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rng = np.random.default_rng(12000)
xs, ys = random_pairs(rng, npairs, la, la, sub=0.08, indel=0.01)
ys = [y[:lb] for y in ys]
eng = Engine(0)
eng.load(xs, 0)
eng.load(ys, 1)
px = np.arange(npairs, dtype=np.int32)
out = {}
for name, no_coop in (("one_warp_per_pair", 1), ("intra_task", 0)):
    eng.set_option("no_coop", no_coop)
    eng.align_pairs(px, px)                      # warm-up: arena allocation
    t0 = time.perf_counter()
    res = eng.align_pairs(px, px)
    dt = time.perf_counter() - t0
    out[name] = dict(seconds=round(dt, 4), kernel_ms=round(eng.stats()["kernel_ms"], 2), kernel=eng.last_kernel, score=int(res["score"][0]))
cells = sum(len(x) * len(y) for x, y in zip(xs, ys))
print(json.dumps(dict(pair=f"{la} x {lb} bp", pairs=npairs, cells=cells, **out,
                      speedup=round(out["one_warp_per_pair"]["kernel_ms"] / out["intra_task"]["kernel_ms"], 2),
                      gcups_intra_task=round(cells / out["intra_task"]["kernel_ms"] / 1e6, 1))))
