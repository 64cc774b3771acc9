"""Both orientations of a C3 tile: one ordinary launch per orientation against taxi_align_rect_both
(one alignment per unordered pair, orientation-sensitive pairs re-aligned), device-resident results.
usage: both_perf.py [nx] [ny] [out.json]"""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch  # noqa: E402
from synth import coi_like  # noqa: E402
from taxi2_b200.engine import Engine  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
seqs = coi_like(nx + ny, seed=650)
eng = Engine(0)
eng.load(seqs, 0)
dev = "cuda"
c_xy = torch.empty((nx * ny, 4), dtype=torch.int32, device=dev); m_xy = torch.empty((nx * ny, 4), dtype=torch.float64, device=dev)
c_yx = torch.empty((ny * nx, 4), dtype=torch.int32, device=dev); m_yx = torch.empty((ny * nx, 4), dtype=torch.float64, device=dev)
w_yx = torch.empty((ny * nx, 4), dtype=torch.int32, device=dev); wm_yx = torch.empty((ny * nx, 4), dtype=torch.float64, device=dev)


def two_launches():
    eng.align_rect_device(0, nx, nx, ny, 0, c_xy.data_ptr(), m_xy.data_ptr())
    eng.align_rect_device(nx, ny, 0, nx, 0, w_yx.data_ptr(), wm_yx.data_ptr())
    eng.sync()


def both():
    eng.align_rect_both_device(0, nx, nx, ny, c_xy.data_ptr(), m_xy.data_ptr(), c_yx.data_ptr(), m_yx.data_ptr())


rec = {"tile": [nx, ny], "ordered_pairs": 2 * nx * ny}
for name, fn in (("two_launches", two_launches), ("both", both)):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    rec[name + "_s"] = round(min(ts), 4)
    rec[name + "_ordered_pairs_per_s"] = round(2 * nx * ny / min(ts))
rec["redo"] = eng.last_redo
rec["redo_frac"] = round(eng.last_redo / (nx * ny), 5)
rec["speedup"] = round(rec["two_launches_s"] / rec["both_s"], 3)
rec["identical"] = bool(torch.equal(c_yx, w_yx) and torch.equal(m_yx.view(torch.int64), wm_yx.view(torch.int64)))
print(json.dumps(rec))
if len(sys.argv) > 3:
    Path(sys.argv[3]).write_text(json.dumps(rec) + "\n")
