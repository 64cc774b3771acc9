"""Host side of VersusAll on its own (no GPU): the task driven by a stand-in engine that hands out
pre-computed random metrics instantly, so what is timed is the block iterator, the undefined-pair
bookkeeping, the native writers and the subset aggregation.  Usage: writer_perf.py [n] [pairs] [profile]
("pairs": also align/aligned_pairs.txt from synthetic gapped strings)"""
import cProfile
import json
import pstats
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
from synth import coi_like  # noqa: E402
from taxi2_b200.multi import MultiEngine  # noqa: E402
from taxi2_b200.partitions import Partition  # noqa: E402
from taxi2_b200.sequences import Sequence, Sequences  # noqa: E402
from taxi2_b200.tasks import VersusAll, versus_all  # noqa: E402


class Instant:
    """Engine stand-in: metrics of a fixed random matrix, no alignment at all."""

    def __init__(self, n):
        rng = np.random.default_rng(1)
        self.metrics = rng.random((n, n, 4))
        self.n = [n, 0]
        self.device = 0

    ny = property(lambda self: self.n[0])

    def close(self): pass
    def set_scores(self, scores): pass
    def set_option(self, key, value): pass
    def load(self, seqs, which=0): pass
    def stats(self): return dict(launches=0, cells=0, kernel_ms=0.0)

    def align_rect(self, x0, nx, y0, ny, want=("metrics",), **kw):
        return {"metrics": self.metrics[x0:x0 + nx, y0:y0 + ny]}

    def align_strings_raw(self, px, py, want=("score",), slot=None):
        """Gapped strings of the same shape the library returns: slots of len(x) + len(y) bytes, the alignment
        (here: 680 random symbols) right-aligned in each."""
        k = len(px)
        off = np.arange(k + 1, dtype=np.int64) * 1300
        start = off[1:] - 680
        if getattr(self, "_strings", None) is None or len(self._strings[0]) < k * 1300:
            rng = np.random.default_rng(2)
            self._strings = tuple(np.frombuffer(b"ACGTACGTACGTACG-", dtype=np.uint8)[rng.integers(0, 16, k * 1300, dtype=np.uint8)] for _ in range(2))
        ox, oy = (a[: k * 1300] for a in self._strings)
        px = np.asarray(px); py = np.asarray(py)
        return ox, oy, start, off, np.zeros(k, dtype=np.int32), {"metrics": self.metrics[px, py]}

    def align_strings(self, px, py):
        return [b"A"] * len(px), [b"C"] * len(px), np.zeros(len(px), dtype=np.int32)


n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
raw = coi_like(n, seed=650)
records = [Sequence(f"seq{k}", s.decode(), {"organism": f"Genus{k % 50} species{(k // 50) % 20}"}) for k, s in enumerate(raw)]
species = Partition({r.id: r.extras["organism"] for r in records})
genera = Partition({r.id: r.extras["organism"].split(" ")[0] for r in records})
multi = MultiEngine.__new__(MultiEngine)
multi.devices, multi.engines, multi.lens = [0], [Instant(n)], [None, None]
real_load = MultiEngine.load
MultiEngine.load = lambda self, seqs, which=0: (self.lens.__setitem__(which, np.array([len(s) for s in seqs])), which == 0 and self.lens.__setitem__(1, None))[0]
versus_all.task_engine = lambda task: multi
task = VersusAll()
task.work_dir = Path(tempfile.mkdtemp())
task.progress_handler = lambda *a: None
task.input.sequences = Sequences(records)
task.input.species, task.input.genera = species, genera
task.params.pairs.write = "pairs" in sys.argv[2:]
t0 = time.perf_counter()
if "profile" in sys.argv[2:]:
    prof = cProfile.Profile()
    prof.runcall(task.start)
    pstats.Stats(prof).sort_stats("cumulative").print_stats(25)
else:
    task.start()
dt = time.perf_counter() - t0
size = sum(f.stat().st_size for f in task.work_dir.rglob("*") if f.is_file())
print(json.dumps(dict(sequences=n, pairs=n * n, seconds=round(dt, 2), pairs_per_s=round(n * n / dt), output_mb=round(size / 1e6, 1),
                      mb_per_s=round(size / 1e6 / dt))))
import shutil  # noqa: E402
shutil.rmtree(task.work_dir)
