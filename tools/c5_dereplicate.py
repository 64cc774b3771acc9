"""BASELINE config C5 through the product path: the Dereplicate task (similarity 0.07, length 10,
no pair / distance files) on n sequences of 300-1500 bp of the C3 species tree.  Prints one JSON
line: wall time, pairs the reference would have visited, pairs the device computed (a superset:
blocks are scheduled ahead of the exclusions), reloads, survivors.
Usage: python tools/c5_dereplicate.py [n] [gpus]"""
import json
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
from synth import coi_like  # noqa: E402
from taxi2_b200.files import FileFormat  # noqa: E402
from taxi2_b200.sequences import Sequence, Sequences  # noqa: E402
from taxi2_b200.tasks import Dereplicate  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t0 = time.perf_counter()
rng = np.random.default_rng(5)
base = coi_like(n, length=1500, seed=5)
records = [Sequence(f"seq{k}", s[: int(rng.integers(300, 1501))].decode(), {"organism": f"Genus{k % 50} species{(k // 50) % 20}"})
           for k, s in enumerate(base)]
t_gen = time.perf_counter() - t0
lens = np.array([len(r.seq) for r in records])
with tempfile.TemporaryDirectory() as tmp:
    task = Dereplicate()
    task.work_dir = Path(tmp) / "out"
    task.progress_handler = lambda *a: None
    task.devices = list(range(gpus))
    task.input = Sequences(records)
    task.output_format = FileFormat.Tabfile
    task.params.pairs.write = False
    task.params.distances.write_linear = task.params.distances.write_matricial = False
    t1 = time.perf_counter()
    res = task.start()
    dt = time.perf_counter() - t1
    kept = sum(1 for _ in open(task.paths.dereplicated)) - 1
print(json.dumps(dict(
    config=f"C5: Dereplicate (similarity 0.07, length 10, metric p, align=True) on {n} sequences of 300-1500 bp, {gpus} GPU(s)",
    n=n, seconds=dt, generation_seconds=t_gen, excluded=len(task.excluded), kept=kept,
    pairs_visited=task.stats["pairs_visited"], pairs_computed=task.stats["pairs_computed"], reloads=task.stats["reloads"],
    ordered_pairs_total=n * n, computed_pairs_per_s=task.stats["pairs_computed"] / dt,
    mean_cells_per_pair=float(lens.mean()) ** 2)))
