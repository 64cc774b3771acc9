"""End-to-end task timing: VersusAll on n COI-like sequences with all distance / summary / subset
outputs, native batch writers on and off.  Prints one JSON line each.
Usage: task_perf.py [n] [native|python] [pairs]   ("pairs": also write align/aligned_pairs.txt)"""
import json
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from synth import coi_like  # noqa: E402
from taxi2_b200.partitions import Partition  # noqa: E402
from taxi2_b200.sequences import Sequence, Sequences  # noqa: E402
from taxi2_b200.tasks import VersusAll  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
modes = [True, False] if len(sys.argv) < 3 or sys.argv[2] == "both" else [sys.argv[2] == "native"]
write_pairs = len(sys.argv) > 3 and sys.argv[3] == "pairs"
raw = coi_like(n, seed=650)
records = [Sequence(f"seq{k}", s.decode(), {"organism": f"Genus{k % 50} species{(k // 50) % 20}"}) for k, s in enumerate(raw)]
species = Partition({r.id: r.extras["organism"] for r in records})
genera = Partition({r.id: r.extras["organism"].split(" ")[0] for r in records})
for native in modes:
    task = VersusAll()
    task.work_dir = Path(tempfile.mkdtemp())
    task.progress_handler = lambda *a: None
    task.input.sequences = Sequences(records)
    task.input.species, task.input.genera = species, genera
    task.params.pairs.write = write_pairs
    task.native_writers = native
    t0 = time.perf_counter()
    task.start()
    dt = time.perf_counter() - t0
    size = sum(f.stat().st_size for f in task.work_dir.rglob("*") if f.is_file())
    print(json.dumps(dict(task="VersusAll", sequences=n, pairs=n * n, native_writers=native, aligned_pairs=write_pairs, seconds=round(dt, 2),
                          pairs_per_s=round(n * n / dt), output_mb=round(size / 1e6, 1))), flush=True)
