"""End-to-end task timing: VersusReference on nq COI-like queries x nr references (C4 geometry,
reduced), all output files except aligned pairs, block path on and off.  One JSON line each.
Usage: task_perf_reference.py [nq] [nr] [native|python|both]"""
import json
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from synth import coi_like  # noqa: E402
from taxi2_b200.sequences import Sequence, Sequences  # noqa: E402
from taxi2_b200.tasks import VersusReference  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
which = sys.argv[3] if len(sys.argv) > 3 else "both"
raw = coi_like(nq + nr, seed=200020)
records = [Sequence(f"seq{k}", s.decode(), {"organism": f"Genus{k % 50} species{(k // 50) % 20}"}) for k, s in enumerate(raw)]
for native in ([True, False] if which == "both" else [which == "native"]):
    task = VersusReference()
    task.work_dir = Path(tempfile.mkdtemp())
    task.progress_handler = lambda *a: None
    task.input.data, task.input.reference = Sequences(records[:nq]), Sequences(records[nq:])
    task.params.pairs.write = False
    task.native_writers = native
    t0 = time.perf_counter()
    task.start()
    dt = time.perf_counter() - t0
    size = sum(f.stat().st_size for f in task.work_dir.rglob("*") if f.is_file())
    print(json.dumps(dict(task="VersusReference", queries=nq, references=nr, pairs=nq * nr, block_path=native, seconds=round(dt, 2),
                          pairs_per_s=round(nq * nr / dt), output_mb=round(size / 1e6, 1))), flush=True)
