"""Where a VersusAll run with every output file spends its time: wall seconds inside the engine's
string / rectangle calls and inside each native writer, summed over the host threads that run them
(they overlap, so the parts can exceed the total).  Usage: task_breakdown.py [n] [pairs]"""
import collections
import json
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from taxi2_b200 import fastwrite  # noqa: E402
from taxi2_b200.engine import Engine  # noqa: E402

total = collections.Counter()
lock = threading.Lock()


def timed(owner, name):
    inner = getattr(owner, name)

    def outer(*a, **k):
        t0 = time.perf_counter()
        try:
            return inner(*a, **k)
        finally:
            with lock:
                total[name] += time.perf_counter() - t0
    setattr(owner, name, outer)


for fn in ("format_pairs", "format_matrix", "format_aligned_pairs", "format_subset_rows", "format_subset_matrix"):
    timed(fastwrite, fn)
for fn in ("align_strings_raw", "align_rect", "align_rect_both", "load"):
    timed(Engine, fn)

sys.argv = [sys.argv[0], sys.argv[1] if len(sys.argv) > 1 else "2000", "native", *sys.argv[2:]]
exec(compile((ROOT / "tools" / "task_perf.py").read_text().replace("Path(__file__)", f"Path({str(ROOT / 'tools' / 'task_perf.py')!r})"), "task_perf.py", "exec"))
print(json.dumps({k: round(v, 2) for k, v in total.items()}))
