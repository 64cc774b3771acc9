"""The versusAll sub-record of bench.py on its own, three times in a row (run-to-run spread of the
whole-job time): all ordered pairs of the first 8192 C3 sequences from one alignment per unordered
pair.  Usage: python tools/versus_all_perf.py [n_gpus]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
data, off = bench.make_sequences(bench.SYM_N)
for rep in range(3):
    r = bench.versus_all_symmetric(data, off, world)
    print(json.dumps({k: r[k] for k in ("seconds", "seconds_of_each_run", "value", "kernel_seconds_sum", "realigned", "identical_to_ordered_path")}))
