import sys, json, time
sys.path.insert(0, "/root/repo")
import bench
data, off = bench.make_sequences(bench.SYM_N)
for rep in range(3):
    r = bench.versus_all_symmetric(data, off, 1)
    print(json.dumps({k: r[k] for k in ("seconds", "value", "kernel_seconds_sum", "realigned", "identical_to_ordered_path")}))
