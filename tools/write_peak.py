"""Write-only and copy bandwidth of the device as torch sees it (memset / fill / copy of 4 GB buffers): the
alignment-free kernels are pure writers (48 B per pair out, operands L2-resident), so the copy
figure in MEASURED_PEAKS.json -- half reads, half writes -- is not their ceiling by itself."""
import json
import torch

n = 1 << 30   # 4 GB of float32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3


t_zero = timed(lambda: a.zero_())
t_fill = timed(lambda: a.fill_(1.5))
t_copy = timed(lambda: b.copy_(a))
t_read = timed(lambda: a.sum())
print(json.dumps(dict(bytes=4 * n, memset_gbs=4 * n / t_zero / 1e9, fill_gbs=4 * n / t_fill / 1e9, copy_gbs_read_plus_write=8 * n / t_copy / 1e9,
                      read_sum_gbs=4 * n / t_read / 1e9)))
