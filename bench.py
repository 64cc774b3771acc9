#!/usr/bin/env python
"""bench.py -- headline benchmark of the TaxI2 pairwise-distance path on B200.

Metric (BASELINE.json): aligned pairs/sec & GCUPS, all-pairs ~650 bp COI-like sequences.
Workload: BASELINE config C3 -- 50 000 synthetic COI-like sequences (seed 650), all ordered
pairs (reference semantics, versus_all.py:746).  One "step" is one TILE_X x TILE_Y tile of that
50k x 50k pair matrix (global Gotoh alignment with Biopython's first-path tie-breaking + the four
distance metrics per pair).  With N GPUs every rank takes its own tile per step (static tile
assignment, weak scaling, no collective on the compute path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs already in HBM,
outputs left in HBM); `e2e` goes through the C-ABI with host buffers (sequence upload + result
download inside the timed region).  `--impl reference` times the CPU restatement of the reference
path (oracle/, "port": Biopython and the Rust distance crate are not installable here) on all
host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

N_SEQ = 50_000
SEQ_LEN = 650
TILE_X = 1536
TILE_Y = 2048
OPS_PER_CELL = 13          # scalar INT32 ops of the score-only 3-state recurrence
INT32_PEAK_FALLBACK = 33.2e12  # lane-ops/s, profiles/int_peak_r01.jsonl (VIMNMX/IADD3, 128 lanes/clk/SM)


def make_sequences(n: int):
    from synth import coi_like
    from taxi2_b200.engine import pack_strings

    seqs = coi_like(n, length=SEQ_LEN, seed=650)
    return pack_strings(seqs)


def tile_of(step: int, rank: int, world: int, n: int) -> tuple[int, int]:
    """Static tile assignment: the ordered pair matrix is cut into TILE_X x TILE_Y tiles,
    enumerated row-major; (step, rank) -> tile index step*world + rank."""
    tiles_y = n // TILE_Y
    tiles_x = n // TILE_X
    t = (step * world + rank) % (tiles_x * tiles_y)
    return (t // tiles_y) * TILE_X, (t % tiles_y) * TILE_Y


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows: list[list[str]] = []
        self.proc = None
        self.gpu = gpu_index

    def __enter__(self):
        if self.gpu < 0:
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # under load = samples above half the max clock
        load = [s for s in sm if mx and s > 0.5 * max(mx)] or sm
        return dict(sm_mhz=float(np.median(load)) if load else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def measure_int32_peak() -> tuple[float, str]:
    """Live IADD3/VIMNMX issue-rate microbenchmark (tools/int_peak.cu) -> lane-ops/s."""
    exe = ROOT / "tools" / "bin" / "int_peak"
    try:
        out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120).stdout
        best = 0.0
        for line in out.splitlines():
            rec = json.loads(line)
            if rec.get("op") in ("IADD3", "VIMNMX") and rec.get("threads_per_sm") == 1024:
                best = max(best, rec["gops_per_s"] * 1e9)
        if best > 0:
            return best, "measured live by tools/bin/int_peak (IADD3/VIMNMX, 1024 threads/SM)"
    except (OSError, subprocess.SubprocessError, ValueError):
        pass
    return INT32_PEAK_FALLBACK, "fallback: profiles/int_peak_r01.jsonl"


def cpu_oracle_throughput(data, off, seconds: float, threads: int = 0) -> dict:
    """Times the CPU restatement (oracle/) on a bounded sample of the same workload."""
    import oracle

    n = len(off) - 1
    rng = np.random.default_rng(1)
    threads = threads or oracle.max_threads()
    lens = np.diff(off)
    pairs = cells = 0
    t0 = time.perf_counter()
    chunk = max(64, 32 * threads)
    while True:
        px = rng.integers(0, n, size=chunk).astype(np.int32)
        py = rng.integers(0, n, size=chunk).astype(np.int32)
        oracle.align_count_pairs(data, off, px, py, None, threads)
        pairs += chunk
        cells += int((lens[px] * lens[py]).sum())
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    return dict(pairs=pairs, cells=cells, seconds=dt, threads=threads)


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    data, off = make_sequences(4096)
    per_step = 4.0
    for _ in range(args.warmup):
        cpu_oracle_throughput(data, off, 0.5)
    t0 = time.perf_counter()
    pairs = cells = 0
    threads = 0
    for _ in range(args.steps):
        r = cpu_oracle_throughput(data, off, per_step)
        pairs += r["pairs"]; cells += r["cells"]; threads = r["threads"]
    dt = time.perf_counter() - t0
    value = pairs / dt
    sample = f"{pairs} random ordered pairs of the C3 generator (first 4096 sequences), ~{per_step:.0f} s per step"
    line = dict(
        impl="reference", metric="aligned_pairs_per_sec", value=value, unit="pairs/s", gcups=cells / dt / 1e9,
        n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * dt / args.steps,
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
        config=workload_config(world),
        cpu_baseline=dict(value=value, unit="pairs/s", cores=threads, kind="port", sample=sample),
        e2e=dict(value=value, unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        note="CPU restatement of Biopython PairwiseAligner + calculate_distances (oracle/), not the binaries themselves",
    )
    print(json.dumps(line), flush=True)


def workload_config(world: int) -> dict:
    return dict(
        workload=f"C3: all ordered pairs of {N_SEQ} synthetic COI-like sequences (~{SEQ_LEN} bp, seed 650); "
                 f"step = one {TILE_X}x{TILE_Y} tile of the pair matrix per GPU",
        pairs_per_step_per_gpu=TILE_X * TILE_Y, scores="match 1, mismatch -1, internal open -8 / extend -1, end open -1 / extend -1",
        metrics="p, p-gaps, jc, k2p", sharding=f"static tile assignment over {world} GPU(s), no data-path collective",
        l2="256 MiB L2 flush between steps; the per-step traceback arena (>2 GB) exceeds L2 on its own",
    )


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="taxi2_b200", choices=["taxi2_b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--nseq", type=int, default=N_SEQ, help=argparse.SUPPRESS)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # a fresh checkout has no built artefacts (kept out of history): rank 0 builds, the others wait
    needed = [ROOT / "taxi2_b200" / "lib" / "libtaxi2_b200.so", ROOT / "oracle" / "libtaxi_oracle.so", ROOT / "tools" / "bin" / "int_peak"]
    if not all(p.exists() for p in needed):
        if local_rank == 0:
            import __graft_entry__

            __graft_entry__.build()
        else:
            deadline = time.time() + 600
            while not all(p.exists() for p in needed) and time.time() < deadline:
                time.sleep(2)
            time.sleep(5)   # let the linker finish writing

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from taxi2_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; taxi2_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line.  This image exports NCCL_DEBUG=VERSION, whose only
        # effect is a "NCCL version ..." line on stdout (NCCL_DEBUG_FILE does not move that one);
        # any other level the caller asked for is kept, with its output sent to stderr.
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            del os.environ["NCCL_DEBUG"]
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.nseq
    data, off = make_sequences(n)
    eng = Engine(local_rank)
    eng.load((data, off), 0)

    npairs = TILE_X * TILE_Y
    d_counts = torch.empty((npairs, 4), dtype=torch.int32, device="cuda")
    d_metrics = torch.empty((npairs, 4), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    lens = np.diff(off)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    debug = bool(os.environ.get("TAXI_BENCH_DEBUG"))

    def step_device(k: int) -> int:
        x0, y0 = tile_of(k, rank, world, n)
        ta = time.perf_counter()
        flush.zero_()
        torch.cuda.synchronize()
        tb = time.perf_counter()
        eng.align_rect_device(x0, TILE_X, y0, TILE_Y, 0, d_counts.data_ptr(), d_metrics.data_ptr())
        tc = time.perf_counter()
        eng.sync()
        td = time.perf_counter()
        if debug:
            print(f"step {k}: flush {1e3*(tb-ta):.1f} ms, enqueue {1e3*(tc-tb):.1f} ms, sync {1e3*(td-tc):.1f} ms", file=sys.stderr)
        return int(lens[x0:x0 + TILE_X].sum()) * int(lens[y0:y0 + TILE_Y].sum())

    # ---- device-resident throughput ("value") ---------------------------------------------------
    for k in range(args.warmup):
        step_device(k)
    st0 = eng.stats()
    barrier()
    with ClockSampler(local_rank if not os.environ.get('TAXI_NO_SAMPLER') else -1) as clocks:
        t0 = time.perf_counter()
        cells = 0
        for k in range(args.steps):
            cells += step_device(args.warmup + k)
        barrier()
        dt = time.perf_counter() - t0
    st1 = eng.stats()
    kernel_ms = st1["kernel_ms"] - st0["kernel_ms"]
    launches = st1["launches"] - st0["launches"]
    kernel_id = eng.last_kernel
    checksum = int(d_counts.sum().item())

    # ---- end to end through the C ABI with host buffers ("e2e") --------------------------------
    from taxi2_b200.engine import PinnedArray

    eng2 = Engine(local_rank)
    h2d = d2h = 0
    # inputs sit in page-locked host memory: the sequence bytes as a whole, the per-tile offsets
    # in two small pinned scratch arrays that are rewritten every step
    pinned_data = PinnedArray(data.shape, np.uint8)
    pinned_data.array[:] = data
    pinned_xoff = PinnedArray((TILE_X + 1,), np.int64)
    pinned_yoff = PinnedArray((TILE_Y + 1,), np.int64)

    def step_e2e(k: int) -> None:
        nonlocal h2d, d2h
        x0, y0 = tile_of(k, rank, world, n)
        np.subtract(off[x0:x0 + TILE_X + 1], off[x0], out=pinned_xoff.array)
        np.subtract(off[y0:y0 + TILE_Y + 1], off[y0], out=pinned_yoff.array)
        xs = (pinned_data.array[off[x0]:off[x0 + TILE_X]], pinned_xoff.array)
        ys = (pinned_data.array[off[y0]:off[y0 + TILE_Y]], pinned_yoff.array)
        ta = time.perf_counter()
        eng2.load(xs, 0)
        eng2.load(ys, 1)
        tb = time.perf_counter()
        out = eng2.align_rect(0, TILE_X, 0, TILE_Y, want=("counts", "metrics"), pinned=True)
        if debug:
            print(f"e2e step {k}: load {1e3*(tb-ta):.1f} ms, align_rect {1e3*(time.perf_counter()-tb):.1f} ms, "
                  f"kernel {eng2.stats()['kernel_ms']:.1f} ms", file=sys.stderr)
        h2d = xs[0].nbytes + xs[1].nbytes + ys[0].nbytes + ys[1].nbytes
        d2h = out["counts"].nbytes + out["metrics"].nbytes

    for k in range(min(args.warmup, 3)):   # untimed: buffers reach their steady size
        step_e2e(k)
    barrier()
    t1 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 3))
    for k in range(e2e_steps):
        step_e2e(args.warmup + k)
    barrier()
    dt_e2e = time.perf_counter() - t1

    # ---- aggregate over ranks (max time, sum work) ----------------------------------------------
    if distributed:
        t = torch.tensor([dt, dt_e2e, kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e, kernel_ms = (float(v) for v in t.tolist())
        w = torch.tensor([cells, launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        cells, launches = (int(v) for v in w.tolist())

    if rank == 0:
        pairs_total = npairs * args.steps * world
        value = pairs_total / dt
        gcups = cells / dt / 1e9
        kernel_s = kernel_ms / 1e3 / 1.0
        peak, peak_how = measure_int32_peak()
        cells_per_launch = cells / max(launches, 1)
        launch_s = kernel_s / max(args.steps, 1)
        achieved = cells_per_launch * OPS_PER_CELL / launch_s
        # traceback codes written per cell: packed kernel 48 B per (lane, column) = 42 cells;
        # general kernel 24 B per 21 cells
        trace_bytes_per_cell = 48.0 / 42.0 if kernel_id in (16, 17) else 24.0 / 21.0
        kernel_name = {16: "gotoh_pair16_kernel<21,0>", 17: "gotoh_pair16_kernel<21,1>"}.get(kernel_id, "gotoh_warp_kernel<21>")
        cpu = cpu_oracle_throughput(data[: off[4096]], off[:4097], args.cpu_seconds)
        line = dict(
            metric="aligned_pairs_per_sec", value=value, unit="pairs/s", gcups=gcups,
            n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * dt / args.steps,
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u16x2" if kernel_id in (16, 17) else "int32", data="synthetic",
            config=workload_config(world),
            roofline=dict(
                bound="int32_alu", kernel=kernel_name, achieved=achieved / 1e9, peak=peak / 1e9, unit="Gop/s",
                frac=achieved / peak,
                # DRAM bytes per launch: ncu --set full measured 1.31 B/cell (1.17 written + 0.14 read) for this kernel
                # (profiles/ncu_gotoh_pair16_21_r01.json), scaled to this launch's cells
                traffic=(1.31 if kernel_id in (16, 17) else 1.42) * cells_per_launch,
                how=f"{OPS_PER_CELL} algorithmic INT32 ops/cell x {cells_per_launch:.3e} cells/launch / {launch_s * 1e3:.1f} ms (CUDA events on the launch stream); peak {peak_how}",
                hbm=dict(achieved=cells_per_launch * trace_bytes_per_cell / launch_s / 1e9, peak=peak_hbm(), unit="GB/s",
                         note="algorithmic HBM traffic of the same kernel: traceback codes written once; read back sparsely"),
            ),
            cpu_baseline=dict(value=cpu["pairs"] / cpu["seconds"], unit="pairs/s", gcups=cpu["cells"] / cpu["seconds"] / 1e9,
                              cores=cpu["threads"], kind="port",
                              sample=f"{cpu['pairs']} random ordered pairs of the first 4096 C3 sequences in {cpu['seconds']:.1f} s"),
            e2e=dict(value=npairs * e2e_steps * world / dt_e2e, unit="pairs/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                     steps=e2e_steps),
            gpu_launches=launches, clocks=clocks.summary(), checksum=checksum,
        )
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


def peak_hbm() -> float:
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        return 6650.0


if __name__ == "__main__":
    main()
