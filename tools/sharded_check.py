"""Multi-GPU path end to end on real devices (run under torchrun, one rank per GPU):
every rank aligns its statically assigned tiles of one pair matrix, the blocks are gathered on
rank 0 over NCCL, and rank 0 compares the result bit for bit with the same matrix computed on its
own GPU alone.  Prints one JSON line.  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29533 tools/sharded_check.py [n] [tile]"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from synth import coi_like  # noqa: E402
from taxi2_b200 import sharding  # noqa: E402
from taxi2_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local_rank)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":   # its one line would land on stdout
    del os.environ["NCCL_DEBUG"]
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

seqs = coi_like(n, seed=4242)
lens = np.array([len(s) for s in seqs])
eng = Engine(local_rank)
eng.load(seqs, 0)
tiles = sharding.make_tiles(lens, lens, tile, tile)


def compute(t):
    out = eng.align_rect(t.x0, t.nx, t.y0, t.ny, want=("counts", "metrics"))
    return np.concatenate([out["counts"].astype(np.float64), out["metrics"]], axis=2)   # (nx, ny, 8)


torch.cuda.synchronize()
t0 = time.perf_counter()
local = sharding.run_sharded(None, lens, lens, tile, tile, rank, world, compute)
t1 = time.perf_counter()
full = sharding.gather_matrix(local, tiles, (8,), np.float64) if world > 1 else None
t2 = time.perf_counter()
if rank == 0:
    alone = eng.align_rect(0, n, 0, n, want=("counts", "metrics"))
    t3 = time.perf_counter()
    want = np.concatenate([alone["counts"].astype(np.float64), alone["metrics"]], axis=2)
    if world == 1:
        full = np.zeros_like(want)
        for t in tiles:
            full[t.x0:t.x0 + t.nx, t.y0:t.y0 + t.ny] = local[t.index]
    same = bool(np.array_equal(full, want, equal_nan=True))
    plan = sharding.assign_tiles(tiles, world)
    print(json.dumps(dict(world=world, n=n, tiles=len(tiles), tiles_per_rank=[len(p) for p in plan],
                          cells_per_rank=[int(sum(t.cells for t in p)) for p in plan], identical_to_single_gpu=same,
                          sharded_compute_s=round(t1 - t0, 3), gather_s=round(t2 - t1, 3), single_gpu_s=round(t3 - t2, 3))), flush=True)
    if not same:
        sys.exit(1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
