"""One small launch of every kernel and variant, for compute-sanitizer (memcheck / racecheck):
packed aligner top-aligned, bottom-aligned and multi-stripe, general int32 (Gotoh and NW score
sets), strings + metrics in one launch, pair lists of mixed lengths, mixed-length rectangle (row
groups), alignment-free rectangle and pair list, best rows, encoders.
Usage: compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
from synth import coi_like, random_pairs  # noqa: E402
from taxi2_b200.engine import Engine  # noqa: E402

eng = Engine(0)
rng = np.random.default_rng(1)
kernels = []


def note(what):
    kernels.append((what, eng.last_kernel))


seqs = coi_like(24, length=650, seed=3)
for general, top in ((0, 0), (0, 1), (1, 0)):
    eng.set_option("force_general", general)
    eng.set_option("force_top", top)
    eng.set_scores(None)
    eng.load(seqs[:9], 0)
    eng.load(seqs[9:], 1)
    eng.align_rect(0, 9, 0, 15)
    note(f"rect 9x15 general={general} top={top}")
    px = np.arange(9, dtype=np.int32)
    eng.align_strings_raw(px, px, want=("counts", "metrics"))
    note("strings+metrics")
eng.set_option("force_general", 0)
eng.set_option("force_top", 0)
xs, ys = random_pairs(rng, 12, 1100, 1500, sub=0.1, indel=0.02)
eng.load(xs, 0); eng.load(ys, 1)
eng.align_pairs(np.arange(12, dtype=np.int32), np.arange(12, dtype=np.int32))
note("multi-stripe pair list")
mixed, _ = random_pairs(rng, 20, 100, 1400, sub=0.1, indel=0.02)
eng.load(mixed, 0)
eng.align_rect(0, 20, 0, 20, want=("metrics",))
note("mixed-length rectangle")
eng.set_scores((1, 0, 0, 0, 0, 0))
eng.load(seqs[:6], 0)
eng.align_rect(0, 6, 0, 6)
note("NW score set")
eng.set_scores(None)
al = np.frombuffer(b"ACGT-N", dtype=np.uint8)
rows = [al[rng.choice(6, 618, p=[.24, .24, .24, .24, .03, .01])].tobytes() for _ in range(300)]
eng.load(rows, 0)
eng.count_rect(0, 300, 0, 300)
eng.count_pairs(np.arange(300, dtype=np.int32), np.arange(300, dtype=np.int32)[::-1].copy())
eng.best_rows(0, 300, 0, 300, 0, align=False)
eng.load(seqs[:8], 0); eng.load(seqs[8:20], 1)
eng.best_rows(0, 8, 0, 12, 0, align=True)
eng.load(seqs, 0)
eng.align_rect_both(0, 10, 10, 14)
note(f"both orientations 10x14 (re-aligned: {eng.last_redo})")
for tile_x in (128, 64):
    eng.set_option("count_kernel", 2); eng.set_option("tc_tile_x", tile_x)
    eng.load(rows, 0)
    eng.count_rect(0, 300, 0, 300)
    note(f"tensor-core counting kernel, x tile {tile_x}")
eng.set_option("count_kernel", 0)
print("kernels:", kernels)
print("sanitize_run ok")
