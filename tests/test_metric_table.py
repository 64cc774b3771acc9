"""The table form of the JC / K2P epilogue of the alignment-free kernels (common.cuh:
metrics_from_counts_table) replayed on the CPU against the oracle's libm formulas
(oracle/taxi_oracle.c:282-285): every (ts, tv) for n <= 160 and a dense sample up to n = 2048.
The tool exits non-zero if the None pattern differs anywhere or a value deviates by more than
5e-13 relative (north_star allows 1e-12)."""
from __future__ import annotations

import json
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_table_form_matches_the_floating_point_formulas(tmp_path):
    exe = tmp_path / "metrics_table_check"
    subprocess.run(["gcc", "-O2", "-o", str(exe), str(ROOT / "tools" / "metrics_table_check.c"), "-lm"], check=True)
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, run.stdout + run.stderr
    rec = json.loads(run.stdout)
    assert rec["cases"] > 5_000_000 and rec["nan_or_zero_mismatches"] == 0
    assert rec["worst_rel_jc"] < 2e-13 and rec["worst_rel_k2p"] < 2e-13
