"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/*.h declares
(no compute calls here); compute entry points fail loudly without a device (no CPU fallback)."""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "taxi2_b200.h"


def declared_symbols() -> list[str]:
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(taxi_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for required in ("taxi_ctx_create", "taxi_load_sequences", "taxi_set_scores", "taxi_align_pairs", "taxi_align_rect",
                     "taxi_align_rect_device", "taxi_align_strings", "taxi_count_rect", "taxi_count_pairs",
                     "taxi_argmin_rows_device", "taxi_last_error"):
        assert required in names


def test_library_exports_every_declared_symbol():
    from taxi2_b200 import _native

    lib = ctypes.CDLL(str(_native.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in {HEADER.name} but not exported"
    # and the Python binding table covers the header one to one
    assert sorted(_native.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from taxi2_b200 import _native
    from taxi2_b200.engine import Engine

    with pytest.raises(_native.TaxiNativeError):
        Engine(0)


def test_product_never_imports_the_oracle():
    for path in (ROOT / "taxi2_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(import oracle|from oracle)", text, flags=re.M), path
    for path in (ROOT / "taxi2_b200" / "csrc").iterdir():
        if path.suffix in (".cu", ".cuh") or path.name == "Makefile":
            assert "oracle" not in path.read_text().lower(), path
