"""Python side of the native batch formatter / subset aggregator (csrc/host_format.cpp).

The reference writes every value through `str.format` and a generator `.send`
(/root/reference/src/itaxotools/taxi2/distances.py:95-186, 244-279; tasks/versus_all.py:278-350)
and aggregates every distance through nested dicts (versus_all.py:57-95, 623-645).  Here the tasks
hand whole blocks of the fp64 result matrix to the library, which appends the rows of the same
files; headers stay in Python.  Formats that are not plain "{:.Nf}" / "{:f}" / "{:.Ne}" specs keep
using the Python handlers.
"""
from __future__ import annotations

import ctypes as C
import re
from math import inf

import numpy as np

from . import _native as N

SEG_X = (0, 1, 2, 3)
SEG_Y = (4, 5, 6, 7)
SEG_SCORES = 8
SEG_COMPARISON = 9
COMPARISON_LABELS = ("no info", "intra-species", "inter-species", "intra-genus", "inter-genus")  # plot.py:15-20

_SPEC = re.compile(r"^\{:(\.\d+)?([fe])\}$")


def printf_format(python_spec: str) -> str | None:
    """'{:.4f}' -> '%.4f'; None when the spec is not a plain fixed / exponent float format."""
    m = _SPEC.match(python_spec)
    if not m:
        return None
    return "%" + (m.group(1) or "") + m.group(2)


class StringTable:
    """list of str -> concatenated utf-8 bytes + int64 offsets (kept alive for the C calls)."""

    def __init__(self, items):
        encoded = [s.encode("utf-8", "surrogateescape") for s in items]
        self.bytes = b"".join(encoded)
        self.off = np.zeros(len(encoded) + 1, dtype=np.int64)
        if encoded:
            np.cumsum([len(b) for b in encoded], out=self.off[1:])
        self._buf = C.create_string_buffer(self.bytes, max(len(self.bytes), 1))

    @property
    def bytes_ptr(self):
        return C.cast(self._buf, C.c_void_p)

    @property
    def off_ptr(self):
        return self.off.ctypes.data_as(C.c_void_p)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def format_pairs(path, segments, xtables, ytables, x0, nx, ny, metrics, undefined, columns, scale, fmt, missing,
                 xgenus=None, xspecies=None, ygenus=None, yspecies=None, threads=0) -> None:
    lib = N.load()
    seg = np.asarray(segments, dtype=np.int32)
    cols = np.asarray(columns, dtype=np.int32)
    xt = (list(xtables) + [None] * 4)[:4]
    yt = (list(ytables) + [None] * 4)[:4]
    xb = (C.c_void_p * 4)(*[(t.bytes_ptr if t else None) for t in xt])
    xo = (C.c_void_p * 4)(*[(t.off_ptr if t else None) for t in xt])
    yb = (C.c_void_p * 4)(*[(t.bytes_ptr if t else None) for t in yt])
    yo = (C.c_void_p * 4)(*[(t.off_ptr if t else None) for t in yt])
    labels = (C.c_char_p * 5)(*[s.encode() for s in COMPARISON_LABELS])
    metrics = np.ascontiguousarray(metrics, dtype=np.float64)
    N.check(lib.taxi_format_pairs(str(path).encode(), _ptr(seg), len(seg), xb, xo, yb, yo, x0, nx, ny, _ptr(metrics), _ptr(undefined),
                                  _ptr(cols), len(cols), float(scale), fmt.encode(), missing.encode(),
                                  _ptr(xgenus), _ptr(xspecies), _ptr(ygenus), _ptr(yspecies), labels, threads))


def format_matrix(path, xids: StringTable, x0, nx, ny, metrics, undefined, column, scale, fmt, missing, threads=0) -> None:
    lib = N.load()
    metrics = np.ascontiguousarray(metrics, dtype=np.float64)
    N.check(lib.taxi_format_matrix(str(path).encode(), xids.bytes_ptr, xids.off_ptr, x0, nx, ny, _ptr(metrics), _ptr(undefined),
                                   int(column), float(scale), fmt.encode(), missing.encode(), threads))


def format_aligned_pairs(path, first_record: bool, xids: StringTable, yids: StringTable, x0, nx, ny,
                         aln_x: np.ndarray, aln_y: np.ndarray, aln_start: np.ndarray, aln_off: np.ndarray, threads=0) -> None:
    """Append the `SequencePairHandler.Formatted` records of a row block (pairs.py:81-97) from the
    raw arrays of `Engine.align_strings_raw`."""
    N.check(N.load().taxi_format_aligned_pairs(str(path).encode(), int(bool(first_record)), xids.bytes_ptr, xids.off_ptr,
                                               yids.bytes_ptr, yids.off_ptr, x0, nx, ny, _ptr(aln_x), _ptr(aln_y),
                                               _ptr(aln_start), _ptr(aln_off), threads))


def format_values(values: np.ndarray, undefined: np.ndarray | None, fmt: str, missing: str, scale: float = 1.0) -> list[str]:
    """[fmt % v or missing ...] for a vector of doubles, formatted by the library in one call."""
    values = np.ascontiguousarray(values, dtype=np.float64)
    n = len(values)
    if n == 0:
        return []
    und = None if undefined is None else np.ascontiguousarray(undefined, dtype=np.uint8)
    cap = 24 * n + 64
    while True:
        buf = C.create_string_buffer(cap)
        got = N.load().taxi_format_values(_ptr(values), n, _ptr(und), float(scale), fmt.encode(), missing.encode(), buf, cap)
        if got >= 0:
            return buf.raw[:got].decode("ascii").split("\t")
        if got > -16:
            N.check(int(got))
        cap = -got + 64


def _pointer_array(arrays):
    return (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])


def format_subset_rows(path, labels: StringTable, kx, ky, select, with_query: bool, mean, vmin, vmax, count, fmt: str, threads=0) -> None:
    """Append the rows of subsets/*/linear/{pairs,identity}.tsv for the selected keys (first-seen order):
    labels, then mean / min / max per metric (lists of per-metric arrays over the keys), "NA" where count is 0."""
    keep = [np.ascontiguousarray(a, dtype=np.float64) for group in (mean, vmin, vmax) for a in group]
    m = len(mean)
    cnt = [np.ascontiguousarray(a, dtype=np.int64) for a in count]
    select = np.ascontiguousarray(select, dtype=np.uint8)
    N.check(N.load().taxi_format_subset_rows(str(path).encode(), labels.bytes_ptr, labels.off_ptr, _ptr(kx), _ptr(ky), _ptr(select), len(kx),
                                             int(bool(with_query)), m, _pointer_array(keep[:m]), _pointer_array(keep[m:2 * m]),
                                             _pointer_array(keep[2 * m:]), _pointer_array(cnt), fmt.encode(), threads))


def format_subset_matrix(path, labels: StringTable, kx, ky, mean, vmin, vmax, count, fmt: str, pieces, threads=0) -> None:
    """Write one subsets/*/matricial/<metric>.tsv: runs of equal kx are its rows, `pieces` the four strings
    around {mean}, {min}, {max} in the statistics template."""
    arrays = [np.ascontiguousarray(a, dtype=np.float64) for a in (mean, vmin, vmax)]
    cnt = np.ascontiguousarray(count, dtype=np.int64)
    N.check(N.load().taxi_format_subset_matrix(str(path).encode(), labels.bytes_ptr, labels.off_ptr, _ptr(kx), _ptr(ky), len(kx),
                                               _ptr(arrays[0]), _ptr(arrays[1]), _ptr(arrays[2]), _ptr(cnt), fmt.encode(),
                                               *[piece.encode("utf-8") for piece in pieces], threads))


class NativeSubsetState:
    """sum / min / max / n / first-seen per (subset_x, subset_y) of one metric column."""

    def __init__(self, nsub: int):
        self.nsub = nsub
        self.sum = np.zeros(nsub * nsub, dtype=np.float64)
        self.min = np.full(nsub * nsub, inf, dtype=np.float64)
        self.max = np.zeros(nsub * nsub, dtype=np.float64)      # the reference starts max at 0.0
        self.count = np.zeros(nsub * nsub, dtype=np.int64)
        self.first_seen = np.full(nsub * nsub, -1, dtype=np.int64)
        self.next_order = np.zeros(1, dtype=np.int64)

    def add_block(self, metrics, undefined, x0, nx, ny, column, scale, xsubset, ysubset) -> None:
        metrics = np.ascontiguousarray(metrics, dtype=np.float64)
        N.check(N.load().taxi_aggregate_subsets(_ptr(metrics), _ptr(undefined), x0, nx, ny, int(column), float(scale), _ptr(xsubset),
                                                _ptr(ysubset), self.nsub, _ptr(self.sum), _ptr(self.min), _ptr(self.max),
                                                _ptr(self.count), _ptr(self.first_seen), _ptr(self.next_order)))

    def items(self):
        """((sx, sy), (min, max, mean, n)) in the order the keys first appeared."""
        keys = np.nonzero(self.first_seen >= 0)[0]
        for k in keys[np.argsort(self.first_seen[keys], kind="stable")]:
            n = int(self.count[k])
            stats = (None, None, None, 0) if not n else (float(self.min[k]), float(self.max[k]), float(self.sum[k]) / n, n)
            yield (int(k) // self.nsub, int(k) % self.nsub), stats
