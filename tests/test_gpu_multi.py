"""Multi-GPU product path on real devices (taxi2_b200/multi.py): N contexts driven by N host
threads in one process give the SAME BITS as one context -- pair matrices, alignment-free
matrices, best matches (row tiles and column tiles combined on the host), and whole task output
trees.  On a one-GPU box the "several GPUs" are several contexts on device 0, which exercises the
same threads, plan and host gather; with more devices visible every device takes part."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN
from synth import coi_like, random_pairs

pytestmark = pytest.mark.gpu


def device_sets():
    from taxi2_b200 import _native

    count = int(_native.load().taxi_device_count())
    sets = [[0, 0], [0, 0, 0]]
    if count >= 2:
        sets.append(list(range(count)))
    return sets


@pytest.fixture(scope="module")
def single():
    from taxi2_b200.engine import Engine

    eng = Engine(0)
    yield eng
    eng.close()


def mixed_sequences():
    rng = np.random.default_rng(2024)
    seqs = coi_like(150, seed=77)
    xs, _ = random_pairs(rng, 40, 300, 1500, sub=0.1, indel=0.02)     # C5-like lengths: rows split by geometry
    return seqs + xs


@pytest.mark.parametrize("devices", device_sets(), ids=lambda d: f"gpus{len(d)}x{len(set(d))}")
def test_sharded_matrices_equal_one_gpu_bit_for_bit(single, devices):
    from taxi2_b200.multi import MultiEngine

    seqs = mixed_sequences()
    refs = coi_like(93, seed=5)
    single.set_scores(None)
    single.load(seqs, 0)
    single.load(refs, 1)
    want = single.align_rect(0, len(seqs), 0, len(refs))
    want_free = single.count_rect(0, len(seqs), 0, len(refs))
    with MultiEngine(devices) as multi:
        multi.load(seqs, 0)
        multi.load(refs, 1)
        for pinned, rows in ((False, 11), (True, None), (False, 1)):
            got = multi.align_matrix(rows_per_tile=rows, pinned=pinned)
            assert got["tiles"] >= min(len(seqs), 4)
            assert np.array_equal(got["score"], want["score"])
            assert np.array_equal(got["counts"], want["counts"])
            assert np.array_equal(got["metrics"], want["metrics"], equal_nan=True)
            assert got["cells"] == int(sum(map(len, seqs))) * int(sum(map(len, refs)))
        free = multi.count_matrix(rows_per_tile=17)
        assert np.array_equal(free["counts"], want_free["counts"])
        assert np.array_equal(free["metrics"], want_free["metrics"], equal_nan=True)
        part = multi.align_matrix(want=("counts",), rows_per_tile=5, x_range=(23, 61))
        assert np.array_equal(part["counts"], want["counts"][23:61])


@pytest.mark.parametrize("devices", device_sets(), ids=lambda d: f"gpus{len(d)}x{len(set(d))}")
@pytest.mark.parametrize("align", [True, False])
def test_best_matches_rows_and_column_tiles(single, devices, align):
    """versus_reference.py:184-188 on N GPUs: per query the first minimum over the references,
    whether a query's references sit on one GPU (row tiles) or are split (column tiles combined
    on the host).  Duplicated references force ties that only the first index may win."""
    from taxi2_b200.multi import MultiEngine

    queries = coi_like(70, seed=9)
    refs = coi_like(40, seed=10)
    refs = refs + refs[:15] + [queries[3], queries[3]]          # exact ties, incl. distance 0 twice
    if not align:
        queries = [q[:600] for q in queries]
        refs = [r[:600] for r in refs]
    single.set_scores(None)
    single.load(queries, 0)
    single.load(refs, 1)
    full = (single.align_rect if align else single.count_rect)(0, len(queries), 0, len(refs), want=("counts", "metrics"))
    for metric in (0, 3):
        col = full["metrics"][..., metric]
        key = np.where(np.isnan(col), np.inf, col)
        want_idx = key.argmin(axis=1)                              # numpy argmin = first minimum
        assert (np.isfinite(key.min(axis=1))).all()
        one = single.best_rows(0, len(queries), 0, len(refs), metric, align)
        assert np.array_equal(one["index"], want_idx)
        with MultiEngine(devices) as multi:
            multi.load(queries, 0)
            multi.load(refs, 1)
            for rows, col_tiles in ((None, 1), (9, 1), (9, 3), (70, 5)):
                got = multi.best_matches(metric, align, rows_per_tile=rows, col_tiles=col_tiles)
                assert np.array_equal(got["index"], want_idx), (metric, rows, col_tiles)
                assert np.array_equal(got["metrics"], full["metrics"][np.arange(len(queries)), want_idx], equal_nan=True)
                assert np.array_equal(got["counts"], full["counts"][np.arange(len(queries)), want_idx])


@pytest.mark.parametrize("devices", device_sets(), ids=lambda d: f"gpus{len(d)}x{len(set(d))}")
def test_symmetric_matrix_equals_all_ordered_pairs(single, devices):
    """versus_all.py:746 aligns (x, y) and (y, x).  align_matrix_symmetric aligns each unordered pair
    once, mirrors the result and re-aligns the orientation-sensitive pairs: the full n x n matrices
    must equal those of aligning every ordered pair, bit for bit -- barcodes (packed kernel with the
    tie bit), block sizes that exercise the diagonal recursion and ragged edges, and a mixed-length
    set that falls back to two launches per tile."""
    from taxi2_b200.multi import MultiEngine

    for name, seqs, blocks in (("coi", coi_like(700, seed=21), ((500, None), (128, 256), (None, None))),
                               ("mixed", mixed_sequences(), ((64, None),))):
        single.set_scores(None)
        single.load(seqs, 0)
        want = single.align_rect(0, len(seqs), 0, len(seqs))
        with MultiEngine(devices) as multi:
            multi.load(seqs, 0)
            for block, max_cols in blocks:
                got = multi.align_matrix_symmetric(block=block, max_cols=max_cols, pinned=(block == 500))
                for key in ("score", "counts"):
                    assert np.array_equal(got[key], want[key]), (name, block, key)
                assert np.array_equal(got["metrics"], want["metrics"], equal_nan=True), (name, block)
                if name == "coi":
                    # (pairs inside the small squares on the diagonal are aligned both ways and never counted as re-aligned)
                    assert 0 < got["redo"] < 0.03 * len(seqs) ** 2 and got["cells"] < 0.7 * 650 * 650 * len(seqs) ** 2
            rows = [(x0, nx) for x0, nx, _ in multi.iter_symmetric_rows(("metrics",), block=150)]
            assert rows == [(x0, min(150, len(seqs) - x0)) for x0 in range(0, len(seqs), 150)]


def test_best_rows_with_undefined_rows(single):
    single.load(["ACGTACGT", "NNNNNNNN", "ACGAACGT"], 0)
    single.load(["NNNNNNNN", "ACGTACGA", "ACGTACGT", "ACGTACGT"], 1)
    got = single.best_rows(0, 3, 0, 4, 0, align=False)
    assert list(got["index"]) == [2, -1, 2]     # one mismatch against reference 2, two against reference 1
    assert np.isnan(got["metrics"][1]).all() and got["metrics"][0, 0] == 0.0
    assert not got["counts"][1].any()


def _tree(path: Path) -> dict:
    return {str(p.relative_to(path)): p.read_bytes() for p in sorted(path.rglob("*")) if p.is_file()}


@pytest.mark.parametrize("devices", device_sets()[-1:], ids=lambda d: f"gpus{len(d)}x{len(set(d))}")
def test_tasks_on_several_gpus_write_the_same_bytes(tmp_path, devices, monkeypatch):
    from test_gpu_tasks import SILENT, load
    from taxi2_b200.tasks import VersusAll, VersusReference, common

    monkeypatch.setattr(common, "MAX_BLOCK_PAIRS", 300)   # six rows per block: every GPU gets several blocks

    seqs, species, genera = load("Taxi2test1_50.tab")
    trees = {}
    for name, devs in (("one", None), ("many", devices)):
        task = VersusAll()
        task.work_dir = tmp_path / name / "all"
        task.progress_handler = SILENT
        task.devices = devs
        task.input.sequences = seqs
        task.input.species, task.input.genera = species, genera
        task.start()
        records = list(seqs)
        ref = VersusReference()
        ref.work_dir = tmp_path / name / "ref"
        ref.progress_handler = SILENT
        ref.devices = devs
        from taxi2_b200.sequences import Sequences
        ref.input.data, ref.input.reference = Sequences(records[:20]), Sequences(records[20:])
        ref.start()
        trees[name] = _tree(tmp_path / name)
    assert trees["one"].keys() == trees["many"].keys() and len(trees["one"]) > 8
    for rel in trees["one"]:
        assert trees["one"][rel] == trees["many"][rel], rel
