"""Host-side mirror of the reference interface (no GPU): Type registry, containers, pair product,
labels, and the reader/writer formats either side of the hot path, against the reference's own
fixture files (tests/golden/{distances,sequences,pairs_*}, copied by make_golden.py)."""
from __future__ import annotations

import re
from pathlib import Path

import pytest

from conftest import GOLDEN
from taxi2_b200.align import PairwiseAligner, Scores
from taxi2_b200.distances import Distance, DistanceHandler, DistanceMetric, Distances
from taxi2_b200.handlers import FileHandler
from taxi2_b200.pairs import SequencePair, SequencePairHandler, SequencePairs
from taxi2_b200.sequences import Sequence, SequenceHandler, Sequences
from taxi2_b200.types import AttrDict, Container, Percentage, Type


def same_modulo_space(a: Path, b: Path) -> bool:
    strip = lambda p: re.sub(r"\s", "", p.read_text())  # noqa: E731
    return strip(a) == strip(b)


# ---- Type registry (reference tests/test_types.py:8-37) -------------------------------------------
def test_type_inheritance():
    class Parent(Type):
        pass

    class Child_A(Parent):
        pass

    class Child_B(Parent):
        pass

    class GrandChild_A(Child_A):
        pass

    class GrandChild_B(Child_A, Parent):
        pass

    assert Child_A in Parent and Child_B in Parent
    assert GrandChild_A in Child_A and GrandChild_A not in Parent
    assert GrandChild_B in Child_A and GrandChild_B in Parent
    assert Child_A() not in Parent
    with pytest.raises(TypeError):
        assert Child_A() not in Parent()
    with pytest.raises(TypeError):
        assert Child_A not in Parent()
    assert Parent.Child_A is Child_A and list(Parent) == [Child_A, Child_B, GrandChild_B]
    assert Child_A() == Child_A() and Child_A() != Child_B() and Child_A().type is Child_A


def test_registries_expose_reference_names():
    assert PairwiseAligner.Biopython in PairwiseAligner
    assert [str(m()) for m in (DistanceMetric.Uncorrected, DistanceMetric.UncorrectedWithGaps,
                               DistanceMetric.JukesCantor, DistanceMetric.Kimura2P)] == ["p", "p-gaps", "jc", "k2p"]
    assert DistanceHandler.Linear.WithExtras in DistanceHandler.Linear
    assert SequenceHandler.Tabfile in SequenceHandler.Tabular and FileHandler.Tabular.Tabfile in FileHandler.Tabular
    assert SequencePairHandler.Formatted in SequencePairHandler


def test_container_percentage_attrdict():
    c = Container(lambda n: iter(range(n)), 3)
    assert list(c) == [0, 1, 2] and list(c) == [0, 1, 2] and len(c) == 3
    with pytest.raises(TypeError):
        Container([1], 2)
    assert str(Percentage(0.1234)) == "12.34%"
    d = AttrDict(a=1)
    d.b = AttrDict(c=2)
    assert d.a == 1 and d["b"].c == 2
    with pytest.raises(AttributeError):
        d.missing


# ---- Scores / labels (align.py:17-35, tests/test_distances.py:503-512) ----------------------------
def test_scores_defaults_and_attribute_access():
    s = Scores()
    assert tuple(s.values()) == (1, -1, -8, -1, -1, -1)
    assert Scores(match_score=5).match_score == 5 and Scores(end_open_gap_score=0)["end_open_gap_score"] == 0
    s.mismatch_score = -3
    assert s["mismatch_score"] == -3


@pytest.mark.parametrize("metric,label", [
    (DistanceMetric.Uncorrected(), "p"), (DistanceMetric.UncorrectedWithGaps(), "p-gaps"),
    (DistanceMetric.JukesCantor(), "jc"), (DistanceMetric.Kimura2P(), "k2p"),
    (DistanceMetric.NCD(), "ncd"), (DistanceMetric.BBC(0), "bbc(0)"), (DistanceMetric.BBC(1), "bbc(1)"),
])
def test_labels(metric, label):
    assert metric == DistanceMetric.fromLabel(label)
    assert label == str(metric)


def test_normalize():
    assert Sequence("a", "ac-g?t-N").normalize().seq == "ACGNTN"
    seqs = Sequences([Sequence("a", "a-c"), Sequence("b", "??")]).normalize()
    assert [s.seq for s in seqs] == ["AC", "NN"] and [s.seq for s in seqs] == ["AC", "NN"]


# ---- pairs (tests/test_pairs.py) ------------------------------------------------------------------
def pairs_simple():
    return [SequencePair(Sequence("id1", "ATC-"), Sequence("id2", "ATG-")),
            SequencePair(Sequence("id1", "ATC-"), Sequence("id3", "-TAA")),
            SequencePair(Sequence("id2", "ATG-"), Sequence("id3", "-TAA"))]


def test_pairs_from_product_is_row_major_and_reiterable():
    xs = [Sequence("id1", "ATC"), Sequence("id2", "ATG")]
    ys = [Sequence("id3", "TAA"), Sequence("id4", "TAC"), Sequence("id5", "TAG")]
    ps = SequencePairs.fromProduct(Sequences(xs), Sequences(ys))
    want = [SequencePair(x, y) for x in xs for y in ys]
    assert list(ps) == want and list(ps) == want


@pytest.mark.parametrize("name,handler", [("pairs_simple.tsv", SequencePairHandler.Tabfile),
                                          ("pairs_simple.formatted", SequencePairHandler.Formatted)])
def test_pair_files(name, handler, tmp_path):
    assert list(SequencePairs.fromPath(GOLDEN / name, handler)) == pairs_simple()
    out = tmp_path / name
    with handler(out, "w") as file:
        for pair in pairs_simple():
            file.write(pair)
    assert out.read_text().replace("\n", "") == (GOLDEN / name).read_text().replace("\n", "")


# ---- distance files (tests/test_distances.py:371-500) ---------------------------------------------
P = DistanceMetric.Uncorrected


def d_simple():
    return [Distance(P(), Sequence("id1", None), Sequence(f"id{k}", None), d) for k, d in ((2, 0.1), (3, 0.2), (4, 0.3))]


def d_multiple():
    metrics = [DistanceMetric.Uncorrected(), DistanceMetric.UncorrectedWithGaps(), DistanceMetric.JukesCantor(),
               DistanceMetric.Kimura2P(), DistanceMetric.NCD(), DistanceMetric.BBC(0)]
    return [Distance(m, Sequence("id1", None), Sequence(f"id{y}", None), round(0.1 * (y - 1) + 0.01 * (k + 1), 2))
            for y in (2, 3, 4) for k, m in enumerate(metrics)]


def d_square(metric=None):
    vals = {(1, 1): 0.0, (1, 2): 0.1, (1, 3): 0.2, (2, 2): 0.0, (2, 3): 0.3, (3, 3): 0.0}
    return [Distance(metric or P(), Sequence(f"id{i}", None), Sequence(f"id{j}", None), vals[(min(i, j), max(i, j))])
            for i in (1, 2, 3) for j in (1, 2, 3)]


def d_rectangle():
    return [Distance(P(), Sequence(f"id{i}", None), Sequence(f"id{j}", None), round(0.1 * i + 0.01 * j, 2))
            for i in (1, 2, 3) for j in range(4, 10)]


def d_missing():
    return [Distance(P(), Sequence(f"id{i}", None), Sequence(f"id{j}", None), 0.0 if i == j else None)
            for i in (1, 2) for j in (1, 2)]


def d_extras():
    rows = [("query1", "K", "reference1", "X", "A", (0.11, 0.12, 0.13, 0.14)),
            ("query1", "K", "reference2", "Y", "B", (0.21, 0.22, 0.23, 0.24)),
            ("query2", "L", "reference3", "Z", "C", (0.31, 0.32, 0.33, None))]
    metrics = [DistanceMetric.Uncorrected(), DistanceMetric.UncorrectedWithGaps(), DistanceMetric.JukesCantor(), DistanceMetric.Kimura2P()]
    return [Distance(m, Sequence(qx, None, dict(voucher=vx)), Sequence(qy, None, dict(voucher=vy, organism=oy)), d)
            for qx, vx, qy, vy, oy, ds in rows for m, d in zip(metrics, ds)]


EXTRAS_KW = dict(idxHeader="seqid", idyHeader="id", tagX="_x", tagY="_y")

READ = [
    (d_simple, "simple.linear", DistanceHandler.Linear, {}),
    (d_multiple, "multiple.linear", DistanceHandler.Linear, {}),
    (d_missing, "missing.linear", DistanceHandler.Linear, {}),
    (list, "empty", DistanceHandler.Linear, {}),
    (lambda: d_square(DistanceMetric.Unknown()), "square.matrix", DistanceHandler.Matrix, {}),
    (list, "empty", DistanceHandler.Matrix, {}),
    (d_square, "square.matrix", DistanceHandler.Matrix, dict(metric=P())),
    (d_rectangle, "rectangle.matrix", DistanceHandler.Matrix, dict(metric=P())),
    (d_missing, "missing.matrix", DistanceHandler.Matrix, dict(metric=P())),
    (d_extras, "extras.tsv", DistanceHandler.Linear.WithExtras, EXTRAS_KW),
    (d_extras, "extras.tsv", DistanceHandler.Linear.WithExtras, dict(idxColumn=0, idyColumn=2, tagX="_x", tagY="_y")),
    (list, "empty", DistanceHandler.Linear.WithExtras, {}),
]

WRITE = [
    (d_simple, "simple.linear", DistanceHandler.Linear, dict(formatter="{:.1f}")),
    (d_multiple, "multiple.linear", DistanceHandler.Linear, dict(formatter="{:.2f}")),
    (d_missing, "missing.linear", DistanceHandler.Linear, dict(formatter="{:.1f}")),
    (list, "empty", DistanceHandler.Linear, dict(formatter="{:.1f}")),
    (d_square, "square.matrix", DistanceHandler.Matrix, dict(formatter="{:.1f}")),
    (d_rectangle, "rectangle.matrix", DistanceHandler.Matrix, dict(formatter="{:.2f}")),
    (d_missing, "missing.matrix", DistanceHandler.Matrix, dict(formatter="{:.1f}")),
    (list, "empty", DistanceHandler.Matrix, dict(formatter="{:.1f}")),
    (d_missing, "missing.formatted.linear", DistanceHandler.Linear, dict(formatter="{:.2e}", missing="nan")),
    (d_missing, "missing.formatted.matrix", DistanceHandler.Matrix, dict(formatter="{:.2e}", missing="nan")),
    (d_extras, "extras.tsv", DistanceHandler.Linear.WithExtras, dict(formatter="{:.2f}", **EXTRAS_KW)),
    (d_missing, "missing.formatted.linear", DistanceHandler.Linear.WithExtras,
     dict(idxHeader="idx", idyHeader="idy", tagX="", tagY="", formatter="{:.2e}", missing="nan")),
    (list, "empty", DistanceHandler.Linear.WithExtras, dict(formatter="{:.1f}")),
]


@pytest.mark.parametrize("fixture,name,handler,kwargs", READ)
def test_read_distances(fixture, name, handler, kwargs):
    got = list(Distances.fromPath(GOLDEN / "distances" / name, handler, **kwargs))
    want = fixture()
    assert len(got) == len(want)
    for d in want:
        assert d in got


@pytest.mark.parametrize("fixture,name,handler,kwargs", WRITE)
def test_write_distances(fixture, name, handler, kwargs, tmp_path):
    out = tmp_path / name
    with handler(out, "w", **kwargs) as file:
        for d in fixture():
            file.write(d)
    assert same_modulo_space(out, GOLDEN / "distances" / name)


# ---- sequence files (tests/test_sequences.py, Tabfile + Fasta only) -------------------------------
def s_simple():
    return [Sequence("id1", "ATC"), Sequence("id2", "ATG"), Sequence("id3", "ATA")]


def s_tagged(tag):
    return [Sequence(i, s, {tag: v}) for i, s, v in (("id1", "ATC", "X"), ("id2", "ATG", "Y"), ("id3", "ATA", "Z"))]


def s_alleles():
    return [Sequence(i, s, {"allele": a, "species": sp}) for i, s, sp in (("id1", "ATC", "X"), ("id2", "ATG", "Y"), ("id3", "ATA", "Z"))
            for a in "ab"]


SEQ_READ = [
    (s_simple, "simple.fas", SequenceHandler.Fasta, {}),
    (s_simple, "simple.multi.fas", SequenceHandler.Fasta, {}),
    (s_simple, "simple.tsv", SequenceHandler.Tabfile, {}),
    (lambda: s_tagged("voucher"), "headers.tsv", SequenceHandler.Tabfile, dict(idHeader="seqid", seqHeader="sequences")),
    (lambda: s_tagged("voucher"), "species.fas", SequenceHandler.Fasta, dict(parse_organism=True, organism_tag="voucher", organism_separator="|")),
    (lambda: s_tagged("organism"), "species.fas", SequenceHandler.Fasta, dict(parse_organism=True)),
    (lambda: s_tagged("organism"), "species.dot.fas", SequenceHandler.Fasta, dict(parse_organism=True, organism_separator=".")),
    (list, "empty", SequenceHandler.Fasta, {}),
    (list, "empty.tsv", SequenceHandler.Tabfile, dict(idHeader="seqid", seqHeader="sequences")),
]

SEQ_WRITE = [
    (s_simple, "simple.tsv", SequenceHandler.Tabfile, {}),
    (lambda: s_tagged("voucher"), "headers.tsv", SequenceHandler.Tabfile, dict(idHeader="seqid", seqHeader="sequences")),
    (s_simple, "simple.fas", SequenceHandler.Fasta, {}),
    (s_simple, "simple.width.fas", SequenceHandler.Fasta, dict(line_width=2)),
    (lambda: s_tagged("organism"), "species.fas", SequenceHandler.Fasta, dict(write_organism=True)),
    (lambda: s_tagged("organism"), "species.dot.fas", SequenceHandler.Fasta, dict(write_organism=True, organism_separator=".")),
    (lambda: s_tagged("voucher"), "species.fas", SequenceHandler.Fasta, dict(write_organism=True, organism_tag="voucher", organism_separator="|")),
    (s_alleles, "alleles.concat.fas", SequenceHandler.Fasta, dict(write_organism=False, concatenate_extras=["species", "allele"])),
    (s_alleles, "alleles.plain.fas", SequenceHandler.Fasta, dict(write_organism=False, concatenate_extras=["allele"])),
    (s_alleles, "alleles.species.fas", SequenceHandler.Fasta,
     dict(write_organism=True, organism_separator="|", organism_tag="species", concatenate_extras=["allele"])),
]


@pytest.mark.parametrize("fixture,name,handler,kwargs", SEQ_READ)
def test_read_sequences(fixture, name, handler, kwargs):
    got = list(Sequences.fromPath(GOLDEN / "sequences" / name, handler, **kwargs))
    want = fixture()
    assert len(got) == len(want)
    for s in want:
        assert s in got


@pytest.mark.parametrize("fixture,name,handler,kwargs", SEQ_WRITE)
def test_write_sequences(fixture, name, handler, kwargs, tmp_path):
    out = tmp_path / name
    with handler(out, "w", **kwargs) as file:
        for s in fixture():
            file.write(s)
    assert same_modulo_space(out, GOLDEN / "sequences" / name)


def test_sample_tab_reads_like_the_reference():
    seqs = list(Sequences.fromPath(GOLDEN / "Taxi2test1_10.tab", SequenceHandler.Tabfile, idHeader="seqid", seqHeader="sequence"))
    assert len(seqs) == 10 and seqs[0].id == "specimen1"
    assert set(seqs[0].extras) == {"specimen_voucher", "organism"}
    assert not seqs[0].seq.endswith("\r")  # CRLF sample, universal newlines
