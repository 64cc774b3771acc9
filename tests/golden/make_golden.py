"""Extract the reference's own known-answer tables into JSON fixtures.

Run once in the build container (where /root/reference exists); the outputs are committed so
the GPU box (which has no /root/reference) can check the oracle and the CUDA path against them.

Sources (relative to /root/reference):
  tests/test_align.py:49-203          -> align_cases.json   (50 cases + 3 "Biopython-only" cases)
  tests/test_distances/metrics.tsv    -> metrics_cases.json (26 rows x 4 metrics, tol 5.1e-4)
  tests/test_distances.py:515-521     -> metrics_cases.json["exact"]
  tests/test_pairs/simple.*           -> pairs_simple.{tsv,formatted}
  samples/Taxi2test1_{10,50,120}.tab  -> sample inputs for parity runs
  tests/test_distances/*, tests/test_sequences/* (tsv, fas) -> distances/, sequences/ (handler fixtures)
  tests/test_statistics.py:113-250    -> statistics_cases.json (17 count + 108 statistic vectors)
  tests/test_statistics/*             -> statistics/ (writer fixtures)
  samples/Taxi2test1_ca200.tab        -> un-aligned 200-row resample for the align=False parity run
  tests/test_partitions/*.{tsv,fas}, tests/test_handlers/*.tsv -> partitions/, handlers/ (reader fixtures)

The reference test modules cannot be imported here (Bio / itaxotools.* are absent), so the
tables are read with `ast` instead of being executed.
"""
from __future__ import annotations

import ast
import json
import shutil
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).parent


def align_cases():
    tree = ast.parse((REF / "tests/test_align.py").read_text())
    default_scores = [1, -1, -8, -1, -1, -1]  # AlignTest.scores default, test_align.py:18
    out = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and node.targets[0].id in ("align_tests", "align_tests_failing"):
            cases = []
            for call in node.value.elts:
                args = [ast.literal_eval(a) for a in call.args]
                x, y = args[0]
                solutions = [list(s) for s in args[1]]
                scores = list(args[2]) if len(args) > 2 else default_scores
                cases.append(dict(x=x, y=y, solutions=solutions, scores=scores))
            out[node.targets[0].id] = cases
    # test_align.py:170-202: the alignments Biopython prints for the three cases the Rust
    # aligner gets wrong (recorded there as comments), with Biopython's score.
    biopython = [
        dict(aligned=["ATATATATATA", "AT-------TA"], score=46),
        dict(aligned=["AAA---TTTAAA", "AAACCC---AAA"], score=4),
        dict(aligned=["ATCG", "-AT-"], score=0),
    ]
    for case, extra in zip(out["align_tests_failing"], biopython):
        case["biopython"] = extra
    return out


def metric_cases():
    rows = []
    lines = (REF / "tests/test_distances/metrics.tsv").read_text().splitlines()
    header = lines[0].split("\t")
    assert header == ["target", "query", "p", "p-gaps", "jc", "k2p"]
    for line in lines[1:]:
        if not line:
            continue
        t, q, *vals = line.split("\t")
        rows.append(dict(x=t, y=q, expected=[None if v == "NA" else float(v) for v in vals]))
    exact = [
        dict(metric="p", x="gg-ccnccta", y="ggaccaccaa", expected=1.0 / 8.0),
        dict(metric="p-gaps", x="gg-ccnccta", y="ggaccaccaa", expected=2.0 / 9.0),
        dict(metric="p", x="---", y="nnn", expected=None),
    ]
    return dict(tolerance=0.00051, labels=header[2:], rows=rows, exact=exact)


def handler_fixtures():
    """Data files of the reference's reader/writer tests for the formats either side of the path
    (tests/test_distances/*, tests/test_sequences/* for Tabfile and Fasta only)."""
    (OUT / "distances").mkdir(exist_ok=True)
    (OUT / "sequences").mkdir(exist_ok=True)
    for path in sorted((REF / "tests/test_distances").iterdir()):
        if path.name != "metrics.tsv":
            shutil.copyfile(path, OUT / "distances" / path.name)
    for name in ("simple.tsv", "headers.tsv", "empty.tsv", "empty", "simple.fas", "simple.multi.fas", "simple.width.fas",
                 "species.fas", "species.dot.fas", "alleles.concat.fas", "alleles.plain.fas", "alleles.species.fas"):
        shutil.copyfile(REF / "tests/test_sequences" / name, OUT / "sequences" / name)


def statistics_cases():
    """CountTest(...) / StatisticTest(...) rows of tests/test_statistics.py; the argument
    expressions ("A" * 100, list(map(lambda ...)), sqrt(8 / 3)) are evaluated, nothing else is."""
    from math import sqrt
    tree = ast.parse((REF / "tests/test_statistics.py").read_text())
    env = {"sqrt": sqrt, "list": list, "map": map, "range": range}
    value = lambda node: eval(compile(ast.Expression(node), "<case>", "eval"), {"__builtins__": {}}, env)  # noqa: E731
    out = {"counts": [], "statistics": []}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.args:
            if node.func.id == "CountTest":
                out["counts"].append({"count": value(node.args[0]), "fixed": value(node.args[1]), "sequence": value(node.args[2])})
            elif node.func.id == "StatisticTest" and isinstance(node.args[0], ast.Attribute):
                out["statistics"].append({"stat": node.args[0].attr, "fixed": value(node.args[1]), "sequences": value(node.args[2])})
    (OUT / "statistics").mkdir(exist_ok=True)
    for path in sorted((REF / "tests/test_statistics").iterdir()):
        shutil.copyfile(path, OUT / "statistics" / path.name)
    return out


def reader_fixtures():
    """Input-path fixtures: partition files (tabular + FASTA titles) and plain tab files."""
    for sub, pattern in (("partitions", ("*.tsv", "*.fas")), ("handlers", ("*.tsv",))):
        (OUT / sub).mkdir(exist_ok=True)
        for pat in pattern:
            for path in sorted((REF / "tests" / f"test_{sub}").glob(pat)):
                shutil.copyfile(path, OUT / sub / path.name)


def samples():
    for name in ("Taxi2test1_10", "Taxi2test1_50", "Taxi2test1_120", "Taxi2test1_ca200"):
        shutil.copyfile(REF / "samples" / f"{name}.tab", OUT / f"{name}.tab")


if __name__ == "__main__":
    (OUT / "align_cases.json").write_text(json.dumps(align_cases(), indent=1) + "\n")
    (OUT / "metrics_cases.json").write_text(json.dumps(metric_cases(), indent=1) + "\n")
    for name in ("simple.tsv", "simple.formatted"):
        shutil.copyfile(REF / "tests/test_pairs" / name, OUT / f"pairs_{name}")
    (OUT / "statistics_cases.json").write_text(json.dumps(statistics_cases(), indent=1) + "\n")
    samples()
    handler_fixtures()
    reader_fixtures()
    print("golden fixtures written to", OUT)
