"""Seeded synthetic inputs shared by tests and bench (COI-like barcodes)."""
from __future__ import annotations

import numpy as np

ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)
COMPOSITION = np.array([0.23, 0.33, 0.15, 0.29])  # A C G T, measured on Taxi2test1_120.tab


def _mutate(rng, seq: np.ndarray, sub: float, indel: float = 0.0, nfrac: float = 0.0) -> np.ndarray:
    seq = seq.copy()
    n = len(seq)
    hit = rng.random(n) < sub
    seq[hit] = ALPHA[rng.choice(4, size=int(hit.sum()), p=COMPOSITION)]
    if indel > 0:
        k = rng.binomial(n, indel)
        for _ in range(k):
            pos = int(rng.integers(0, len(seq)))
            if rng.random() < 0.5 and len(seq) > 1:
                seq = np.delete(seq, pos)
            else:
                seq = np.insert(seq, pos, ALPHA[rng.choice(4, p=COMPOSITION)])
    if nfrac > 0:
        seq[rng.random(len(seq)) < nfrac] = ord("N")
    return seq


def coi_like(n: int, length: int = 650, seed: int = 650, genera: int = 50, species: int = 20) -> list[bytes]:
    """Hierarchical barcode set: root -> genera (12 % subs) -> species (6 %) -> individuals
    (1 % subs, 2 % single-base indels, 0.02 % N).  Individuals are dealt round-robin so any
    prefix of the list mixes close and distant pairs."""
    rng = np.random.default_rng(seed)
    root = ALPHA[rng.choice(4, size=length, p=COMPOSITION)]
    gen = [_mutate(rng, root, 0.12) for _ in range(genera)]
    spp = [[_mutate(rng, g, 0.06) for _ in range(species)] for g in gen]
    out = []
    for k in range(n):
        g = k % genera
        s = (k // genera) % species
        out.append(_mutate(rng, spp[g][s], 0.01, 0.02, 0.0002).tobytes())
    return out


def random_pairs(rng, n: int, lo: int, hi: int, sub: float = 0.15, indel: float = 0.03, alphabet: bytes = b"ACGTN") -> tuple[list[bytes], list[bytes]]:
    """n related (x, y) pairs with ragged lengths in [lo, hi]."""
    al = np.frombuffer(alphabet, dtype=np.uint8)
    xs, ys = [], []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        x = al[rng.integers(0, len(al), size=L)]
        y = x.copy()
        hit = rng.random(L) < sub
        y[hit] = al[rng.integers(0, len(al), size=int(hit.sum()))]
        k = rng.binomial(L, indel)
        for _ in range(k):
            pos = int(rng.integers(0, max(len(y), 1)))
            if rng.random() < 0.5 and len(y) > 1:
                run = int(rng.integers(1, 4))
                y = np.delete(y, slice(pos, pos + run))
            else:
                run = int(rng.integers(1, 4))
                y = np.insert(y, pos, al[rng.integers(0, len(al), size=run)])
        if len(y) == 0:
            y = x[:1].copy()
        xs.append(x.tobytes())
        ys.append(y.tobytes())
    return xs, ys


def read_tab_sequences(path, normalize: bool = True) -> tuple[list[str], list[str]]:
    """ids and sequences (normalized unless told otherwise) of a TaxI2 sample .tab (seqid ... sequence)."""
    ids, seqs = [], []
    with open(path, "r", encoding="utf-8", errors="surrogateescape") as f:
        header = f.readline().rstrip("\n").split("\t")
        ci, cs = header.index("seqid"), header.index("sequence")
        for line in f:
            line = line[:-1] if line.endswith("\n") else line
            if not line:
                continue
            row = line.split("\t")
            ids.append(row[ci])
            seqs.append(row[cs].replace("?", "N").replace("-", "").upper() if normalize else row[cs])
    return ids, seqs
