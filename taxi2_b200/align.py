"""Pairwise aligner API: drop-in for /root/reference/src/itaxotools/taxi2/align.py.

`PairwiseAligner.Biopython(scores)` keeps the name every reference task uses
(versus_all.py:532, versus_reference.py:105, dereplicate.py:213, decontaminate.py:177,
decontaminate2.py:194) but runs on the GPU: global alignment with the six `Scores`, returning
the alignment Biopython's `aligner.align(x, y)[0]` returns, as two gapped strings.
There is no CPU fallback; without the CUDA library or a device the constructor raises.

The out-of-scope `PairwiseAligner.Rust` backend (flagged as sub-optimal by the reference itself,
align.py:57-60) is not provided.
"""
from __future__ import annotations

from itertools import islice
from typing import Iterator

import numpy as np

from .pairs import SequencePair, SequencePairs
from .sequences import Sequence
from .types import Type


class Scores(dict):
    """Alignment scores; keys are also attributes (align.py:17-35)."""

    defaults = dict(
        match_score=1,
        mismatch_score=-1,
        internal_open_gap_score=-8,
        internal_extend_gap_score=-1,
        end_open_gap_score=-1,
        end_extend_gap_score=-1,
    )

    def __init__(self, **kwargs):
        super().__init__(self.defaults | kwargs)
        self.__dict__ = self

    def __repr__(self):
        attrs = ", ".join(f"{k}={v}" for k, v in self.items())
        return f"<{type(self).__name__}: {attrs}>"


class PairwiseAligner(Type):
    def __init__(self, scores: Scores = None):
        self.scores = scores or Scores()

    def align(self, pair: SequencePair) -> SequencePair:
        raise NotImplementedError()

    def align_pairs_parallel(self, pairs: SequencePairs) -> Iterator[SequencePair]:
        # the reference's multiprocessing.Pool variant (align.py:45-48, unused by every task):
        # batching on the device already is the parallel form
        yield from self.align_pairs(pairs)

    def align_pairs(self, pairs: SequencePairs) -> SequencePairs:
        return SequencePairs((self.align(pair) for pair in pairs))


class Biopython(PairwiseAligner):
    """Biopython-compatible global aligner on sm_100a (the class name is the reference's)."""

    batch_size = 4096  # pairs aligned per device launch in align_pairs()

    def __init__(self, scores: Scores = None, device: int = 0):
        super().__init__(scores)
        from .engine import default_engine, scores_vector  # raises without library / GPU

        self._vector = scores_vector(dict(self.scores))
        self._engine = default_engine(device)

    # -- single pair (align.py:151-157) ---------------------------------------------------------
    def align(self, pair: SequencePair) -> SequencePair:
        return self._align_batch([pair])[0]

    # -- lazy, order-preserving, re-iterable stream (align.py:50-51) ------------------------------
    def align_pairs(self, pairs: SequencePairs) -> SequencePairs:
        def generate():
            it = iter(pairs)
            while True:
                chunk = list(islice(it, self.batch_size))
                if not chunk:
                    return
                yield from self._align_batch(chunk)

        return SequencePairs(generate)

    def _align_batch(self, chunk: list[SequencePair]) -> list[SequencePair]:
        eng = self._engine
        eng.set_scores(self._vector)
        # de-duplicate sequence strings inside the batch: a row-major product repeats them
        index: dict[str, int] = {}
        px = np.empty(len(chunk), dtype=np.int32)
        py = np.empty(len(chunk), dtype=np.int32)
        for k, pair in enumerate(chunk):
            px[k] = index.setdefault(pair.x.seq, len(index))
            py[k] = index.setdefault(pair.y.seq, len(index))
        eng.load(list(index), 0)
        ax, ay, _ = eng.align_strings(px, py)
        return [
            SequencePair(
                Sequence(pair.x.id, ax[k].decode("latin-1"), pair.x.extras),
                Sequence(pair.y.id, ay[k].decode("latin-1"), pair.y.extras),
            )
            for k, pair in enumerate(chunk)
        ]
