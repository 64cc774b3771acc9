"""A stand-in for taxi2_b200.engine.Engine whose results come from the CPU oracle, so the HOST
logic around the device (tile scheduling, greedy replays, block iteration) can be tested without a
GPU.  TEST INFRASTRUCTURE ONLY: nothing in the product imports this."""
from __future__ import annotations

import numpy as np

import oracle
from taxi2_b200.engine import pack_strings
from taxi2_b200.multi import MultiEngine


class OracleEngine:
    def __init__(self, device: int = 0, scores=None):
        self.device = device
        self.scores = scores
        self.n = [0, 0]
        self.sets = [None, None]
        self.calls = 0

    def close(self):
        pass

    def set_scores(self, scores):
        self.scores = None if scores is None else [int(v) for v in scores]

    def set_option(self, key, value):
        pass

    def load(self, seqs, which=0):
        data, off = seqs if isinstance(seqs, tuple) else pack_strings(seqs)
        self.sets[which] = (np.array(data), np.array(off))
        self.n[which] = len(off) - 1
        if which == 0:
            self.sets[1] = None
            self.n[1] = 0

    @property
    def ny(self):
        return self.n[1] if self.sets[1] is not None else self.n[0]

    def _joined(self):
        xd, xo = self.sets[0]
        yd, yo = self.sets[1] if self.sets[1] is not None else self.sets[0]
        return np.concatenate([xd, yd]), np.concatenate([xo, yo[1:] + xo[-1]]), len(xo) - 1

    def _pairs(self, px, py, align):
        data, off, nx = self._joined()
        px = np.asarray(px, dtype=np.int32)
        py = np.asarray(py, dtype=np.int32) + nx
        self.calls += 1
        if align:
            return oracle.align_count_pairs(data, off, px, py, self.scores)
        return oracle.count_pairs(data, off, px, py)

    def _rect(self, x0, nx, y0, ny, align, want):
        px, py = np.divmod(np.arange(nx * ny), max(ny, 1))
        res = self._pairs(px + x0, py + y0, align)
        out = {}
        if "score" in want and align:
            out["score"] = res["score"].reshape(nx, ny)
        if "counts" in want:
            out["counts"] = res["counts"].reshape(nx, ny, 4)
        if "metrics" in want:
            out["metrics"] = res["metrics"].reshape(nx, ny, 4)
        return out

    def align_rect(self, x0, nx, y0, ny, want=("score", "counts", "metrics"), **kw):
        return self._rect(x0, nx, y0, ny, True, want)

    def align_rect_both(self, x0, nx, y0, ny, want=("score", "counts", "metrics"), out=None, out_t=None):
        """Both orientations, each aligned on its own by the oracle (set 0 x set 0 only)."""
        if self.sets[1] is not None:
            raise NotImplementedError("stand-in: both orientations of one set only")
        xy = self._rect(x0, nx, y0, ny, True, want)
        yx = self._rect(y0, ny, x0, nx, True, want)
        self.last_redo = 0
        for res, given in ((xy, out), (yx, out_t)):
            if given is not None:
                for key in want:
                    given[key][...] = res[key]
        return xy, yx

    def count_rect(self, x0, nx, y0, ny, want=("counts", "metrics"), **kw):
        return self._rect(x0, nx, y0, ny, False, want)

    def align_pairs(self, px, py, want=("score", "counts", "metrics")):
        res = self._pairs(px, py, True)
        return {k: res[k] for k in want}

    def count_pairs(self, px, py, want=("counts", "metrics")):
        res = self._pairs(px, py, False)
        return {k: res[k] for k in want}

    def align_strings(self, px, py):
        xd, xo = self.sets[0]
        yd, yo = self.sets[1] if self.sets[1] is not None else self.sets[0]
        ax, ay, sc = [], [], []
        for i, j in zip(px, py):
            a, b, s = oracle.align(xd[xo[i]:xo[i + 1]].tobytes(), yd[yo[j]:yo[j + 1]].tobytes(), self.scores)
            ax.append(a.encode("latin-1")); ay.append(b.encode("latin-1")); sc.append(int(round(s)))
        return ax, ay, np.array(sc, dtype=np.int32)

    def align_strings_raw(self, px, py, want=("score",), slot=None):
        xd, xo = self.sets[0]
        yd, yo = self.sets[1] if self.sets[1] is not None else self.sets[0]
        ax, ay, score = self.align_strings(px, py)
        n = len(ax)
        off = np.zeros(n + 1, dtype=np.int64)
        for k, (i, j) in enumerate(zip(px, py)):
            off[k + 1] = off[k] + (xo[i + 1] - xo[i]) + (yo[j + 1] - yo[j])
        ox = np.zeros(max(int(off[-1]), 1), dtype=np.uint8)
        oy = np.zeros(max(int(off[-1]), 1), dtype=np.uint8)
        start = np.zeros(max(n, 1), dtype=np.int64)
        for k in range(n):
            start[k] = off[k + 1] - len(ax[k])
            ox[start[k]:off[k + 1]] = np.frombuffer(ax[k], dtype=np.uint8)
            oy[start[k]:off[k + 1]] = np.frombuffer(ay[k], dtype=np.uint8)
        if "counts" not in want and "metrics" not in want:
            return ox, oy, start, off, score
        res = self._pairs(px, py, True)
        return ox, oy, start, off, score, {k: res[k] for k in ("score", "counts", "metrics") if k in want or k == "score"}

    def stats(self):
        return dict(launches=self.calls, cells=0, kernel_ms=0.0)


def oracle_multi(ngpus: int = 2) -> MultiEngine:
    m = MultiEngine.__new__(MultiEngine)
    m.devices = list(range(ngpus))
    m.engines = [OracleEngine(d) for d in range(ngpus)]
    m.lens = [None, None]
    return m
