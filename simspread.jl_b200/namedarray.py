"""Minimal stand-in for NamedArrays.jl's NamedMatrix (the container the reference API speaks:
`.array`, `names(A, d)`, `setnames!`, indexing by name lists).  Host-side
bookkeeping only -- no numerics happen here."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


class NamedArray:
    def __init__(self, array, names=None):
        a = np.asarray(array)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        if a.ndim != 2:
            raise ValueError("NamedArray holds matrices")
        self.array = a
        if names is None:  # NamedArrays.jl default names are "1", "2", ...
            names = ([str(i + 1) for i in range(a.shape[0])], [str(j + 1) for j in range(a.shape[1])])
        r, c = [str(x) for x in names[0]], [str(x) for x in names[1]]
        if len(r) != a.shape[0] or len(c) != a.shape[1]:
            raise ValueError("names do not match the array shape")
        self._names: List[List[str]] = [r, c]

    # -- NamedArrays.jl API --------------------------------------------------------------------
    def names(self, d: int | None = None):
        """`names(A, d)` with Julia's 1-based dimension; `names(A)` returns both."""
        if d is None:
            return [list(self._names[0]), list(self._names[1])]
        return list(self._names[d - 1])

    def setnames(self, v: Sequence[str], d: int):
        """`setnames!(A, v, d)`."""
        v = [str(x) for x in v]
        if len(v) != self.array.shape[d - 1]:
            raise ValueError("wrong number of names")
        self._names[d - 1] = v

    @property
    def shape(self):
        return self.array.shape

    def size(self, d: int):
        return self.array.shape[d - 1]

    def copy(self) -> "NamedArray":
        return NamedArray(self.array.copy(), self.names())

    def _positions(self, d: int) -> dict:
        """name -> first position along dimension d; cached against the identity of the name list (setnames and
        every constructor install a fresh list, so a stale cache cannot be hit)."""
        names = self._names[d - 1]
        cache = self.__dict__.setdefault("_pos_cache", {})
        hit = cache.get(d)
        if hit is not None and hit[0] is names and hit[1] == len(names):
            return hit[2]
        pos = {}
        for i, n in enumerate(names):
            pos.setdefault(n, i)
        cache[d] = (names, len(names), pos)
        return pos

    def index_of(self, wanted: Sequence[str], d: int) -> np.ndarray:
        pos = self._positions(d)
        try:
            return np.array([pos[str(w)] for w in wanted], dtype=np.int32)
        except KeyError as e:
            raise KeyError(f"name {e.args[0]!r} not found along dimension {d}") from None

    def __getitem__(self, key) -> "NamedArray":
        """`A[rownames, colnames]`: gathers in the given order, returns a fresh NamedArray."""
        rk, ck = key
        ri = np.arange(self.shape[0]) if isinstance(rk, slice) else self.index_of(_aslist(rk), 1)
        ci = np.arange(self.shape[1]) if isinstance(ck, slice) else self.index_of(_aslist(ck), 2)
        return NamedArray(self.array[np.ix_(ri, ci)],
                          ([self._names[0][i] for i in ri], [self._names[1][j] for j in ci]))

    def __eq__(self, other):  # NamedArray == compares values only (test/runtests.jl:14)
        o = other.array if isinstance(other, NamedArray) else np.asarray(other)
        return self.array.shape == o.shape and bool(np.all(self.array == o))

    def __repr__(self):
        return f"NamedArray{self.array.shape}(rows={self._names[0][:3]}.., cols={self._names[1][:3]}..)"


def _aslist(x):
    return [x] if isinstance(x, str) else list(x)
