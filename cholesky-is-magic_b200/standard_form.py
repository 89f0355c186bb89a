"""standard-form struct (standard-form.lisp:8-16):  min c'x  s.t.  Ax = b,  l <= x <= u.

Fields keep the Lisp names: ``c`` is a vector of (index . value) pairs sorted by index, ``A`` a
vector of triplets.  Triplets are stored as three parallel arrays (row, col, value) instead of a
vector of structs; ``A_dense`` optionally carries a dense m x n matrix for the dense configurations
(the Lisp `densify`, standard-form.lisp:137-157, builds the same thing from the triplets)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Triplets:
    row: np.ndarray
    col: np.ndarray
    value: np.ndarray

    def __len__(self):
        return len(self.value)


@dataclass
class StandardForm:
    nvars: int
    ncons: int
    c: list                      # [(index, value), ...] sorted by index
    A: Triplets | None
    b: np.ndarray
    l: np.ndarray
    u: np.ndarray
    type: list = field(default_factory=list)
    initial_vars: int | None = None
    A_dense: np.ndarray | None = None   # optional dense m x n view of A

    def c_dense(self):
        v = np.zeros(self.nvars)
        for i, x in self.c:
            v[i] = x
        return v


def rescale_sf(sf: StandardForm) -> StandardForm:
    """rescale-sf (standard-form.lisp:107-134): scale each row (and b) by 1/max|a_ij| when that
    maximum is >= 1e-6.  In place, like the Lisp."""
    norm = np.zeros(sf.ncons)
    if sf.A_dense is not None:
        norm = np.abs(sf.A_dense).max(axis=1)
    else:
        np.maximum.at(norm, sf.A.row, np.abs(sf.A.value))
    norm = np.where(norm < 1e-6, 1.0, 1.0 / np.where(norm < 1e-6, 1.0, norm))
    sf.b = sf.b * norm
    if sf.A_dense is not None:
        sf.A_dense = sf.A_dense * norm[:, None]
    else:
        sf.A.value = sf.A.value * norm[sf.A.row]
    return sf
