"""Batched small dense LPs (BASELINE config 5): many independent `affine-scaling` runs
(affine-scaling.lisp:265-297) advanced together on the GPU through nes_batch_*."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import nes
from .affine_scaling import make_affine_state
from .sparse_cholesky import cholmod_common


class Batch:
    """B problems of identical shape m x n.  `A` has shape (B, m, n)."""

    def __init__(self, A, c=None, b=None, l=None, u=None, x=None):
        com = cholmod_common()
        A = np.asarray(A, dtype=np.float64)
        self.B, self.m, self.n = A.shape
        # per problem column-major, problems back to back
        A_all = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))
        ptr = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.float64).ctypes.data_as(nes._dp)
        keep = [None if v is None else np.ascontiguousarray(v, dtype=np.float64) for v in (c, b, l, u, x)]
        self.ptr = com.lib.nes_batch_create(A_all.ctypes.data_as(nes._dp), self.B, self.m, self.n,
                                            *[None if k is None else k.ctypes.data_as(nes._dp) for k in keep],
                                            com.ptr)
        if not self.ptr:
            raise nes.NesError(f"nes_batch_create failed: {com.error()}")

    @classmethod
    def from_standard_forms(cls, sfs):
        """One affine-scaling-state per standard form (make-affine-state, affine-scaling.lisp:52-90)."""
        sts = [make_affine_state(sf) for sf in sfs]
        A = np.stack([st.A_dense for st in sts])
        return cls(A, c=np.stack([st.c for st in sts]), b=np.stack([st.b for st in sts]),
                   l=np.stack([st.l for st in sts]), u=np.stack([st.u for st in sts]),
                   x=np.stack([st.x0 for st in sts]))

    def normal_solve(self, s, rhs):
        """x_b with (A_b diag s_b)(A_b diag s_b)' x_b = rhs_b; returns (x, status)."""
        com = cholmod_common()
        s_ = None if s is None else np.ascontiguousarray(s, dtype=np.float64)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        x = np.empty((self.B, self.m))
        status = np.zeros(self.B, dtype=np.int32)
        com.check(com.lib.nes_batch_normal_solve(self.ptr, None if s_ is None else s_.ctypes.data_as(nes._dp),
                                                 rhs.ctypes.data_as(nes._dp), x.ctypes.data_as(nes._dp),
                                                 status.ctypes.data_as(nes._ip), com.ptr), "nes_batch_normal_solve")
        return x, status

    def affine_scaling(self, max_iter=100000):
        """Returns (objectives, x, residual norms, iterations), each per problem."""
        com = cholmod_common()
        iters = np.zeros(self.B, dtype=np.int32)
        obj, res = np.empty(self.B), np.empty(self.B)
        com.check(com.lib.nes_batch_affine_solve(self.ptr, max_iter, iters.ctypes.data_as(nes._ip),
                                                 obj.ctypes.data_as(nes._dp), res.ctypes.data_as(nes._dp), com.ptr),
                  "nes_batch_affine_solve")
        x = np.empty((self.B, self.n))
        com.check(com.lib.nes_batch_get_x(self.ptr, x.ctypes.data_as(nes._dp), com.ptr), "nes_batch_get_x")
        return obj, x, res, iters

    def free(self):
        if self.ptr:
            com = cholmod_common()
            h = C.c_void_p(self.ptr)
            assert com.lib.nes_batch_free(C.byref(h), com.ptr) != 0
            self.ptr = None


# ---- multi-GPU: the batch index is split across ranks, no data-path collective ----
def shard_range(nbatch, world, rank):
    """Problems [lo, hi) of rank `rank`: contiguous, sizes differ by at most one."""
    base, extra = divmod(nbatch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedBatch:
    """The LPs of a batch that belong to this rank (one process per GPU).  Every rank builds it from the
    same list of standard forms; results stay local until `gather` is called (test / reporting only)."""

    def __init__(self, sfs, world, rank):
        self.nbatch, self.world, self.rank = len(sfs), world, rank
        self.lo, self.hi = shard_range(self.nbatch, world, rank)
        self.batch = Batch.from_standard_forms(sfs[self.lo:self.hi]) if self.hi > self.lo else None

    def affine_scaling(self, max_iter=100000):
        if self.batch is None:
            return np.empty(0), np.empty((0, 0)), np.empty(0), np.empty(0, dtype=np.int32)
        return self.batch.affine_scaling(max_iter)

    def free(self):
        if self.batch is not None:
            self.batch.free()
            self.batch = None


def gather(local, world):
    """Concatenate per-rank result arrays in rank order (torch.distributed must be initialised)."""
    if world == 1:
        return local
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, local)
    return tuple(np.concatenate([p[k] for p in parts if len(p[k])]) for k in range(len(local)))
