"""ctypes binding of libnes.so (include/nes.h) -- the same stubs a maintainer writes with sb-alien.

This mirrors sparse-cholesky.lisp:1-342 (the ``define-alien-routine`` block) one-to-one: each
``lib.nes_*`` prototype below corresponds to the CHOLMOD routine cited in include/nes.h.  The thin
classes (`Common`, `Matrix`, `Factor`) only manage handle lifetime the way the Lisp's
``with-cholmod`` / ``with-alien ... :local`` forms do.  No arithmetic happens in Python, and there is
no fallback: a missing library or GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnes.so")

NES_OK = 0
NES_NOT_POSDEF = 1
NES_DIV_BY_ZERO = 2
NES_MAXITER = 3
NES_ERR_NO_DEVICE = -1
NES_ERR_OUT_OF_MEMORY = -2
NES_ERR_INVALID = -4
NES_ERR_CUDA = -5
NES_ERR_COMM = -6

STAGE_FORM, STAGE_FACTOR, STAGE_SOLVE, STAGE_GEMV, STAGE_VECTOR = range(5)
STAGE_NAMES = ("form", "factor", "solve", "gemv", "vector")

_ACCESSORS = (
    ("print", C.c_int), ("print_function", C.c_void_p), ("dbound", C.c_double),
    ("supernodal_switch", C.c_double), ("supernodal", C.c_int), ("selected", C.c_int),
    ("itype", C.c_int), ("dtype", C.c_int), ("status", C.c_int), ("fl", C.c_double),
    ("lnz", C.c_double), ("anz", C.c_double), ("modfl", C.c_double),
    ("malloc_count", C.c_size_t), ("memory_usage", C.c_size_t), ("memory_inuse", C.c_size_t),
    ("rowfacfl", C.c_double), ("aatfl", C.c_double), ("blas_ok", C.c_int),
)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p
_vpp = C.POINTER(C.c_void_p)


class NesError(RuntimeError):
    pass


def _prototypes(lib):
    def fn(name, restype, *argtypes):
        f = getattr(lib, name)
        f.restype = restype
        f.argtypes = list(argtypes)

    fn("nes_allocate", _vp)
    fn("nes_release", None, _vp)
    for n in ("nes_start", "nes_finish", "nes_defaults", "nes_free_work", "nes_synchronize"):
        fn(n, C.c_int, _vp)
    fn("nes_set_device", C.c_int, _vp, C.c_int)
    fn("nes_last_error", C.c_char_p, _vp)
    fn("nes_version", C.c_int, _ip)
    for field, typ in _ACCESSORS:
        fn("nes_get_" + field, typ, _vp)
        fn("nes_set_" + field, typ, _vp, typ)
    fn("nes_get_minor", C.c_int, _vp)
    fn("nes_dense_to_matrix", _vp, _dp, C.c_size_t, C.c_size_t, C.c_size_t, _vp)
    fn("nes_triplet_to_sparse", _vp, _ip, _ip, _dp, C.c_size_t, C.c_size_t, C.c_size_t, _vp)
    fn("nes_csc_to_matrix", _vp, _ip, _ip, _dp, C.c_size_t, C.c_size_t, _vp)
    fn("nes_generate_dense", _vp, C.c_size_t, C.c_size_t, C.c_ulonglong, _vp)
    fn("nes_copy_matrix", _vp, _vp, _vp)
    fn("nes_free_matrix", C.c_int, _vpp, _vp)
    fn("nes_scale", C.c_int, _dp, C.c_int, _vp, _vp)
    fn("nes_unscale", C.c_int, _vp, _vp)
    for n in ("nes_matrix_nrow", "nes_matrix_ncol", "nes_matrix_nnz"):
        fn(n, C.c_size_t, _vp)
    fn("nes_matrix_is_dense", C.c_int, _vp)
    fn("nes_scale_rows_maxabs", C.c_int, _vp, _dp, _vp)
    fn("nes_sdmult", C.c_int, _vp, C.c_int, _dp, _dp, _dp, _dp, _vp)
    fn("nes_analyze", _vp, _vp, _vp)
    fn("nes_factorize", C.c_int, _vp, _vp, _vp)
    fn("nes_solve", C.c_int, C.c_int, _vp, _dp, _dp, _vp)
    fn("nes_solve2", C.c_int, C.c_int, _vp, _dp, _dp, _vp)
    fn("nes_free_factor", C.c_int, _vpp, _vp)
    fn("nes_solve_dense", C.c_int, _dp, C.c_size_t, C.c_size_t, _dp, _dp, _vp)
    fn("nes_factor_to_dense", C.c_int, _vp, _dp, C.c_size_t, _ip, _vp)
    fn("nes_normal_matrix_to_dense", C.c_int, _vp, _dp, C.c_size_t, _vp)
    fn("nes_factor_residual", C.c_int, _vp, _vp, _dp, _vp)
    fn("nes_kkt_newton", C.c_int, _vp, _vp, C.c_int, *([_dp] * 12), _vp)
    fn("nes_pdas_create", _vp, _vp, *([_dp] * 8), C.c_int, _vp)
    fn("nes_pdas_free", C.c_int, _vpp, _vp)
    fn("nes_pdas_violation", C.c_int, _vp, _dp, _vp)
    fn("nes_pdas_newton_direction", C.c_int, _vp, _dp, _vp)
    fn("nes_pdas_apply_step", C.c_int, _vp, C.c_double, _vp)
    fn("nes_pdas_repair", C.c_int, _vp, _dp, _vp)
    fn("nes_pdas_recentre", C.c_int, _vp, _dp, _vp)
    fn("nes_pdas_one_iteration", C.c_int, _vp, C.c_int, _dp, _vp)
    fn("nes_pdas_solve", C.c_int, _vp, C.c_int, _ip, _dp, _dp, _vp)
    fn("nes_pdas_get", C.c_int, _vp, C.c_int, _dp, _vp)
    fn("nes_pdas_set", C.c_int, _vp, C.c_int, _dp, _vp)
    fn("nes_affine_create", _vp, _vp, *([_dp] * 5), _vp)
    fn("nes_affine_free", C.c_int, _vpp, _vp)
    fn("nes_affine_residual", C.c_int, _vp, _dp, _vp)
    fn("nes_affine_repair", C.c_int, _vp, _dp, _vp)
    fn("nes_affine_direction", C.c_int, _vp, C.c_int, _dp, _vp)
    fn("nes_affine_apply", C.c_int, _vp, C.c_double, _vp)
    fn("nes_affine_one_iteration", C.c_int, _vp, C.c_int, _dp, _vp)
    fn("nes_affine_solve", C.c_int, _vp, C.c_int, _ip, _dp, _dp, _vp)
    fn("nes_affine_get", C.c_int, _vp, C.c_int, _dp, _vp)
    fn("nes_batch_create", _vp, _dp, C.c_int, C.c_int, C.c_int, *([_dp] * 5), _vp)
    fn("nes_batch_free", C.c_int, _vpp, _vp)
    fn("nes_batch_normal_solve", C.c_int, _vp, _dp, _dp, _dp, _ip, _vp)
    fn("nes_batch_affine_solve", C.c_int, _vp, C.c_int, _ip, _dp, _dp, _vp)
    fn("nes_batch_get_x", C.c_int, _vp, _dp, _vp)
    fn("nes_timing_enable", C.c_int, _vp, C.c_int)
    fn("nes_timing_reset", C.c_int, _vp)
    fn("nes_timing_get", C.c_int, _vp, C.c_int, _dp, C.POINTER(C.c_longlong))
    fn("nes_get_launch_count", C.c_longlong, _vp)
    fn("nes_get_form_flops", C.c_double, _vp)
    fn("nes_comm_unique_id", C.c_int, C.c_char_p)
    fn("nes_comm_init", C.c_int, _vp, C.c_int, C.c_int, C.c_char_p)
    fn("nes_comm_finalize", C.c_int, _vp)
    fn("nes_comm_rank", C.c_int, _vp)
    fn("nes_comm_nranks", C.c_int, _vp)
    fn("nes_dist_plan", C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, C.c_int)
    fn("nes_dist_plan_grid", C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, C.c_int, _ip, _ip)
    fn("nes_dist_plan_msgs", C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, C.c_int)
    fn("nes_dist_set_grid", C.c_int, _vp, C.c_int, C.c_int)
    fn("nes_dist_layout", C.c_int, _vp, C.c_int, _ip, _ip, _ip)
    fn("nes_mark_begin", C.c_int, _vp)
    fn("nes_mark_end", C.c_int, _vp, _dp)
    fn("nes_approx_create", _vp, _vp, _dp, _dp, _dp, _dp, _ip, _ip, _dp, _ip, C.c_int, _dp, C.c_double, C.c_int,
       C.c_double, _vp)
    fn("nes_approx_free", C.c_int, _vpp, _vp)
    fn("nes_approx_value_gradient", C.c_int, _vp, _dp, _dp, _dp, _dp, _vp)
    fn("nes_approx_get", C.c_int, _vp, C.c_int, _dp, _vp)
    fn("nes_approx_solve", C.c_int, _vp, C.c_int, _dp, _dp, _ip, _ip, _dp, _vp)
    fn("nes_approx_set_variant", C.c_int, _vp, C.c_int, C.c_double, _vp)
    fn("nes_approx_set_subproblem", C.c_int, _vp, C.c_double, _dp, C.c_double, _vp)
    fn("nes_symbolic_create", _vp, C.c_int, C.c_int, _ip, _ip, C.c_int, C.c_int, C.c_char_p, C.c_size_t)
    fn("nes_symbolic_ints", C.c_longlong, _vp, C.c_char_p, C.POINTER(_ip))
    fn("nes_symbolic_longs", C.c_longlong, _vp, C.c_char_p, C.POINTER(C.POINTER(C.c_longlong)))
    fn("nes_symbolic_scalar", C.c_double, _vp, C.c_char_p)
    fn("nes_symbolic_free", None, _vp)
    fn("nes_set_ordering_leaf", C.c_int, _vp, C.c_int)


_lib = None

_SYM_INTS = ("perm", "first", "nr", "ld", "rows", "rowptr", "sparent", "level", "lvlptr", "childptr", "child",
             "relptr", "rel", "cut", "ldu", "nb0", "tbptr", "tb", "owner", "ei", "ej", "segptr", "seg_tid", "inptr", "in_s")
_SYM_LONGS = ("off", "uoff", "edest", "vptr", "xu_off", "xv_off")
_SYM_SCALARS = ("anz", "aatfl", "lnz", "fl", "lsize", "usize", "nsuper", "nlevels")


def symbolic_analyze(colptr, rowidx, m, n, nranks=1, nd_leaf=0):
    """Host-only symbolic analysis (cholmod_analyze's role; csrc/sparse_symbolic.cu) as a dict of numpy
    arrays / scalars.  Needs no GPU."""
    lib = load_library()
    cp = np.ascontiguousarray(colptr, dtype=np.int32)
    ri = np.ascontiguousarray(rowidx, dtype=np.int32)
    err = C.create_string_buffer(512)
    h = lib.nes_symbolic_create(m, n, cp.ctypes.data_as(_ip), ri.ctypes.data_as(_ip), nranks, nd_leaf, err, 512)
    if not h:
        raise NesError("nes_symbolic_create failed: " + err.value.decode())
    out = {}
    try:
        for name in _SYM_INTS:
            p = _ip()
            cnt = lib.nes_symbolic_ints(h, name.encode(), C.byref(p))
            out[name] = np.ctypeslib.as_array(p, shape=(cnt,)).copy() if cnt > 0 else np.zeros(0, np.int32)
        for name in _SYM_LONGS:
            p = C.POINTER(C.c_longlong)()
            cnt = lib.nes_symbolic_longs(h, name.encode(), C.byref(p))
            out[name] = np.ctypeslib.as_array(p, shape=(cnt,)).copy() if cnt > 0 else np.zeros(0, np.int64)
        for name in _SYM_SCALARS:
            out[name] = lib.nes_symbolic_scalar(h, name.encode())
    finally:
        lib.nes_symbolic_free(h)
    return out


def load_library():
    """dlopen libnes.so.  Raises NesError (never falls back) when the extension is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NesError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (make -C cholesky-is-magic_b200/csrc).  There is no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _prototypes(_lib)
    return _lib


def unique_id() -> bytes:
    """ncclGetUniqueId through the library (call on rank 0, ship the 128 bytes to the other ranks)."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    if lib.nes_comm_unique_id(buf) != 0:
        raise NesError("nes_comm_unique_id failed (libnccl.so.2 not loadable)")
    return buf.raw


def dist_plan(m, nranks, rank):
    """Tiles of tril(M) owned by `rank` as an (ntiles, 2) int array (host-only, no GPU needed)."""
    lib = load_library()
    n = lib.nes_dist_plan(m, nranks, rank, None, None, 0)
    if n < 0:
        raise NesError("nes_dist_plan: bad arguments")
    rows = np.empty(max(n, 1), dtype=np.int32)
    cols = np.empty(max(n, 1), dtype=np.int32)
    lib.nes_dist_plan(m, nranks, rank, rows.ctypes.data_as(_ip), cols.ctypes.data_as(_ip), n)
    return np.stack([rows[:n], cols[:n]], axis=1)


def dist_plan_grid(m, P, Q, rank, nbo=0, chunk_rows=0):
    """(tiles (ntiles, 2), broadcasts per factorization, broadcasts rooted at `rank`) on a P x Q grid."""
    lib = load_library()
    nm, nr = C.c_int(0), C.c_int(0)
    n = lib.nes_dist_plan_grid(m, nbo, P, Q, rank, chunk_rows, None, None, 0, C.byref(nm), C.byref(nr))
    if n < 0:
        raise NesError("nes_dist_plan_grid: bad arguments")
    rows = np.empty(max(n, 1), dtype=np.int32)
    cols = np.empty(max(n, 1), dtype=np.int32)
    lib.nes_dist_plan_grid(m, nbo, P, Q, rank, chunk_rows, rows.ctypes.data_as(_ip), cols.ctypes.data_as(_ip), n,
                           C.byref(nm), C.byref(nr))
    return np.stack([rows[:n], cols[:n]], axis=1), nm.value, nr.value


def dist_plan_msgs(m, P, Q, nbo=0, chunk_rows=0, head_blocks=-1):
    """The broadcasts of a distributed factorization in order, as dicts (host-only planner)."""
    lib = load_library()
    n = lib.nes_dist_plan_msgs(m, nbo, P, Q, chunk_rows, head_blocks, None, 0)
    if n < 0:
        raise NesError("nes_dist_plan_msgs: bad arguments")
    buf = np.zeros(8 * max(n, 1), dtype=np.int32)
    lib.nes_dist_plan_msgs(m, nbo, P, Q, chunk_rows, head_blocks, buf.ctypes.data_as(_ip), n)
    keys = ("panel", "root", "has_diag", "row_start", "nblocks", "bh", "stride", "dep")
    return [dict(zip(keys, map(int, buf[8 * i: 8 * i + 8]))) for i in range(n)]


def vec(a):
    """Contiguous float64 view/copy + its ctypes pointer (make-dense, sparse-cholesky.lisp:346)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def ivec(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


class Common:
    """cholmod_common + with-cholmod (sparse-cholesky.lisp:344, 389-406)."""

    def __init__(self, device=None, timing=False):
        self.lib = load_library()
        self.ptr = self.lib.nes_allocate()
        if not self.ptr:
            raise NesError("nes_allocate failed")
        if device is not None:
            self.lib.nes_set_device(self.ptr, int(device))
        if not self.lib.nes_start(self.ptr):
            msg = self.error()
            self.lib.nes_release(self.ptr)
            self.ptr = None
            raise NesError(f"nes_start failed: {msg}")
        self.lib.nes_defaults(self.ptr)
        if timing:
            self.lib.nes_timing_enable(self.ptr, 1)

    def error(self):
        return (self.lib.nes_last_error(self.ptr) or b"").decode()

    def check(self, rc, what):
        if rc < 0:
            raise NesError(f"{what}: status {rc}: {self.error()}")
        return rc

    def close(self):
        if self.ptr:
            self.lib.nes_finish(self.ptr)
            self.lib.nes_release(self.ptr)
            self.ptr = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __getattr__(self, name):
        # cholmod-get-<field> accessors (sparse-cholesky.lisp:8-38)
        for field, _ in _ACCESSORS:
            if name == field:
                return getattr(self.lib, "nes_get_" + field)(self.ptr)
        raise AttributeError(name)

    def set(self, field, value):
        return getattr(self.lib, "nes_set_" + field)(self.ptr, value)

    @property
    def minor(self):
        return self.lib.nes_get_minor(self.ptr)

    def free_work(self):
        return self.lib.nes_free_work(self.ptr)

    def synchronize(self):
        return self.check(self.lib.nes_synchronize(self.ptr), "nes_synchronize")

    def timing_reset(self):
        self.lib.nes_timing_reset(self.ptr)

    def timing(self):
        out = {}
        for i, name in enumerate(STAGE_NAMES):
            ms, cnt = C.c_double(0), C.c_longlong(0)
            self.lib.nes_timing_get(self.ptr, i, C.byref(ms), C.byref(cnt))
            out[name] = (ms.value, cnt.value)
        return out

    def comm_init(self, nranks, rank, unique_id: bytes):
        """Join the NCCL communicator of the job (rank 0's id from `unique_id()`)."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self.check(self.lib.nes_comm_init(self.ptr, nranks, rank, buf), "nes_comm_init")

    def mark_begin(self):
        self.check(self.lib.nes_mark_begin(self.ptr), "nes_mark_begin")

    def mark_end(self):
        ms = C.c_double(0)
        self.check(self.lib.nes_mark_end(self.ptr, C.byref(ms)), "nes_mark_end")
        return ms.value

    @property
    def launches(self):
        return self.lib.nes_get_launch_count(self.ptr)

    @property
    def form_flops(self):
        """Algorithmic flops of the last formation launch timed as stage "form"."""
        return self.lib.nes_get_form_flops(self.ptr)


class Matrix:
    """A device-resident constraint matrix (the Lisp's (* cholmod-sparse) handle)."""

    def __init__(self, common, ptr):
        if not ptr:
            raise NesError(f"matrix creation failed: {common.error()}")
        self.common, self.ptr = common, ptr

    @classmethod
    def from_dense(cls, common, A):
        A = np.asarray(A, dtype=np.float64)
        m, n = A.shape
        Af = np.asfortranarray(A)
        return cls(common, common.lib.nes_dense_to_matrix(Af.ctypes.data_as(_dp), m, n, m, common.ptr))

    @classmethod
    def from_triplets(cls, common, rows, cols, vals, m, n):
        r, rp = ivec(rows)
        c_, cp = ivec(cols)
        v, vp_ = vec(vals)
        return cls(common, common.lib.nes_triplet_to_sparse(rp, cp, vp_, len(v), m, n, common.ptr))

    @classmethod
    def from_csc(cls, common, colptr, rowidx, vals, m, n):
        p, pp = ivec(colptr)
        i, ip_ = ivec(rowidx)
        v, vp_ = vec(vals)
        return cls(common, common.lib.nes_csc_to_matrix(pp, ip_, vp_, m, n, common.ptr))

    @classmethod
    def generate_dense(cls, common, m, n, seed):
        return cls(common, common.lib.nes_generate_dense(m, n, seed, common.ptr))

    def copy(self):
        return Matrix(self.common, self.common.lib.nes_copy_matrix(self.ptr, self.common.ptr))

    @property
    def shape(self):
        lib = self.common.lib
        return lib.nes_matrix_nrow(self.ptr), lib.nes_matrix_ncol(self.ptr)

    @property
    def nnz(self):
        return self.common.lib.nes_matrix_nnz(self.ptr)

    @property
    def is_dense(self):
        return bool(self.common.lib.nes_matrix_is_dense(self.ptr))

    def scale(self, s):
        s, sp = vec(s)
        assert len(s) == self.shape[1]
        ok = self.common.lib.nes_scale(sp, 2, self.ptr, self.common.ptr)
        if not ok:
            raise NesError(f"nes_scale failed: {self.common.error()}")
        return self

    def unscale(self):
        self.common.lib.nes_unscale(self.ptr, self.common.ptr)
        return self

    def scale_rows_maxabs(self):
        out = np.empty(self.shape[0])
        self.common.check(
            self.common.lib.nes_scale_rows_maxabs(self.ptr, out.ctypes.data_as(_dp), self.common.ptr),
            "nes_scale_rows_maxabs")
        return out

    def sdmult(self, x, transpose=False, y=None, alpha=1.0, beta=None):
        m, n = self.shape
        if beta is None:
            beta = 1.0 if y is not None else 0.0
        x, xp = vec(x)
        ny = n if transpose else m
        assert len(x) == (m if transpose else n)
        out = np.zeros(ny) if y is None else np.array(y, dtype=np.float64, copy=True)
        assert len(out) == ny
        a = (C.c_double * 2)(alpha, 0.0)
        b = (C.c_double * 2)(beta, 0.0)
        ok = self.common.lib.nes_sdmult(self.ptr, 1 if transpose else 0, a, b, xp,
                                        out.ctypes.data_as(_dp), self.common.ptr)
        if not ok:
            raise NesError(f"nes_sdmult failed: {self.common.error()}")
        return out

    def normal_matrix(self):
        m = self.shape[0]
        out = np.zeros((m, m), order="F")
        self.common.check(
            self.common.lib.nes_normal_matrix_to_dense(self.ptr, out.ctypes.data_as(_dp), m, self.common.ptr),
            "nes_normal_matrix_to_dense")
        return np.tril(out) + np.tril(out, -1).T

    def free(self):
        if self.ptr:
            h = C.c_void_p(self.ptr)
            ok = self.common.lib.nes_free_matrix(C.byref(h), self.common.ptr)
            assert ok != 0
            self.ptr = None


class Factor:
    """cholmod_factor handle (+ solve2 workspaces)."""

    def __init__(self, common, matrix):
        self.common = common
        self.ptr = common.lib.nes_analyze(matrix.ptr, common.ptr)
        if not self.ptr:
            raise NesError(f"nes_analyze failed: {common.error()}")
        self.m = matrix.shape[0]

    def factorize(self, matrix):
        """cholmod_set_status 0; cholmod_factorize; status check (sparse-cholesky.lisp:418-421).
        Returns True when the factorization succeeded, False on a non-positive pivot."""
        c = self.common
        c.set("status", 0)
        ok = c.lib.nes_factorize(matrix.ptr, self.ptr, c.ptr)
        if not ok:
            raise NesError(f"nes_factorize failed: {c.error()}")
        return c.status == 0

    def solve(self, b):
        b, bp = vec(b)
        x = np.empty(self.m)
        rc = self.common.lib.nes_solve(0, self.ptr, bp, x.ctypes.data_as(_dp), self.common.ptr)
        self.common.check(rc, "nes_solve")
        return x

    def residual(self, matrix):
        """||L L' - M||_F / ||M||_F over the whole matrix, computed on the device (nes_factor_residual)."""
        out = (C.c_double * 3)()
        self.common.check(self.common.lib.nes_factor_residual(matrix.ptr, self.ptr, out, self.common.ptr),
                          "nes_factor_residual")
        return out[0]

    def to_dense(self):
        out = np.zeros((self.m, self.m), order="F")
        self.common.check(
            self.common.lib.nes_factor_to_dense(self.ptr, out.ctypes.data_as(_dp), self.m, None,
                                                self.common.ptr), "nes_factor_to_dense")
        return out

    def free(self):
        if self.ptr:
            h = C.c_void_p(self.ptr)
            ok = self.common.lib.nes_free_factor(C.byref(h), self.common.ptr)
            assert ok != 0
            self.ptr = None
