"""B200-native normal-equations engine for the interior-point LP solvers of
pkhuong/cholesky-is-magic (affine-scaling, primal-dual-affine-scaling, newton-solve,
sparse-newton-solve).

Layout
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/nes.h) -> libnes.so
  nes.py           ctypes binding of the C ABI (what the Lisp does with sb-alien)
  sparse_cholesky.py, newton_solve.py, pdas.py, affine_scaling.py
                   host-side mirror of the reference's Lisp entry points (same names, argument
                   meaning and failure behaviour), calling only the C ABI
  standard_form.py, lpgen.py, read_mps.py
                   problem input: standard-form struct, synthetic LP generators, MPS reader

The directory name carries a hyphen (it is named after the reference), so import it through
``_pkg.load()`` at the repo root, which registers it as ``cholesky_is_magic_b200``.
"""
from . import nes  # noqa: F401
from .nes import Common, Factor, Matrix, NesError, load_library  # noqa: F401
