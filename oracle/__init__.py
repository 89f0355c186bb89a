"""CPU oracle: a NumPy/SciPy restatement of the reference's interior-point hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``cholesky-is-magic_b200/``) imports
this; it is used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` as the checker and the timed CPU baseline.

PARITY UNPINNED: the reference (pkhuong/cholesky-is-magic) ships no golden vectors, its tests are
unseeded randomized property checks (newton-solve.lisp:163-211, sparse-newton-solve.lisp:177-269),
and its arithmetic lives in two un-vendored, un-pinned third-party libraries that cannot be built
here:
  * SuiteSparse CHOLMOD (cholmod_analyze/factorize/solve/solve2/sdmult/scale; call sites
    sparse-cholesky.lisp:409-431, 461-473, 506-614; header "CHOLMOD/include/cholmod.h" at
    wrapper.c:1, dylib path at sparse-cholesky.lisp:1; no version anywhere in the tree),
  * matlisp (BLAS/LAPACK-backed real-matrix ops; call sites throughout newton-solve.lisp).
No Common Lisp implementation exists in this image either.  The oracle therefore restates the
Lisp control flow and algebra operation-for-operation and plays CHOLMOD/LAPACK's role with
OpenBLAS (dsyrk-like products + dpotrf/dpotrs through scipy.linalg).  It is pinned only by the
reference's own property tests, which tests/test_oracle.py re-runs with the reference's
generators and thresholds (four KKT residuals <= 1e-6), plus the literal IPM rules (init, step,
stop) transcribed from primal-dual-affine-scaling.lisp / affine-scaling.lisp, and -- independently of
this restatement -- by tests/test_known_answers.py: its PDAS and affine scaling must reach the optimum
that HiGHS (scipy.optimize.linprog) finds on the same LPs, within the reference's stop tolerance.
"""
