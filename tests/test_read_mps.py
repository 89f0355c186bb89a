"""read-mps / to-standard-form restatement (read-mps.lisp:272-326, standard-form.lisp:18-105)."""
import io
import math

import numpy as np
import pytest

from cholesky_is_magic_b200 import lpgen, read_mps

SAMPLE = """NAME test lp
OBJSENSE
 MAX
ROWS
 N COST
 N IGNORED
 E R1
 G R2
 L R3
 L R4
COLUMNS
 X COST 1 R1 2.0
 X R2 1d0
 Y COST 2.5 R3 1
 Y IGNORED 7
 Z R4 1e0 R1 -1
RHS
 RHS R1 4 R2 1
 RHS R3 10
 RHS COST 99
RANGES
 RNG R4 3
BOUNDS
 UP BND X 4
 FR BND Y
 MI BND Z
ENDATA
"""


def test_reader_and_standard_form():
    mps = read_mps.read_mps(io.StringIO(SAMPLE))
    assert mps.name == "test lp" and mps.sense == "max"
    assert [r.name for r in mps.row_data] == ["R1", "R2", "R3", "R4"]
    assert mps.obj_row == [(0, 1.0), (1, 2.5)]
    sf = read_mps.to_standard_form(mps)
    # rows: R1 equality, R2 >= (slack -1), R3 <= (slack +1), R4 ranged [-3, 0] (slack in [0,3], b = ub)
    assert sf.ncons == 4 and sf.initial_vars == 3 and sf.nvars == 6
    np.testing.assert_allclose(sf.b, [4.0, 1.0, 10.0, 0.0])
    assert sf.type == [None, ">", "<", None]
    np.testing.assert_allclose(sf.l, [0, -math.inf, -math.inf, 0, 0, 0])
    np.testing.assert_allclose(sf.u, [4, math.inf, 0, math.inf, math.inf, 3])
    assert sf.c == [(0, -1.0), (1, -2.5)]                 # negated for max
    A = np.zeros((4, 6))
    A[sf.A.row, sf.A.col] = sf.A.value
    want = np.array([[2, 0, -1, 0, 0, 0], [1, 0, 0, -1, 0, 0], [0, 1, 0, 0, 1, 0], [0, 0, 1, 0, 0, 1.0]])
    np.testing.assert_allclose(A, want)


@pytest.mark.parametrize("bad", ["ROWS\n", "NAME\nROWS\n", "NAME x\nROWS\n E R\nCOLUMNS\n X R 1\nENDATA\n",
                                 "NAME x\nROWS\n E R\nCOLUMNS\n X Q 1\nRHS\nENDATA\n"])
def test_reader_rejects_what_the_lisp_asserts_on(bad):
    with pytest.raises(read_mps.MpsError):
        read_mps.read_mps(io.StringIO(bad))


def test_write_then_read_round_trip(tmp_path):
    sf = lpgen.sparse_lp(40, 90, nnz_per_col=4, bandwidth=10, seed=3)
    sf.u[::7] = 12.5
    sf.l[::5] = -1.0
    p = tmp_path / "lp.mps"
    read_mps.write_mps(p, sf)
    sf2 = read_mps.to_standard_form(read_mps.read_mps_file(p))
    assert (sf2.nvars, sf2.ncons) == (sf.nvars, sf.ncons)
    np.testing.assert_array_equal(sf2.b, sf.b)
    np.testing.assert_array_equal(sf2.l, sf.l)
    np.testing.assert_array_equal(sf2.u, sf.u)
    np.testing.assert_array_equal(sf2.c_dense(), sf.c_dense())
    A1 = np.zeros((40, 90)); A1[sf.A.row, sf.A.col] = sf.A.value
    A2 = np.zeros((40, 90)); A2[sf2.A.row, sf2.A.col] = sf2.A.value
    np.testing.assert_array_equal(A1, A2)
