import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import _pkg  # noqa: E402

_pkg.load()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture
def common():
    """with-cholmod extent for one test (sparse-cholesky.lisp:400-406)."""
    from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
    with with_cholmod() as c:
        yield c
        # leak check of sparse-newton-solve.lisp:255-258 on every GPU test
        c.free_work()
        assert c.malloc_count == 0, f"leaked {c.malloc_count} blocks"
        assert c.memory_inuse == 0, f"leaked {c.memory_inuse} bytes"
