"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/nes.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from cholesky_is_magic_b200 import nes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nes.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(nes_[a-z0-9_]+)\s*\(", text))
    for field in re.findall(r"NES_DECLARE_ACCESSOR\((\w+),", text):
        if field != "FIELD":
            names.add("nes_get_" + field)
            names.add("nes_set_" + field)
    names.discard("nes_get_")
    names.discard("nes_set_")
    return sorted(names)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(nes.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 80
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/nes.h but not exported: {missing}"


def test_wrapper_c_accessor_fields_are_all_present():
    # the 19 fields of wrapper.c:31-52
    fields = ["print", "print_function", "dbound", "supernodal_switch", "supernodal", "selected",
              "itype", "dtype", "status", "fl", "lnz", "anz", "modfl", "malloc_count", "memory_usage",
              "memory_inuse", "rowfacfl", "aatfl", "blas_ok"]
    assert [f for f, _ in nes._ACCESSORS] == fields


def test_accessors_work_without_a_device():
    lib = nes.load_library()
    c = lib.nes_allocate()
    assert c
    assert lib.nes_set_status(c, 7) == 0      # set returns the old value (wrapper.c:24-29)
    assert lib.nes_get_status(c) == 7
    assert lib.nes_set_dbound(c, 1e-3) == 0.0
    assert lib.nes_get_dbound(c) == 1e-3
    assert lib.nes_defaults(c) == 1
    assert lib.nes_get_dbound(c) == 0.0       # cholmod_defaults: dbound = 0
    assert lib.nes_get_supernodal_switch(c) == 40.0
    lib.nes_release(c)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nes.NesError, match="no CPU fallback"):
        nes.Common()
    # un-started context: every compute entry point refuses
    lib = nes.load_library()
    c = lib.nes_allocate()
    assert not lib.nes_generate_dense(4, 4, 0, c)
    assert lib.nes_get_status(c) == nes.NES_ERR_NO_DEVICE
    lib.nes_release(c)
