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


def _c_prototypes():
    """name -> number of parameters, from include/nes.h (accessor macro expanded)."""
    text = open(os.path.join(ROOT, "include", "nes.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for name, args in re.findall(r"\b(nes_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text):
        args = args.strip()
        protos[name] = 0 if args in ("", "void") else args.count(",") + 1
    for field in re.findall(r"NES_DECLARE_ACCESSOR\((\w+),", text):
        if field != "FIELD":
            protos["nes_get_" + field] = 1
            protos["nes_set_" + field] = 2
    return protos


def test_lisp_glue_binds_only_exported_symbols_with_the_right_arity():
    """lisp/nes-glue.lisp cannot be executed here (no Common Lisp in the image), but every C name it binds with
    define-alien-routine must exist in libnes.so and take as many arguments as include/nes.h declares."""
    src = open(os.path.join(ROOT, "lisp", "nes-glue.lisp")).read()
    src = re.sub(r";[^\n]*", "", src)                         # strip comments
    lib = ctypes.CDLL(nes.LIB_PATH)
    protos = _c_prototypes()
    bound = {}
    for m in re.finditer(r'\(define-alien-routine\s+\("(nes_[a-z0-9_]+)"\s+[^)]+\)', src):
        # the form continues until its parentheses balance: return type, then one (name type) list per argument
        i, depth = m.start(), 0
        while True:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
            if depth == 0:
                break
        body = src[m.end():i - 1].strip()
        # drop the return type: either a symbol or one parenthesised form
        if body.startswith("("):
            d, j = 0, 0
            while True:
                d += {"(": 1, ")": -1}.get(body[j], 0)
                j += 1
                if d == 0:
                    break
            rest = body[j:]
        else:
            rest = body.split(None, 1)[1] if len(body.split(None, 1)) > 1 else ""
        nargs, d = 0, 0
        for ch in rest:
            if ch == "(":
                if d == 0:
                    nargs += 1
                d += 1
            elif ch == ")":
                d -= 1
        bound[m.group(1)] = nargs
    # the accessor macro binds nes_get_/nes_set_ for the 19 fields through format strings
    for field in re.findall(r'\(def #:[a-z-]+ "([a-z_]+)"', src):
        bound["nes_get_" + field] = 1
        bound["nes_set_" + field] = 2
    assert len(bound) >= 19 * 2 + 15
    for name, nargs in bound.items():
        assert hasattr(lib, name), f"{name} bound in nes-glue.lisp but not exported"
        assert name in protos, f"{name} not declared in include/nes.h"
        assert protos[name] == nargs, f"{name}: glue passes {nargs} arguments, nes.h declares {protos[name]}"
    # the names the other Lisp files of the reference call are all defined by the glue
    for lisp_name in ("with-cholmod", "make-sparse-from-triplet-vector", "scale-sparse!", "scale-sparse", "solve-sparse",
                      "solve-sparse-one-shot", "solve-sparse-recycle", "free-sparse-state", "solve-dense", "sparse-m*",
                      "cholmod-copy-sparse", "cholmod-free-sparse", "cholmod-analyze", "cholmod-free-work",
                      "cholmod-get-status", "cholmod-get-anz", "make-solve-sparse-state", "affine-A-copy", "flush"):
        target = lisp_name.replace("make-solve-sparse-state", "defstruct solve-sparse-state")
        assert (f"(defun {lisp_name} " in src or f"(defmacro {lisp_name} " in src or f" {lisp_name})" in src
                or target in src or "#:" + lisp_name.replace("cholmod-get-", "") in src), lisp_name
