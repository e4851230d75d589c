"""CPU tests of the drop-in boundary: libwbeuler.so loads, exports every symbol include/wbeuler.h
declares, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "wbeuler.h")


def declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(wb_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    import wbeuler
    return wbeuler.lib()


def test_header_declares_the_reference_routines():
    syms = declared_symbols()
    for s in ("wb_fv2d_compute_update_exact", "wb_fv2d_compute_max_speed", "wb_fv2d_evolve", "wb_fv2d_create"):
        assert s in syms


def test_every_declared_symbol_is_exported(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/wbeuler.h but not exported: {missing}"


def test_bad_arguments_are_rejected_without_touching_the_gpu(lib):
    import wbeuler
    p = wbeuler.FV2DParams(64, 64, 3, 2, 1.4, 1.0, 1.0, 0.5, 0, -1, 0, 1)   # nvar = 3
    h = C.c_void_p()
    assert lib.wb_fv2d_create(C.byref(h), C.byref(p)) == -1
    assert b"nvar" in lib.wb_last_error()
    assert lib.wb_fv2d_create(None, None) == -1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    import wbeuler
    with pytest.raises(wbeuler.WBError) as e:
        wbeuler.FV2D(32, 32)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under fvm-source-wb_b200/ may reference it."""
    pkg = os.path.join(ROOT, "fvm-source-wb_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".f90")):
                if re.search(r"\boracle\b", open(os.path.join(d, f), errors="ignore").read()):
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def _c_struct_members(name):
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct \{((?:(?!typedef struct).)*?)\}\s*" + name + r"\s*;", txt, flags=re.S).group(1)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, names = decl.split(None, 1)
        out += [(ctype, n.strip()) for n in names.split(",")]
    return out


def _fortran_type_members(path, name):
    txt = re.sub(r"!.*", "", open(path).read())
    body = re.search(r"type,\s*bind\(C\)\s*::\s*" + name + r"(.*?)end type", txt, flags=re.S | re.I).group(1)
    out = []
    for line in body.strip().splitlines():
        m = re.match(r"\s*(integer\(c_int\)|real\(c_double\))\s*::\s*(.*)", line.strip(), flags=re.I)
        if m:
            ctype = "int" if "c_int" in m.group(1).lower() else "double"
            out += [(ctype, n.strip()) for n in m.group(2).split(",")]
    return out


SHIMS = [("wb_shim_2d.f90", "wb_fv2d_params", "FV2DParams"), ("wb_shim_dg2d.f90", "wb_dg2d_params", "DG2DParams"),
         ("wb_shim_fvm1d.f90", "wb_fvm1d_params", "FVM1DParams"), ("wb_shim_fv1d.f90", "wb_fv1d_params", "FV1DParams"),
         ("wb_shim_dg1d.f90", "wb_dg1d_params", "DG1DParams")]


@pytest.mark.parametrize("shim,struct,mirror", SHIMS)
def test_fortran_shims_mirror_the_c_structs_and_bind_exported_symbols(lib, shim, struct, mirror):
    """The ISO_C_BINDING shims cannot be compiled here (no Fortran compiler), so the parts that would fail silently are
    checked textually: the bind(C) derived type has the C struct's members in the same order with the same types, every
    bind(C, name=...) is a symbol the library exports, and the ctypes mirror used by the tests has the same layout."""
    path = os.path.join(ROOT, "fvm-source-wb_b200", "fortran", shim)
    c_members = _c_struct_members(struct)
    assert _fortran_type_members(path, struct) == c_members
    bound = re.findall(r'bind\(C,\s*name="(wb_[a-z0-9_]+)"\)', open(path).read())
    assert bound and all(hasattr(lib, s) for s in bound), [s for s in bound if not hasattr(lib, s)]
    import wbeuler
    ct = getattr(wbeuler, mirror)
    assert [(("int" if t is C.c_int else "double"), n) for n, t in ct._fields_] == c_members


def _c_prototypes():
    """name -> list of argument kinds of every `int wb_*(...)` / `const char* wb_*(...)` prototype of the header"""
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const\s+char\s*\*|long\s+long)\s+(wb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", txt):
        kinds = []
        for a in m.group(2).split(","):
            a = " ".join(a.split())
            if a in ("", "void"):
                continue
            if "**" in a:
                kinds.append("handle_out")
            elif re.search(r"wb_\w+_params\s*\*", a):
                kinds.append("struct_in")
            elif re.search(r"\bwb_\w+\s*\*", a):
                kinds.append("handle")
            elif re.search(r"\bdouble\s*\*", a):
                kinds.append("ptr_double")
            elif re.search(r"\bint\s*\*", a):
                kinds.append("ptr_int")
            elif re.search(r"\bvoid\s*\*", a):
                kinds.append("ptr_void")
            elif re.search(r"\bdouble\b", a):
                kinds.append("val_double")
            elif re.search(r"\bint\b", a):
                kinds.append("val_int")
            else:
                kinds.append("?" + a)
        out[m.group(1)] = kinds
    return out


def _fortran_interfaces(path):
    """name -> list of argument kinds of every bind(C) function interface of a shim (dummy order of the function statement)"""
    txt = re.sub(r"!.*", "", open(path).read())
    txt = re.sub(r"&\s*\n\s*", " ", txt)                       # continuation lines
    out = {}
    for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name=\"(\w+)\"\)(.*?)end function", txt, flags=re.S | re.I):
        dummies = [d.strip().lower() for d in m.group(2).split(",") if d.strip()]
        kind = {}
        for line in m.group(4).splitlines():
            d = re.match(r"\s*(type\(c_ptr\)|type\(wb_\w+\)|real\(c_double\)|integer\(c_int\))\s*(,[^:]*)?::\s*(.*)", line, flags=re.I)
            if not d:
                continue
            base, attrs, names = d.group(1).lower(), (d.group(2) or "").lower(), d.group(3)
            for nm in re.split(r",(?![^(]*\))", names):
                nm = re.sub(r"\(.*\)", "", nm).strip().lower()
                byval = "value" in attrs
                if base == "type(c_ptr)":
                    k = "handle" if byval else "handle_out"
                elif base.startswith("type(wb_"):
                    k = "struct_in"
                elif base == "real(c_double)":
                    k = "val_double" if byval else "ptr_double"
                else:
                    k = "val_int" if byval else "ptr_int"
                kind[nm] = k
        assert m.group(1).lower() == m.group(3).lower()
        out[m.group(3)] = [kind.get(d, "?" + d) for d in dummies]
    return out


@pytest.mark.parametrize("shim", [s[0] for s in SHIMS])
def test_fortran_interface_blocks_match_the_c_prototypes(shim):
    """Every bind(C) interface of a shim against the prototype of include/wbeuler.h: same number of arguments, and for each
    one the same passing convention -- `value` for C scalars, by reference (no `value`) for C pointers, type(c_ptr),value
    for the opaque handle and type(c_ptr),intent(out) for the handle** of *_create.  A mismatch here is exactly what a
    Fortran compiler would NOT catch (the interface is trusted) and what would corrupt the call at run time."""
    protos = _c_prototypes()
    ifs = _fortran_interfaces(os.path.join(ROOT, "fvm-source-wb_b200", "fortran", shim))
    assert len(ifs) >= 4
    for name, kinds in ifs.items():
        if name == "wb_last_error":
            assert kinds == []
            continue
        assert name in protos, name
        assert kinds == protos[name], (name, kinds, protos[name])


def test_python_binding_rejects_wrong_shapes():
    """The library copies prod(shape) doubles from every array pointer it is handed: the binding checks dtype, layout AND
    shape before a pointer leaves Python (no GPU needed)."""
    import numpy as np
    import wbeuler
    a = np.zeros((4, 3, 4))
    assert wbeuler._ptr(a, (4, 3, 4))
    with pytest.raises(wbeuler.WBError):
        wbeuler._ptr(a, (3, 4, 4))              # transposed-but-contiguous
    with pytest.raises(wbeuler.WBError):
        wbeuler._ptr(a[:2], (4, 3, 4))          # too small
    with pytest.raises(wbeuler.WBError):
        wbeuler._ptr(a.astype(np.float32), (4, 3, 4))
    with pytest.raises(wbeuler.WBError):
        wbeuler._ptr([[0.0]], (1, 1))           # not an ndarray


def test_output_number_format_is_fortrans_1pe12_5(lib):
    """output_file writes '(7(1PE12.5,1X))' (benchmark_2d.f90:139): one digit before the point, five after, two-digit
    exponent -- and gfortran's three-digit form without the letter when the exponent needs it."""
    def f(v):
        buf = C.create_string_buffer(16)
        lib.wb_format_1pe12_5(C.c_double(v), buf)
        return buf.value.decode()
    assert f(1.0) == " 1.00000E+00" and f(-0.5) == "-5.00000E-01" and f(0.0) == " 0.00000E+00"
    assert f(123456.789) == " 1.23457E+05" and f(-9.999996e-7) == "-1.00000E-06"
    assert f(1e-100) == " 1.00000-100" and f(-2.5e120) == "-2.50000+120"
    assert f(float("nan")).strip() == "NaN" and f(float("inf")).strip() == "Infinity"
    assert all(len(f(v)) == 12 for v in (1.0, -1e-300, 3e200, float("nan")))
