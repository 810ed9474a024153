"""The C-ABI library builds, loads and exports every symbol include/mcbrat_cuda.h declares.
No compute call is made here (no GPU in this container)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from mcbrat3d_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcbrat_cuda.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcb_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.dirname(_lib.LIB_PATH)], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return C.CDLL(_lib.LIB_PATH)


def test_header_declares_what_the_binding_lists():
    assert _declared() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in _declared():
        assert hasattr(lib, name), "missing export " + name


def test_only_sm_100a_code_is_embedded():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.mcb_options) == 84          # 11 reference-facing words + 9 tune knobs + 1 reserved
    assert C.sizeof(_lib.mcb_counters) == 128
    assert _lib.EVENT_DTYPE.itemsize == 96
    assert _lib.EVENT_DTYPE.fields["path"][1] == 48 and _lib.EVENT_DTYPE.fields["dir"][1] == 80


def test_create_fails_loudly_without_a_gpu(lib):
    """No CPU fallback: without a CUDA device mcb_create reports an error and hands back no handle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    lib.mcb_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    rc = lib.mcb_create(0, C.byref(h))
    assert rc != 0 and not h.value
    from mcbrat3d_b200 import domains
    from mcbrat3d_b200.monteCarloRadiativeTransfer import new_Integrator
    d, _ = domains.homogeneous_slab()
    with pytest.raises(_lib.McbError):
        new_Integrator(d)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under mcbrat3d_b200/ may reference it."""
    pkg = os.path.join(ROOT, "mcbrat3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)


def test_no_environment_variables_on_the_launch_path():
    """Measurement knobs live in mcb_options.tune*; the CUDA sources read no environment variable."""
    csrc = os.path.join(ROOT, "mcbrat3d_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh")):
            assert "getenv" not in open(os.path.join(csrc, f)).read(), f


def test_fortran_module_binds_every_declared_symbol():
    """fortran/mcbrat_cuda_mod.f90 (the ISO_C_BINDING layer of the reference-side integration) has a bind(C) interface
    for every entry point of the header, lists it as public, and mirrors the option struct field by field."""
    text = open(os.path.join(ROOT, "fortran", "mcbrat_cuda_mod.f90")).read()
    bound = set(re.findall(r'bind\(C,\s*name="(mcb_\w+)"\)', text))
    assert bound == set(_declared()), (sorted(set(_declared()) - bound), sorted(bound - set(_declared())))
    public = re.search(r"public :: mcb_create.*?\n\n", text, re.S).group(0)
    for name in _declared():
        assert re.search(r"\b%s\b" % name, public), "not public: " + name
    block = re.search(r"type, bind\(C\), public :: mcb_options(.*?)end type mcb_options", text, re.S).group(1)
    block = re.sub(r"!.*", "", block)
    fields = []
    for decl in re.findall(r"::\s*(.+)", block):
        for f in decl.split(","):
            m = re.match(r"\s*(\w+)(?:\((\d+)\))?", f)
            fields += [m.group(1)] * int(m.group(2) or 1)
    want = []
    for n, t in _lib.mcb_options._fields_:
        want += [n] * (C.sizeof(t) // 4)
    assert fields == want, (fields, want)
