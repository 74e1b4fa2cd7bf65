"""The drop-in boundary without a GPU: libsph_b200.so loads, exports every symbol include/sph_b200.h declares,
the ctypes structs match the C layout, and the product path fails LOUDLY (no CPU fallback) when no sm_100
device is visible."""
import ctypes as C
import os
import re
import subprocess

import pytest
from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "sph_b200.h")


@pytest.fixture(scope="module")
def libsph():
    from astrophysical_sph_b200 import libsph as L

    L.build()          # nvcc cross-compiles sm_100a without a GPU
    return L


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(libsph):
    names = declared_functions()
    assert len(names) >= 20
    lib = C.CDLL(libsph.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sph_b200.h but not exported"
    assert sorted(libsph.ABI_SYMBOLS) == names
    assert lib.sph_abi_version() == 2


def test_library_is_sm100a_only(libsph):
    out = subprocess.run(["cuobjdump", "-lelf", libsph.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts_match_header(libsph):
    # compile a tiny C program against the public header and compare sizeof/offsetof with ctypes
    import tempfile

    prog = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "sph_b200.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu %d %d %zu\n", sizeof(sph_params), offsetof(sph_params, m), offsetof(sph_params, device),
               sizeof(sph_step_info), sizeof(sph_timings), offsetof(sph_params, U_iso), SPH_FLAG_COUNT_VISITS,
               SPH_FLAG_SERIAL_PHASES, offsetof(sph_params, flags));
        return 0;
    }"""
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        vals = [int(x) for x in subprocess.check_output([exe]).split()]
    P = libsph.SphParams
    assert vals == [C.sizeof(P), P.m.offset, P.device.offset, C.sizeof(libsph.SphStepInfo), C.sizeof(libsph.SphTimings),
                    P.U_iso.offset, libsph.FLAG_COUNT_VISITS, libsph.FLAG_SERIAL_PHASES, P.flags.offset]


def test_no_device_fails_loudly(libsph):
    if libsph.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(libsph.SphError) as e:
        libsph.SphB200(1000)
    assert e.value.code == libsph.SPH_ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)


def test_argument_validation_precedes_device_probe(libsph):
    lib = libsph.lib()
    h = C.c_void_p()
    p = libsph.SphParams(10, 50, 0, 1.0, 1.0, 5 / 3, 1.0, 0.5, 1.0, 2.0, 0.0, 0, 0)      # N < 64
    assert lib.sph_create(C.byref(p), C.byref(h)) == libsph.SPH_ERR_INVALID
    p = libsph.SphParams(1000, 500, 0, 1.0, 1.0, 5 / 3, 1.0, 0.5, 1.0, 2.0, 0.0, 0, 0)   # Kh too large
    assert lib.sph_create(C.byref(p), C.byref(h)) == libsph.SPH_ERR_INVALID
    assert lib.sph_create(None, C.byref(h)) == libsph.SPH_ERR_INVALID
    assert b"sph_create" in lib.sph_last_error(None)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "astrophysical-sph_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
