"""The C-ABI library loads on a CPU-only box and exports every function include/slicer_b200.h declares.
No compute call is made here; creating a handle without a GPU must fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from slicer_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "slicer_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(slicer_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(capi.EXPORTS)


def test_library_exports_every_symbol():
    capi.build()
    lib = C.CDLL(capi.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_struct_layout_matches_header():
    # sizes the C compiler gives the header's structs (guards the ctypes mirrors against drift)
    import subprocess
    import tempfile

    prog = r"""
    #include <stdio.h>
    #include "slicer_b200.h"
    int main(void){ printf("%zu %zu %zu\n", sizeof(slicer_config), sizeof(slicer_plane_desc), sizeof(slicer_stats)); return 0; }
    """
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(capi.Config), C.sizeof(capi.PlaneDesc), C.sizeof(capi.Stats)]


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SlicerError) as e:
        capi.Slicer(npix_max=16)
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)


def test_plain_c_example_compiles_and_links(tmp_path):
    """examples/minimal.c uses the ABI from C99 (no C++ or torch types in the signatures) and links against the library."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "minimal")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "minimal.c"),
                        "-L" + os.path.join(root, "slicer_b200", "_build"), "-lslicer_b200", "-Wl,-rpath," + os.path.join(root, "slicer_b200", "_build"),
                        "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
