"""Error budget of the lean projection (slicer_b200/csrc/lean_math.h): the SAME header the CUDA kernels compile is built for
the host together with tools/lean_math_check.cpp and run against the reference's chain with this machine's libm.  The MUFU
seeds of the GPU are emulated by correctly rounded float reciprocals perturbed by up to 2^-21.5 (worse than the hardware's).
Nothing the guard lets through may differ from the reference (decision or float bits), and the observed error must stay far
inside the guard."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("lean") / "lean_math_check")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tools", "lean_math_check.cpp"), "-lm"])
    return exe


@pytest.mark.parametrize("fov_deg,npix", [(2.0, 256), (5.0, 2048), (5.0, 8192), (20.0, 1024), (34.0, 512)])
def test_lean_projection_inside_its_guard(checker, fov_deg, npix):
    r = subprocess.run([checker, "3000000", str(17 + npix), str(fov_deg), str(npix)], capture_output=True, text=True)
    out = json.loads(r.stdout)
    assert r.returncode == 0 and out["mismatch_unflagged"] == 0 and out["rejected_by_lean_only"] == 0, out
    assert out["accepted_ref"] > 1_000_000
    assert out["max_dv_units_2m53"] <= 16 < out["eta_units_2m53"]       # observed error (~4 units) against the guard (64 units of 2^-53)
    assert 0 < out["flagged"] < out["accepted_ref"] // 1000            # the libm path is the exception


def test_wide_fields_are_left_to_the_general_chain(checker):
    r = subprocess.run([checker, "1000", "1", "60.0", "512"], capture_output=True, text=True)
    assert r.returncode == 2 and "too wide" in r.stdout
