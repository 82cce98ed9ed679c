"""CPU tests of the drop-in boundary: libfedem_b200.so loads, exports every symbol that
include/fedem_b200.h declares, and fails loudly (no CPU fallback) when no B200 is present."""
import ctypes as C
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "fedem_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    # fsr_* entry points + the exported names kept from the reference (stressInterface.C, solverInterface.C)
    return sorted(set(re.findall(r"\b(fsr_[a-z0-9_]+|initSolverArgs|solveStress|solveGage|solveModes|(?:get|save)Part[A-Za-z]+)\s*\(", txt)))


def test_header_symbols_are_exported():
    from fedem_solvers_b200 import _lib
    lib = _lib.load_library()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fedem_b200.h but not exported"
    bound = {s[0] for s in _lib.SYMBOLS}
    assert set(names) <= bound, set(names) - bound


def test_fortran_interface_covers_header():
    """fortran/fedem_b200_mod.f90 (ISO_C_BINDING interface blocks, compiled by the reference's
    Fortran host code) binds every C entry point by its exact name."""
    f90 = open(os.path.join(ROOT, "fortran", "fedem_b200_mod.f90")).read().lower()
    for n in _declared_symbols():
        assert f'name="{n.lower()}"' in f90 or f"name='{n.lower()}'" in f90, n


def test_part_state_entry_points_without_parts():
    """getPartStressStateSize & co. before any part is registered: the reference's 'not allocated' answer."""
    from fedem_solvers_b200 import _lib
    lib = _lib.load_library()
    assert lib.getPartDeformationStateSize(7) == -999 and lib.getPartStressStateSize(7) == -999
    buf = np.zeros(8)
    assert lib.savePartStressState(7, buf.ctypes.data_as(_lib._D), 8) and buf[3] == 7.0   # header only, like the reference
    assert not lib.savePartStressState(7, buf.ctypes.data_as(_lib._D), 3)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fedem_solvers_b200 import StressRecovery, FsrError, fatigue
    from fedem_solvers_b200.model import plate_part
    part = plate_part(2, 2, ngen=1, seed=1)
    with pytest.raises(FsrError, match="no CUDA device|CUDA"):
        StressRecovery(part)
    with pytest.raises(FsrError):
        fatigue(np.zeros((1, 4)), 1.0, [15.117, 17.146, 4.0, 5.0])


def test_product_does_not_import_oracle():
    """The product package must never touch oracle/ (only tests/, smoke() and bench's CPU legs may)."""
    pkg = os.path.join(ROOT, "fedem_solvers_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle_bind" not in src and "liboracle" not in src and "orc_" not in src, f


def _build_c_driver():
    import subprocess
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "c_abi_driver")
    src = os.path.join(ROOT, "tests", "c_abi_driver.c")
    libdir = os.path.join(ROOT, "fedem_solvers_b200", "lib")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L", libdir, "-lfedem_b200", "-lm", f"-Wl,-rpath,{libdir}"])
    return exe


def test_plain_c_caller_links_and_fails_loudly_without_gpu():
    """a C translation unit that sees only include/fedem_b200.h compiles, links and runs"""
    import subprocess
    exe = _build_c_driver()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    import torch
    if not torch.cuda.is_available():
        assert "no usable B200" in r.stdout


@pytest.mark.gpu
def test_plain_c_caller_on_gpu():
    import subprocess
    exe = _build_c_driver()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "C ABI driver OK" in r.stdout, r.stdout + r.stderr
