"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/mbrf.h declares,
and fails loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol(mbrf):
    from multiband_rf_pulse_design_b200 import _lib
    declared = _lib.declared_symbols()
    assert len(declared) >= 15
    handle = ctypes.CDLL(_lib.library_path())
    missing = [s for s in declared if not hasattr(handle, s)]
    assert not missing, f"declared in include/mbrf.h but not exported: {missing}"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes signature table out of sync with include/mbrf.h"


def test_no_torch_or_cxx_types_in_header():
    import re
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "mbrf.h")).read(), flags=re.S)
    for banned in ("torch", "at::", "std::", "template", "class "):
        assert banned not in text


def test_only_sm100a_code_in_library(mbrf):
    from multiband_rf_pulse_design_b200 import _lib
    try:
        out = subprocess.run(["cuobjdump", "--list-elf", _lib.library_path()], capture_output=True, text=True).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not on PATH")
    assert "sm_100a" in out
    assert all("sm_100a" in line for line in out.splitlines() if "ELF file" in line)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multiband_rf_pulse_design_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".c", ".m")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.replace("the oracle", ""), f"{f} mentions the oracle"


def test_fails_loudly_without_device(mbrf):
    if mbrf.lib().mbrf_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(mbrf.MbrfError) as e:
        mbrf.blochC(np.ones(4), np.zeros(4), 1e-5, 1.0, 1.0, np.zeros(3), 0.0)
    assert e.value.code == -2 and "no CPU path" in e.value.message
    with pytest.raises(mbrf.MbrfError):
        mbrf.abrx(np.ones(4) * 0.1, np.ones(4), np.zeros(3))


def test_abrx_argument_errors_match_reference(mbrf):
    # abrx.c:44-45 — same message as the reference's mexErrMsgTxt
    with pytest.raises(ValueError, match="rf and gradient vectors are of different lengths"):
        mbrf.abrx(np.ones(8), np.ones(7), np.zeros(3))


def test_solver_option_range(mbrf):
    # host-side knobs of the PDHG solver (include/mbrf.h): indices 0..5, positive values only; no device needed
    lib = mbrf.lib()
    defaults = {0: 0.9, 1: 0.2, 2: 0.8, 3: 0.36, 4: 0.5, 5: 1.0}
    for which in (-1, 6, 99):
        assert lib.mbrf_pdhg_set_option(which, 1.0) != 0
    for bad in (0.0, -1.0, float("nan")):
        assert lib.mbrf_pdhg_set_option(2, bad) != 0
    for which in range(6):
        assert lib.mbrf_pdhg_set_option(which, defaults[which]) == 0
    for mode in (-1, 3):
        assert lib.mbrf_pdhg_set_halpern(mode) != 0
    assert lib.mbrf_pdhg_set_halpern(2) == 0
