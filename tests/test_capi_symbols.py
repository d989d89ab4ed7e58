"""The C-ABI library loads on a CPU-only box and exports every symbol include/ddm_b200.h
declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "ddm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ddm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    names = _declared_functions()
    for must in ("ddm_create", "ddm_destroy", "ddm_simulate", "ddm_simulate_trialwise", "ddm_last_output_dlpack",
                 "ddm_set_normals_debug", "ddm_last_error", "ddm_export_normals", "ddm_microbench"):
        assert must in names
    assert len(names) >= 20


def test_library_exports_every_declared_symbol():
    from bayesflow_nddms_b200 import _build, _capi

    assert os.path.exists(_build.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_build.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/ddm_b200.h but not exported"
    # the ctypes prototypes cover the same surface
    assert sorted(_capi.SIGNATURES) == _declared_functions()
    assert _capi.load().ddm_version() == 200


def test_enums_match_header():
    from bayesflow_nddms_b200 import _capi

    src = open(os.path.join(ROOT, "include", "ddm_b200.h")).read()
    for name, val in (("DDM_MODEL_BASIC", _capi.MODEL_BASIC), ("DDM_MODEL_ALPHA", _capi.MODEL_ALPHA),
                      ("DDM_MODEL_ALPHA_DC", _capi.MODEL_ALPHA_DC), ("DDM_MODEL_ALPHA_SCALE", _capi.MODEL_ALPHA_SCALE),
                      ("DDM_MODEL_ALPHA_SCALE2", _capi.MODEL_ALPHA_SCALE2), ("DDM_MODEL_TRIALWISE", _capi.MODEL_TRIALWISE),
                      ("DDM_ERR_NEGATIVE_BOUND", _capi.ERR_NEGATIVE_BOUND), ("DDM_FLAG_OUT_F32", _capi.FLAG_OUT_F32),
                      ("DDM_FLAG_KEEP_STEPS", _capi.FLAG_KEEP_STEPS), ("DDM_FLAG_FORCE_GENERIC", _capi.FLAG_FORCE_GENERIC),
                      ("DDM_FLAG_OUT_STATE", _capi.FLAG_OUT_STATE), ("DDM_FLAG_F32_NORMALS", _capi.FLAG_F32_NORMALS)):
        m = re.search(rf"\b{name}\s*=\s*(-?\d+)", src)
        assert m and int(m.group(1)) == val, name
    # ddm_stats layout mirrors the header field order
    fields = re.search(r"typedef struct ddm_stats \{(.*?)\} ddm_stats;", src, flags=re.S).group(1)
    fields = re.sub(r"/\*.*?\*/", "", fields, flags=re.S)
    order = [n for decl in fields.split(";") if decl.strip() for n in re.findall(r"(\w+)\s*(?:,|$)", decl.strip().split(None, 1)[1])]
    assert order == [f for f, _ in _capi.Stats._fields_]


def test_no_cpu_fallback_without_a_gpu():
    """On a box without a B200 the product path must fail loudly, never fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import bayesflow_nddms_b200 as pkg
    from bayesflow_nddms_b200 import basic_ddm_dc

    with pytest.raises(pkg.DDMError, match="no CPU fallback"):
        pkg.DDMSimulator()
    pkg.set_default_simulator(None)
    with pytest.raises(pkg.DDMError):
        basic_ddm_dc.simulate_trials([1.0, 1.0, 0.5, 0.3, 1.0], 10)


def test_product_package_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "bayesflow_nddms_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "ddm_oracle" not in text.replace("oracle/ddm_oracle.c:orc_philox_normals6", ""), f
