"""The C-ABI library loads on a CPU-only box and exports every symbol include/xyce_b200.h declares;
creating a context without a GPU fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import xyce_b200

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "xyce_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xgpu_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = xyce_b200.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n


def test_field_lists_are_consistent():
    lib = xyce_b200.load_library()
    for which in range(5):
        names = lib.xgpu_b4_field_names(which).decode().split()
        assert len(names) == lib.xgpu_b4_field_count(which) > 0
        assert len(set(names)) == len(names)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        xyce_b200.Engine(0)


def test_benchmark_fixture_matches_field_lists():
    from xyce_b200 import workloads as wl
    lib = xyce_b200.load_library()
    rec = wl.load_b4_records()
    for which, key in enumerate(("model_d", "model_i", "size_d", "inst_d", "inst_i")):
        assert list(rec["names_" + key]) == lib.xgpu_b4_field_names(which).decode().split()


def test_specialised_kernel_mode_tuples_are_consistent():
    """bsim4_spec_tuples.def is the single source of the mode tuples: scripts/gen_spec.py substitutes them by name in
    XB_B4_MODEL_I order, the launcher checks model cards against the same rows -- the name order must be the library's."""
    import re
    src = open(os.path.join(ROOT, "scripts", "gen_spec.py")).read()
    names_py = eval(src.split("NAMES = ")[1].split("]")[0] + "]")
    lib = xyce_b200.load_library()
    names = lib.xgpu_b4_field_names(1).decode().split()
    assert names_py == names and len(names) == 17
    rows = re.findall(r"^\s*X\(([-0-9, ]+)\)", open(os.path.join(ROOT, "xyce_b200", "csrc", "bsim4_spec_tuples.def")).read(), re.M)
    ids = [int(r.split(",")[0]) for r in rows]
    assert ids == list(range(len(ids))) and all(len(r.split(",")) == 18 for r in rows)
    hdr = open(os.path.join(ROOT, "xyce_b200", "csrc", "bsim4_spec_tuples.def")).read()
    assert "kNumSpecTuples = %d" % len(ids) in hdr
    # every tuple has a launcher in the library
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", xyce_b200.capi.LIB_PATH], capture_output=True, text=True).stdout
    for i in ids:
        assert ("launch_b4_group_a2x%s" % ("" if i == 0 else str(i))) in syms
