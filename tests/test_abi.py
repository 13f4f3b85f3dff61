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


def test_specialised_kernel_mode_set_is_consistent():
    """the mode set substituted by scripts/gen_spec.py must be the one the launcher checks model cards against"""
    import importlib.util
    src = open(os.path.join(ROOT, "scripts", "gen_spec.py")).read()
    spec = eval("dict(" + src.split("SPEC = dict(")[1].split(")\n")[0] + ")")
    hdr = open(os.path.join(ROOT, "xyce_b200", "csrc", "b4_kernels.cuh")).read()
    modes = [int(v) for v in hdr.split("kSpecModes[17] = {")[1].split("}")[0].split(",")]
    lib = xyce_b200.load_library()
    names = lib.xgpu_b4_field_names(1).decode().split()
    assert len(names) == len(modes) == 17
    for n, m in zip(names, modes):
        if m == -2:
            assert n not in spec
        else:
            assert spec[n] == m, n
