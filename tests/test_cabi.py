"""The C-ABI library loads without a GPU, exports every symbol include/b200dp.h declares, the Python
mirror of its constants agrees with the header, and the CUDA path fails LOUDLY (no fallback) here."""

import ctypes
import os
import re

import pytest

from accelerate_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "b200dp.h")).read()


def test_library_is_built_for_sm100a_and_loads():
    native.build()
    lib = native.load()
    assert lib.b200dp_version() == 100
    assert lib.b200dp_control_bytes() > 64 * 1024


def test_every_declared_symbol_is_exported():
    declared = set(re.findall(r"\b(b200dp_[a-z_0-9]+)\s*\(", HEADER))
    lib = ctypes.CDLL(native.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)


def test_constants_match_header():
    defs = dict(re.findall(r"#define\s+B200DP_([A-Z_0-9]+)\s+(-?\d+)", HEADER))
    for py_name, c_name in [("MAX_WORLD", "MAX_WORLD"), ("MAX_BLOCKS", "MAX_BLOCKS"), ("MAX_SLOTS", "MAX_SLOTS"),
                            ("SMALL_MAX", "SMALL_MAX"), ("MT_MAX", "MT_MAX"), ("F32", "F32"), ("BF16", "BF16"),
                            ("F16", "F16"), ("F64", "F64"), ("I32", "I32"), ("I64", "I64"), ("U8", "U8"),
                            ("HANDLE_CUDAIPC", "HANDLE_CUDAIPC"), ("HANDLE_POSIXFD", "HANDLE_POSIXFD"),
                            ("CH_GRAD", "CH_GRAD"), ("CH_USER", "CH_USER"), ("CH_SMALL", "CH_SMALL"), ("CH_PUSH", "CH_PUSH"), ("F_NO_END_BARRIER", "F_NO_END_BARRIER"),
                            ("ALGO_AUTO", "ALGO_AUTO"), ("ALGO_TWOSHOT", "ALGO_TWOSHOT"), ("ALGO_NVLS", "ALGO_NVLS"),
                            ("ALGO_ONESHOT", "ALGO_ONESHOT"), ("F_STATS", "F_STATS"), ("Q_LAUNCHES", "Q_LAUNCHES"),
                            ("Q_STATUS", "Q_STATUS"), ("Q_COMM_STREAM", "Q_COMM_STREAM"), ("ETIMEOUT", "ETIMEOUT")]:  # fmt: skip
        assert getattr(native, py_name) == int(defs[c_name]), py_name


def test_header_cites_the_reference_interface_for_each_data_path_entry():
    for fn in ["b200dp_allreduce_bucket", "b200dp_finalize_norm", "b200dp_pack", "b200dp_pull", "b200dp_exchange_small",
               "b200dp_pad_copy", "b200dp_allreduce_oneshot"]:  # fmt: skip
        pos = HEADER.index(f"int {fn}(")
        assert "src/accelerate/" in HEADER[max(0, pos - 1500) : pos], fn


def test_cuda_path_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.B200dpError, match="no CPU fallback"):
        native.Context(0, 1, 0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "accelerate_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
