"""CPU tests of the drop-in boundary: libftrl_b200.so loads, exports every symbol include/ftrl_b200.h
declares, and fails loudly (no CPU fallback) when no CUDA device is present.  No compute calls."""
import ctypes as C
import os
import re

import pytest

import ftrl_ffm_b200 as pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ftrl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ftrl_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_functions() == sorted(pkg.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    pkg.build_library()
    lib = pkg.load_library()
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in include/ftrl_b200.h but not exported"
    assert lib.ftrl_abi_version() == 1


def test_config_struct_layout_and_defaults():
    """cmd_option.h:49-63 defaults"""
    lib = pkg.load_library()
    cfg = pkg.Config()
    lib.ftrl_config_default(C.byref(cfg))
    assert (cfg.model_type, cfg.n_fields, cfg.n_feats, cfg.n_factors) == (2, 8, 10000, 16)
    assert abs(cfg.init_stddev - 0.02) < 1e-9 and cfg.init_mean == 0.0
    assert abs(cfg.w_alpha - 1e-4) < 1e-10 and cfg.w_beta == 1.0 and abs(cfg.w_l1 - 0.1) < 1e-8 and cfg.w_l2 == 5.0
    assert cfg.mode == pkg.MODE_BATCH and cfg.world_size == 1
    assert C.sizeof(pkg.Config) == 112
    assert C.sizeof(pkg.BatchStats) == 96


def test_bad_arguments_return_status_not_exceptions():
    lib = pkg.load_library()
    h = C.c_void_p()
    assert lib.ftrl_create(None, C.byref(h)) == -1
    cfg = pkg.Config()
    lib.ftrl_config_default(C.byref(cfg))
    cfg.model_type = 7
    rc = lib.ftrl_create(C.byref(cfg), C.byref(h))
    assert rc == -1 and b"model_type" in lib.ftrl_last_error(None)
    assert lib.ftrl_sync(None) == -1
    with pytest.raises(ValueError):
        pkg.FtrlModel("XGB", 10)


def test_no_cpu_fallback_without_a_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.FtrlError) as e:
        pkg.FtrlModel("LR", 16)
    assert e.value.status == -2  # FTRL_ERR_CUDA


def test_product_never_imports_the_oracle():
    """nothing under ftrl-ffm_b200/ may import, link or call the checker (comments may mention it)"""
    banned = ("import oracle", "from oracle", "cpu_model", "libftrl_oracle", "libftrl_ref", "ftrl_oracle_",
              "ftrl_ref_", "ftrl_oracle.h")
    for root, _, files in os.walk(os.path.join(ROOT, "ftrl-ffm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(root, f), errors="ignore").read()
                for b in banned:
                    assert b not in text, f"{f} references {b}"
