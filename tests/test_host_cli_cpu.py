"""CPU tests of the host program's argument handling (no GPU is touched before the options are accepted) and of
the bench contract of the `--impl reference` arm (the reference's own CPU path; runs without a GPU)."""
import json
import os
import subprocess
import sys

import pytest

from ftrl_ffm_b200.build import MAIN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.exists(MAIN), reason="host program not built")
def test_main_rejects_bad_options_before_touching_the_gpu(tmp_path):
    svm = tmp_path / "t.svm"
    svm.write_text("1 3:1 5:0.5\n0 2:1\n")
    # src/task/ftrl_offline.cpp:29-32: invalid model_type -> message + failure exit
    out = subprocess.run([MAIN, "--train_data", str(svm), "--model_type", "XYZ"], capture_output=True, text=True,
                         timeout=60)
    assert out.returncode != 0 and "model_type" in (out.stderr + out.stdout)
    # cmd_option.cpp:109-113: FFM needs libffm input
    out = subprocess.run([MAIN, "--train_data", str(svm), "--model_type", "FFM"], capture_output=True, text=True,
                         timeout=60)
    assert out.returncode != 0 and "FFM model requires libffm data format" in (out.stderr + out.stdout)
    # missing data file
    out = subprocess.run([MAIN, "--train_data", str(tmp_path / "none.svm"), "--model_type", "LR"], capture_output=True,
                         text=True, timeout=60)
    assert out.returncode != 0


@pytest.mark.skipif(not os.path.exists(MAIN), reason="host program not built")
def test_main_help_lists_reference_flags_and_additive_ones():
    out = subprocess.run([MAIN, "--help"], capture_output=True, text=True, timeout=60)
    text = out.stdout + out.stderr
    for flag in ("--model_path", "--model_type", "--online", "--n_fields", "--n_feats", "--n_factors", "--train_data",
                 "--eval_data", "--init_mean", "--init_stddev", "--w_alpha", "--w_beta", "--w_l1", "--w_l2",
                 "--n_threads", "--n_epochs", "--batch_size", "--device", "--seed", "--csr_cache"):
        assert flag in text, flag


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--cpu-samples", "256"], capture_output=True, text=True, timeout=300,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["higher_is_better"] is True
    assert line["metric"] == "FTRL-FFM train samples/sec" and line["value"] > 0
    assert line["config"]["workload"].startswith("cfg4: FFM n_fields=39 n_feats=10000000 k=8 batch=65536/GPU")
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["gpu_launches"] == 0
