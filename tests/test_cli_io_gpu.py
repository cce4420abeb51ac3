"""GPU tests of the drop-in `main` program and the model-file formats (through the C ABI)."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import GOLDEN, assert_close, load_npz
from oracle.cpu_model import CpuModel, have_ref
import ftrl_ffm_b200 as pkg
from ftrl_ffm_b200.build import MAIN

pytestmark = pytest.mark.gpu


def run_main(*args):
    out = subprocess.run([MAIN] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    return out.stdout


def cfg1_text(tmp, libffm=True):
    g = load_npz("cfg1.npz")
    data = {"row_ptr": g["row_ptr"], "field": g["field"].astype(np.int32), "feat": g["feat"], "val": g["val"],
            "label": g["label"].astype(np.int32)}
    path = os.path.join(tmp, "cfg1.ffm" if libffm else "cfg1.svm")
    pkg.synth.write_text(data, path, "libffm" if libffm else "libsvm")
    return g, path


@pytest.mark.parametrize("mt,libffm", [("FFM", True), ("LR", False), ("FM", False)])
def test_main_sequential_reproduces_reference_epoch_lines(mt, libffm):
    """`--batch_size 1` (reference-exact) in online mode = file order, n_threads 1 of the reference:
    BASELINE.md section 2 prints 0.6907/0.6883/0.6867/0.6854/0.6842 train, 0.6893/.../0.6837 eval"""
    with tempfile.TemporaryDirectory() as tmp:
        g, path = cfg1_text(tmp, libffm)
        out = run_main("--train_data", path, "--eval_data", path, "--model_type", mt, "--n_epochs", 5,
                       "--batch_size", 1, "--n_threads", 2)
        tr = [float(x) for x in re.findall(r"train loss: ([0-9.]+)", out)]
        ev = [float(x) for x in re.findall(r"eval loss: ([0-9.]+)", out)]
        assert len(tr) == 5 and len(ev) == 5
        np.testing.assert_allclose(tr, np.round(g[mt + "_train_loss"], 4), atol=1.01e-4)
        np.testing.assert_allclose(ev, np.round(g[mt + "_eval_loss"], 4), atol=1.01e-4)
        assert re.search(r"epoch 1 train time: [0-9.]+s, train loss: 0\.69", out)


def test_main_csr_cache_gives_the_same_epoch_lines():
    """--csr_cache true (additive flag): epochs run from the parsed / cached CSR in file order and print the
    same losses as the streaming text path; the second run is served from <file>.csr"""
    with tempfile.TemporaryDirectory() as tmp:
        g, path = cfg1_text(tmp, True)
        common = ["--train_data", path, "--eval_data", path, "--model_type", "FFM", "--n_epochs", 2, "--batch_size", 1]
        ref = run_main(*common)
        first = run_main(*common, "--csr_cache", "true")
        # train and eval are the same file here: the train load parses and writes the image, the eval load reads it
        assert os.path.exists(path + ".csr") and first.count("binary image") == 1
        second = run_main(*common, "--csr_cache", "true")
        assert second.count("binary image") == 2
        losses = lambda out: re.findall(r"(train|eval) loss: ([0-9.]+)", out)
        assert losses(ref) == losses(first) == losses(second) and len(losses(ref)) == 4


def test_main_offline_minibatch_and_model_path():
    with tempfile.TemporaryDirectory() as tmp:
        g, path = cfg1_text(tmp, True)
        model = os.path.join(tmp, "ffm.zst")
        out = run_main("--train_data", path, "--eval_data", path, "--online", "false", "--n_epochs", 5,
                       "--batch_size", 256, "--seed", 3, "--model_path", model, "--n_threads", 4)
        assert "Total number of samples loaded: 10000" in out and "parsing data time:" in out
        ev = [float(x) for x in re.findall(r"eval loss: ([0-9.]+)", out)]
        assert abs(ev[-1] - g["FFM_eval_loss"][4]) < 0.002          # north-star criterion 3
        # the saved file is the reference layout: [bias][lin_w][vec_w]
        m = pkg.FtrlModel("FFM", 10000, 8, 16)
        m.load_compressed_model(model)
        data = {"row_ptr": g["row_ptr"], "field": g["field"].astype(np.int32), "feat": g["feat"], "val": g["val"],
                "label": g["label"].astype(np.int32)}
        pred, pl = m.predict(data["row_ptr"], data["field"], data["feat"], data["val"], data["label"])
        assert abs(pl / 10000 - ev[-1]) < 1.01e-4
        assert abs(pkg.synth.auc(data["label"], pred) - g["FFM_auc"][4]) < 0.002


def test_main_rejects_bad_input_like_the_reference():
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "x.svm")
        open(p, "w").write("1 3:1 4:1\n0 5:1\n")
        r = subprocess.run([MAIN, "--train_data", p, "--model_type", "FFM"], capture_output=True, text=True)
        assert r.returncode != 0 and "FFM model requires libffm data format" in r.stderr   # cmd_option.cpp:110-113
        r = subprocess.run([MAIN, "--train_data", p, "--bogus", "1"], capture_output=True, text=True)
        assert r.returncode != 0 and "unknown argument" in r.stderr and "Usage" in r.stdout
        open(p, "w").write("1 3:1 oops\n")
        r = subprocess.run([MAIN, "--train_data", p, "--model_type", "LR"], capture_output=True, text=True)
        assert r.returncode != 0 and "wrong input" in r.stdout                              # parser.cpp:25


# ---- model files ---------------------------------------------------------------------------------
def test_load_files_written_by_the_reference():
    """golden files produced by the reference's own save_compressed_model / save_model"""
    g = load_npz("model_files.npz")
    m = pkg.FtrlModel("LR", 50)
    m.load_compressed_model(os.path.join(GOLDEN, "ref_lr.zst"))
    assert m.bias == g["lr_bias"][0]
    assert np.array_equal(m.lin_w, g["lr_lin_w"])                       # bit-equal (tests/test_model.cpp:51-66)
    m = pkg.FtrlModel("FFM", 50, 4, 4)
    m.load_compressed_model(os.path.join(GOLDEN, "ref_ffm.zst"))
    assert m.bias == g["ffm_bias"][0] and np.array_equal(m.lin_w, g["ffm_lin_w"])
    assert np.array_equal(m.vec_w, g["ffm_vec_w"])                      # tests/test_model.cpp:86-102
    m2 = pkg.FtrlModel("FFM", 50, 4, 4)
    m2.load_model(os.path.join(GOLDEN, "ref_ffm.txt"))
    np.testing.assert_allclose(m2.vec_w, g["ffm_vec_w"], rtol=0, atol=0)  # shortest round-trip digits
    np.testing.assert_allclose(m2.lin_w, g["ffm_lin_w"], rtol=1e-5)       # ostream default precision (6 digits)


def test_save_load_round_trips_like_reference_tests():
    rng = np.random.default_rng(2)
    sample = {"row_ptr": [0, 7], "field": [1, 1, 1, 3, 12, 111, 8], "feat": [3, 0, 2, 10, 4, 1, 8],
              "val": [3, 1, 0, 1, 0, 0, 8]}                              # tests/test_model.cpp:70-71
    with tempfile.TemporaryDirectory() as tmp:
        for mt in ("LR", "FM", "FFM"):
            a = pkg.FtrlModel(mt, 50, 4, 4, seed=1)
            b = pkg.FtrlModel(mt, 50, 4, 4, seed=2)
            pa, _ = a.predict(**sample)
            pb, _ = b.predict(**sample)
            assert pa[0] != pb[0]
            path = os.path.join(tmp, mt + ".zst")
            a.save_compressed_model(path, 10)
            b.load_compressed_model(path)
            pb, _ = b.predict(**sample)
            assert pa[0] == pb[0]                                         # bit-equal
            if mt == "FFM":
                c = pkg.FtrlModel(mt, 50, 4, 4, seed=3)
                tpath = os.path.join(tmp, "ffm.txt")
                a.save_model(tpath)
                c.load_model(tpath)
                pc, _ = c.predict(**sample)
                assert abs(pc[0] - pa[0]) <= 1e-4 * max(1.0, abs(pa[0]))  # tests/test_model.cpp:82
        # wrong-size file is an error, not a crash
        lr = pkg.FtrlModel("LR", 51)
        with pytest.raises(pkg.FtrlError):
            lr.load_compressed_model(os.path.join(tmp, "LR.zst"))


@pytest.mark.skipif(not have_ref(), reason="prebuilt oracle/_ref did not travel")
def test_reference_loads_our_files():
    """files written by ftrl_save_model / ftrl_save_model_text are read back by the reference's own loaders"""
    with tempfile.TemporaryDirectory() as tmp:
        a = pkg.FtrlModel("FFM", 50, 4, 4, seed=5)
        zpath, tpath = os.path.join(tmp, "a.zst"), os.path.join(tmp, "a.txt")
        a.save_compressed_model(zpath, 3)
        a.save_model(tpath)
        r = CpuModel("ref", "FFM", 50, 4, 4)
        r._fn("load_compressed")(r.h, zpath.encode())
        st = r.get_state()
        assert np.array_equal(st["vec_w"], a.vec_w) and np.array_equal(st["lin_w"], a.lin_w)
        assert st["bias"][0] == np.float32(a.bias)
        r2 = CpuModel("ref", "FFM", 50, 4, 4)
        r2._fn("load_text")(r2.h, tpath.encode())
        st2 = r2.get_state()
        assert np.array_equal(st2["vec_w"], a.vec_w)
        np.testing.assert_allclose(st2["lin_w"], a.lin_w, rtol=1e-5)
