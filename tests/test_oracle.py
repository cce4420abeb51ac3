"""CPU tests: the C restatement (oracle/ftrl_oracle.c) against (a) the golden fixtures generated from the
reference and (b) the reference itself (oracle/_ref) when that build is present.  Bit-exact."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_npz, split_prefixed
from oracle.cpu_model import CpuModel, have_ref, scalar_fns
import ftrl_ffm_b200 as pkg

TRAJ = ["traj_lr.npz", "traj_fm.npz", "traj_fm_k5.npz", "traj_ffm.npz", "traj_ffm_dupfield.npz", "traj_ffm_k3.npz"]
MT = {"traj_lr.npz": "LR", "traj_fm.npz": "FM", "traj_fm_k5.npz": "FM"}


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_scalars_match_reference_tests():
    """tests/test_utils.cpp:13-24,40-43 of the reference"""
    fn = scalar_fns("oracle")
    assert fn["sgn"](1.0) == 1 and fn["sgn"](0.0) == -1 and fn["sgn"](-2.0) == -1
    assert abs(fn["sigmoid"](0.0) - 0.5) < 1e-7
    assert abs(fn["sigmoid"](1.0) - 0.7311) < 1e-4 and abs(fn["sigmoid"](-2.0) - 0.1192) < 1e-4
    assert abs(fn["loss"](1, 2.0) - 0.1269) < 1e-4 and abs(fn["loss"](0, 1.0) - 1.3133) < 1e-4


def test_scalars_golden():
    with open(os.path.join(GOLDEN, "scalars.json")) as f:
        g = json.load(f)
    fn = scalar_fns("oracle")
    m = CpuModel("oracle", "LR", 4, **g["hyper"])
    for x, want in g["sgn"]:
        assert np.float32(fn["sgn"](x)) == np.float32(want)
    for x, want in g["sigmoid"]:
        assert bits(fn["sigmoid"](x)) == bits(want)
    for y, x, want in g["loss"]:
        got = fn["loss"](y, x)
        assert got == want or (np.isinf(got) and np.isinf(want)) or abs(got - want) <= 1e-15 * abs(want)
    for n, z, want in g["weight"]:
        assert bits(m.weight(n, z)) == bits(want), (n, z)


def test_appendix_b_known_answers():
    """SURVEY.md appendix B vectors, regenerated from the reference into appendix_b.json"""
    with open(os.path.join(GOLDEN, "appendix_b.json")) as f:
        g = json.load(f)
    m = CpuModel("oracle", "LR", 8)
    for step in g["lr_steps"]:
        lg = m.train([0, 0], [3, 5], [1.0, 0.5], 1)
        st = m.get_state()
        assert bits(lg) == bits(step["logit"])
        assert list(bits([st["lin_w"][3], st["lin_z"][3], st["lin_n"][3]])) == list(bits(step["feat3"]))
        assert list(bits([st["lin_w"][5], st["lin_z"][5], st["lin_n"][5]])) == list(bits(step["feat5"]))
        assert list(bits(st["bias"])) == list(bits(step["bias"]))
    # hand check of SURVEY appendix B.1: step 1 w3 = 0.4 / 15005
    assert abs(g["lr_steps"][1]["feat3"][0] - 0.4 / 15005) < 1e-11
    m = CpuModel("oracle", "FFM", 8, 3, 2)
    m.set_state({"lin_w": np.array([0.01 * (i + 1) for i in range(8)], np.float32),
                 "vec_w": np.array([[0.1 * (i + 1) - 0.05 * j for j in range(6)] for i in range(8)], np.float32),
                 "bias": np.array([0.25, 0, 0], np.float32)})
    assert bits(m.predict([0, 1, 2], [1, 4, 6], [1.0, 0.5, 2.0])) == bits(g["ffm_forward"]["logit"])
    assert abs(g["ffm_forward"]["logit"] - 0.723749995) < 1e-8
    assert bits(m.predict([0, 1, 2], [1, 4, 6], [1.0, 0.5, 2.0], True)) == bits(g["ffm_forward"]["prob"])
    m = CpuModel("oracle", "FM", 8, 1, 2)
    m.set_state({"lin_w": np.array([0.01 * (i + 1) for i in range(8)], np.float32),
                 "vec_w": np.array([[0.1 * (i + 1) - 0.05 * j for j in range(2)] for i in range(8)], np.float32),
                 "bias": np.array([0.25, 0, 0], np.float32)})
    assert bits(m.predict([0, 0, 0], [1, 4, 6], [1.0, 0.5, 2.0])) == bits(g["fm_forward"]["logit"])
    assert abs(g["fm_forward"]["logit"] - 1.63624978) < 1e-7


@pytest.mark.parametrize("name", TRAJ)
def test_trajectory_golden_bit_exact(name):
    g = load_npz(name)
    mt = MT.get(name, "FFM")
    m = CpuModel("oracle", mt, int(g["n_feats"]), int(g["n_fields"]), int(g["k"]))
    m.set_state(split_prefixed(g, "s0_"))
    b = split_prefixed(g, "b_")
    logits, loss = m.train_csr(**b)
    assert np.array_equal(bits(logits), bits(g["logits"]))
    assert loss == float(g["loss"])
    st = m.get_state()
    for k, v in split_prefixed(g, "s1_").items():
        assert np.array_equal(bits(st[k]), bits(v)), k
    pred, pl = m.predict_csr(b["row_ptr"], b["field"], b["feat"], b["val"], b["label"])
    assert np.array_equal(bits(pred), bits(g["pred"]))
    assert pl == float(g["pred_loss"])


def test_cfg1_quality_anchor():
    """config 1 (data/libffm_data.txt, 5 epochs, file order): BASELINE.md section 2 numbers; FFM == FM == LR
    bit-for-bit from a cold start."""
    g = load_npz("cfg1.npz")
    data = {"row_ptr": g["row_ptr"], "field": g["field"].astype(np.int32), "feat": g["feat"], "val": g["val"],
            "label": g["label"].astype(np.int32)}
    np.testing.assert_allclose(g["FFM_train_loss"], [0.690715, 0.688291, 0.686674, 0.685360, 0.684224], atol=1e-6)
    np.testing.assert_allclose(g["FFM_eval_loss"], [0.689268, 0.687415, 0.685981, 0.684769, 0.683700], atol=1e-6)
    np.testing.assert_allclose(g["FFM_auc"], [0.949438, 0.950929, 0.951696, 0.952191, 0.952556], atol=1e-6)
    for mt in ("LR", "FM"):
        assert np.array_equal(g[mt + "_train_loss"], g["FFM_train_loss"])
    m = CpuModel("oracle", "FFM", 10000, 8, 16)
    for ep in range(5):
        _, ls = m.train_csr(**data)
        assert ls / 10000 == g["FFM_train_loss"][ep]
        pred, pl = m.predict_csr(data["row_ptr"], data["field"], data["feat"], data["val"], data["label"])
        assert pl / 10000 == g["FFM_eval_loss"][ep]
        assert abs(pkg.synth.auc(data["label"], pred) - g["FFM_auc"][ep]) < 1e-12
    st = m.get_state()
    assert np.array_equal(bits(st["lin_z"]), bits(g["FFM_final_lin_z"]))
    assert not st["vec_z"].any() and not st["vec_n"].any()  # cold-start invariant (appendix B.5)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("mt,k,nfl", [("LR", 1, 1), ("FM", 6, 1), ("FFM", 4, 5), ("FFM", 3, 3)])
def test_oracle_vs_reference_random(mt, k, nfl):
    rng = np.random.default_rng(100 + k)
    nf = 70
    a, r = CpuModel("oracle", mt, nf, nfl, k), CpuModel("ref", mt, nf, nfl, k)
    st = pkg.synth.random_state(rng, nf, a.row_len)
    a.set_state(st)
    r.set_state(st)
    for it in range(5):
        b = pkg.synth.random_csr(rng, 60, nf, nfl, dup_feat=(mt != "FFM"), dup_field=(it % 2 == 1))
        la, sa = a.train_csr(**b)
        lr_, sr = r.train_csr(**b)
        assert np.array_equal(bits(la), bits(lr_)) and sa == sr
    sa, sr = a.get_state(), r.get_state()
    for key in sa:
        assert np.array_equal(bits(sa[key]), bits(sr[key])), key


def test_batch_semantics_equal_sequential_at_batch_one():
    """the derived minibatch rule at n_rows == 1 equals the reference update where ffm.cpp:118 does not bite:
    LR and FM always; FFM from a cold start"""
    rng = np.random.default_rng(5)
    for mt, k, nfl in (("LR", 1, 1), ("FM", 4, 1)):
        nf = 40
        a, s = CpuModel("oracle", mt, nf, nfl, k), CpuModel("oracle", mt, nf, nfl, k)
        st = pkg.synth.random_state(rng, nf, a.row_len)
        a.set_state(st)
        s.set_state(st)
        b = pkg.synth.random_csr(rng, 80, nf, nfl, dup_feat=False, oob_frac=0.0)
        for r in range(80):
            one = pkg.synth.slice_csr(b, r, r + 1)
            la, _ = a.train_batch_csr(**one)
            ls, _ = s.train_csr(**one)
            assert abs(la[0] - ls[0]) <= 1e-6 * max(1, abs(ls[0]))
        sa, ss = a.get_state(), s.get_state()
        for key in sa:
            np.testing.assert_allclose(sa[key], ss[key], rtol=2e-6, atol=1e-5, err_msg=key)
