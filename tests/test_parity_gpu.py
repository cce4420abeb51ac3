"""GPU parity tests (run with -m gpu on the B200 box).  Every call goes through the C ABI
(libftrl_b200.so); the oracle (oracle/ftrl_oracle.c) and, when its prebuilt .so travelled with the
snapshot, the reference itself (oracle/_ref) are the checkers.

Tolerances (BASELINE.json north_star):
  * forward logits at fixed weights: 1e-5 relative
  * batch-size-1 / sequential z,n,w trajectory: 1e-5 relative
  * minibatch training: final logloss / AUC within 0.002 of the reference's single-threaded run
"""
import numpy as np
import pytest

from conftest import assert_close, assert_state_close, load_npz, split_prefixed
from oracle.cpu_model import CpuModel, have_ref
import ftrl_ffm_b200 as pkg

pytestmark = pytest.mark.gpu

TRAJ = [("traj_lr.npz", "LR"), ("traj_fm.npz", "FM"), ("traj_fm_k5.npz", "FM"), ("traj_ffm.npz", "FFM"),
        ("traj_ffm_dupfield.npz", "FFM"), ("traj_ffm_k3.npz", "FFM")]
RTOL = 1e-5


def gpu_model(mt, nf, nfl, k, mode="batch", **kw):
    return pkg.FtrlModel(mt, n_feats=nf, n_fields=nfl, n_factors=k, mode=mode, **kw)


# ---------------------------------------------------------------------------------------------
# 1. sequential mode == the reference's per-sample trajectory (golden fixtures from the reference)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,mt", TRAJ)
def test_sequential_trajectory_matches_reference_golden(name, mt):
    g = load_npz(name)
    m = gpu_model(mt, int(g["n_feats"]), int(g["n_fields"]), int(g["k"]), mode="sequential")
    m.set_state(split_prefixed(g, "s0_"))
    b = split_prefixed(g, "b_")
    logits, loss = m.train(**b)
    assert_close(logits, g["logits"], RTOL, 1e-6, "logits")
    assert abs(loss - float(g["loss"])) <= 1e-9 * max(1.0, abs(float(g["loss"])))
    assert_state_close(m.get_state(), split_prefixed(g, "s1_"), RTOL, name=name)
    pred, pl = m.predict(b["row_ptr"], b["field"], b["feat"], b["val"], b["label"])
    assert_close(pred, g["pred"], RTOL, 1e-6, "predict")
    assert abs(pl - float(g["pred_loss"])) <= 1e-9 * max(1.0, abs(float(g["pred_loss"])))
    exact = np.mean(np.asarray(logits).view(np.uint32) == g["logits"].view(np.uint32))
    print(f"{name}: {exact * 100:.1f}% of logits bit-identical to the reference")


@pytest.mark.parametrize("name,mt", [("traj_ffm.npz", "FFM"), ("traj_fm.npz", "FM"), ("traj_lr.npz", "LR")])
def test_batch_size_one_calls_reproduce_trajectory(name, mt):
    """literally one sample per call (north_star: 'with batch size 1 ...')"""
    g = load_npz(name)
    m = gpu_model(mt, int(g["n_feats"]), int(g["n_fields"]), int(g["k"]), mode="sequential")
    o = CpuModel("oracle", mt, int(g["n_feats"]), int(g["n_fields"]), int(g["k"]))
    st0 = split_prefixed(g, "s0_")
    m.set_state(st0)
    o.set_state(st0)
    b = split_prefixed(g, "b_")
    n = 60
    for r in range(n):
        one = pkg.synth.slice_csr(b, r, r + 1)
        lg, _ = m.train(**one)
        lo, _ = o.train_csr(**one)
        assert_close(lg, lo, RTOL, 1e-6, f"logit row {r}")
    assert_state_close(m.get_state(), o.get_state(), RTOL, name=name)


def test_sequential_appendix_b_lr_steps():
    import json, os
    from conftest import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "appendix_b.json")))
    m = gpu_model("LR", 8, 1, 1, mode="sequential")
    m.set_state({"bias": np.zeros(3, np.float32), "lin_w": np.zeros(8, np.float32),
                 "lin_n": np.zeros(8, np.float32), "lin_z": np.zeros(8, np.float32)})
    for step in g["lr_steps"]:
        lg, _ = m.train([0, 2], [0, 0], [3, 5], [1.0, 0.5], [1])
        st = m.get_state()
        assert_close(lg[0], step["logit"], RTOL, 1e-9, "logit")
        assert_close([st["lin_w"][3], st["lin_z"][3], st["lin_n"][3]], step["feat3"], RTOL, 1e-9, "feat3")
        assert_close([st["lin_w"][5], st["lin_z"][5], st["lin_n"][5]], step["feat5"], RTOL, 1e-9, "feat5")
        assert_close(st["bias"], step["bias"], RTOL, 1e-9, "bias")


# ---------------------------------------------------------------------------------------------
# 2. forward logits at fixed weights (predict), both modes
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["batch", "sequential"])
@pytest.mark.parametrize("mt,nfl,k", [("FFM", 8, 16), ("FFM", 39, 4), ("FFM", 39, 8), ("FFM", 5, 3), ("FFM", 6, 2),
                                      ("FM", 1, 16), ("FM", 1, 5), ("LR", 1, 1)])
def test_forward_logits_fixed_weights(mode, mt, nfl, k):
    rng = np.random.default_rng(nfl * 100 + k)
    nf = 500
    m = gpu_model(mt, nf, nfl, k, mode=mode)
    o = CpuModel("oracle", mt, nf, nfl, k)
    st = pkg.synth.random_state(rng, nf, o.row_len)
    st["vec_w"] = rng.normal(0, 0.3, (nf, o.row_len)).astype(np.float32) if o.row_len else None
    st = {k_: v for k_, v in st.items() if v is not None}
    st["lin_w"] = rng.normal(0, 0.3, nf).astype(np.float32)
    st["bias"][0] = 0.25
    m.set_state(st)
    o.set_state(st)
    b = pkg.synth.random_csr(rng, 300, nf, nfl, max_nnz=min(nfl, 39) if mt == "FFM" else 30, dup_feat=True)
    got, gl = m.predict(b["row_ptr"], b["field"], b["feat"], b["val"], b["label"])
    want, wl = o.predict_csr(b["row_ptr"], b["field"], b["feat"], b["val"], b["label"])
    if mode == "sequential":
        assert_close(got, want, RTOL, 2e-6, "logit")
    else:
        # minibatch mode sums the terms of a logit in a different order than the reference.  Per logit: 1e-5
        # relative, plus the fp32 re-association bound of THAT sample -- 2e-6 (about sqrt(n) eps for its n <= 800
        # terms) times the sum of the magnitudes of its terms, which the oracle evaluates on |w|, |x|
        oa = CpuModel("oracle", mt, nf, nfl, k)
        oa.set_state({k_: np.abs(v) for k_, v in st.items()})
        mag, _ = oa.predict_csr(b["row_ptr"], b["field"], b["feat"], np.abs(b["val"]), b["label"])
        err = np.abs(np.asarray(got, np.float64) - want)
        tol = RTOL * np.abs(want) + 2e-6 * np.asarray(mag, np.float64)
        assert (err <= tol).all(), f"logit: worst excess {np.max(err - tol):.3e} at {int(np.argmax(err - tol))}"
    assert abs(gl - wl) <= 1e-5 * max(1.0, abs(wl))
    gp, _ = m.predict(b["row_ptr"], b["field"], b["feat"], b["val"], None, output_prob=True)
    wp, _ = o.predict_csr(b["row_ptr"], b["field"], b["feat"], b["val"], None, output_prob=True)
    assert_close(gp, wp, RTOL, 1e-6, "prob")


def test_appendix_b_forward_vectors():
    m = gpu_model("FFM", 8, 3, 2)
    m.set_state({"lin_w": np.array([0.01 * (i + 1) for i in range(8)], np.float32),
                 "vec_w": np.array([[0.1 * (i + 1) - 0.05 * j for j in range(6)] for i in range(8)], np.float32),
                 "bias": np.array([0.25, 0, 0], np.float32)})
    out, _ = m.predict([0, 3], [0, 1, 2], [1, 4, 6], [1.0, 0.5, 2.0])
    assert abs(out[0] - 0.723749995) <= 1e-5 * 0.7237
    m = gpu_model("FM", 8, 1, 2)
    m.set_state({"lin_w": np.array([0.01 * (i + 1) for i in range(8)], np.float32),
                 "vec_w": np.array([[0.1 * (i + 1) - 0.05 * j for j in range(2)] for i in range(8)], np.float32),
                 "bias": np.array([0.25, 0, 0], np.float32)})
    out, _ = m.predict([0, 3], [0, 0, 0], [1, 4, 6], [1.0, 0.5, 2.0])
    assert abs(out[0] - 1.63624978) <= 1e-5 * 1.636


# ---------------------------------------------------------------------------------------------
# 3. minibatch mode == the derived minibatch rule (oracle.train_batch_csr), live state
# ---------------------------------------------------------------------------------------------
CASES = [("FFM", 6, 4, {}), ("FFM", 8, 16, {}), ("FFM", 39, 8, {"max_nnz": 39}), ("FFM", 39, 4, {"max_nnz": 39}),
         ("FFM", 5, 3, {}), ("FFM", 6, 2, {}), ("FFM", 4, 8, {"dup_field": True}),
         ("FFM", 6, 4, {"dup_feat": True}), ("FM", 1, 16, {"dup_feat": True, "max_nnz": 20}),
         ("FM", 1, 5, {"dup_feat": True}), ("FM", 1, 6, {}), ("LR", 1, 1, {"dup_feat": True, "max_nnz": 20})]


@pytest.mark.parametrize("mt,nfl,k,kw", CASES)
def test_minibatch_matches_derived_rule(mt, nfl, k, kw):
    rng = np.random.default_rng(7 + nfl + k)
    nf = 300
    m = gpu_model(mt, nf, nfl, k)
    o = CpuModel("oracle", mt, nf, nfl, k)
    st = pkg.synth.random_state(rng, nf, o.row_len)
    m.set_state(st)
    o.set_state(st)
    for it, nrows in enumerate((1, 7, 256, 2000)):
        b = pkg.synth.random_csr(rng, nrows, nf, nfl, **kw)
        got, gl = m.train(**b)
        want, wl = o.train_batch_csr(**b)
        assert_close(got, want, RTOL, 2e-6, f"logits batch {it}")
        assert abs(gl - wl) <= 1e-6 * max(1.0, abs(wl))
        assert_state_close(m.get_state(), o.get_state(), rtol=2e-5, atol=2e-6, atol_z=2e-4, name=f"{mt} batch {it}")


def test_minibatch_hot_rows_span_many_chunks():
    """a feature present in every sample: its row is cut into chunks and recombined"""
    rng = np.random.default_rng(3)
    nf, nfl, k = 200, 5, 4
    for mt in ("FFM", "FM", "LR"):
        m = gpu_model(mt, nf, nfl, k)
        o = CpuModel("oracle", mt, nf, nfl, k)
        st = pkg.synth.random_state(rng, nf, o.row_len)
        m.set_state(st)
        o.set_state(st)
        b = pkg.synth.random_csr(rng, 5000, nf, nfl, max_nnz=5, min_nnz=5, oob_frac=0.0)
        b["feat"][b["row_ptr"][:-1]] = 17          # first feature of every row is id 17
        b["feat"][b["row_ptr"][:-1] + 1] = 18 + (np.arange(5000) % 2)
        got, gl = m.train(**b)
        want, wl = o.train_batch_csr(**b)
        assert_close(got, want, RTOL, 2e-6, "logits")
        # sums over 5000 occurrences accumulate in fp32 on the GPU (fp64 in the oracle)
        assert_state_close(m.get_state(), o.get_state(), rtol=1e-4, atol=1e-5, atol_z=5e-3, name=mt)


@pytest.mark.parametrize("mt,nfl,k", [("LR", 1, 1), ("FM", 1, 8)])
def test_minibatch_at_batch_one_equals_reference_sequence(mt, nfl, k):
    """no ffm.cpp:118 term in LR/FM: minibatch mode with one sample per call is the reference trajectory"""
    rng = np.random.default_rng(21)
    nf = 100
    m = gpu_model(mt, nf, nfl, k)
    o = CpuModel("oracle", mt, nf, nfl, k)
    st = pkg.synth.random_state(rng, nf, o.row_len)
    m.set_state(st)
    o.set_state(st)
    b = pkg.synth.random_csr(rng, 40, nf, nfl, dup_feat=False, oob_frac=0.0)
    for r in range(40):
        one = pkg.synth.slice_csr(b, r, r + 1)
        got, _ = m.train(**one)
        want, _ = o.train_csr(**one)
        assert_close(got, want, RTOL, 2e-6, f"row {r}")
    assert_state_close(m.get_state(), o.get_state(), rtol=2e-5, atol=2e-6, atol_z=2e-4, name=mt)


# ---------------------------------------------------------------------------------------------
# 4. config 1 end to end (data/libffm_data.txt, 5 epochs)
# ---------------------------------------------------------------------------------------------
def cfg1_data():
    g = load_npz("cfg1.npz")
    data = {"row_ptr": g["row_ptr"], "field": g["field"].astype(np.int32), "feat": g["feat"], "val": g["val"],
            "label": g["label"].astype(np.int32)}
    return g, data


@pytest.mark.parametrize("mt", ["FFM", "FM", "LR"])
def test_cfg1_sequential_equals_reference_curve(mt):
    g, data = cfg1_data()
    if mt != "FFM":
        data = dict(data, field=np.zeros_like(data["field"]))
    m = gpu_model(mt, 10000, 8, 16, mode="sequential")
    for ep in range(5):
        _, ls = m.train(**data)
        assert abs(ls / 10000 - g[mt + "_train_loss"][ep]) < 1e-6
        pred, pl = m.predict(data["row_ptr"], data["field"], data["feat"], data["val"], data["label"])
        assert abs(pl / 10000 - g[mt + "_eval_loss"][ep]) < 1e-6
        assert abs(pkg.synth.auc(data["label"], pred) - g[mt + "_auc"][ep]) < 1e-5
    st = m.get_state()
    assert_close(st["lin_z"], g[mt + "_final_lin_z"], RTOL, 1e-4, "lin_z")
    assert_close(st["lin_n"], g[mt + "_final_lin_n"], RTOL, 1e-6, "lin_n")
    if mt != "LR":
        assert not st["vec_z"].any() and not st["vec_n"].any()  # cold-start invariant, appendix B.5
    if mt == "FFM":
        assert m.has_zero_weights()                            # tests/test_task.cpp:31,41 (FFM only)


@pytest.mark.parametrize("batch", [64, 1000])
def test_cfg1_minibatch_quality_within_0p002(batch):
    g, data = cfg1_data()
    m = gpu_model("FFM", 10000, 8, 16)
    for ep in range(5):
        for r0 in range(0, 10000, batch):
            m.train(**pkg.synth.slice_csr(data, r0, min(r0 + batch, 10000)), want_logits=False)
    pred, pl = m.predict(data["row_ptr"], data["field"], data["feat"], data["val"], data["label"])
    assert abs(pl / 10000 - g["FFM_eval_loss"][4]) < 0.002
    assert abs(pkg.synth.auc(data["label"], pred) - g["FFM_auc"][4]) < 0.002


# ---------------------------------------------------------------------------------------------
# 5. edge cases of the reference's tests: empty / ragged input, everything filtered, wide samples
# ---------------------------------------------------------------------------------------------
def test_empty_and_degenerate_batches():
    for mt in ("FFM", "FM", "LR"):
        for mode in ("batch", "sequential"):
            m = gpu_model(mt, 50, 4, 4, mode=mode)
            st0 = m.get_state()
            lg, ls = m.train([0], [], [], [], [])
            assert len(lg) == 0 and ls == 0.0
            st1 = m.get_state()
            for key in st0:
                assert np.array_equal(st0[key], st1[key]), (mt, mode, key)
            # rows with no (valid) features: logit = bias, only the bias moves
            b = {"row_ptr": [0, 0, 2, 2], "field": [9, 1], "feat": [3, 100], "val": [1.0, 2.0], "label": [1, 0, 1]}
            if mt != "FFM":
                b["feat"] = [-1, 100]
            o = CpuModel("oracle", mt, 50, 4, 4)
            o.set_state(st0)
            lg, _ = m.train(**b)
            lo, _ = (o.train_batch_csr if mode == "batch" else o.train_csr)(**b)
            assert_close(lg, lo, RTOL, 1e-6, "logits")
            assert_state_close(m.get_state(), o.get_state(), name=f"{mt}/{mode}")


def test_remove_out_range_like_reference_test_model():
    """tests/test_model.cpp:27-29,46-48: LR keeps 1 of {(1,-1,3),(1,0,1),(1,100,0)} at n_feats=50;
    FFM with field 44 >= n_fields drops everything"""
    m = gpu_model("LR", 50, 4, 4)
    w = np.zeros(50, np.float32)
    w[0] = 0.5
    m.set_state({"bias": np.zeros(3, np.float32), "lin_w": w})
    out, _ = m.predict([0, 3], [1, 1, 1], [-1, 0, 100], [3.0, 1.0, 0.0])
    assert out[0] == 0.5
    m = gpu_model("FFM", 50, 4, 4)
    m.set_state({"bias": np.array([0.125, 0, 0], np.float32)})
    out, _ = m.predict([0, 3], [1, 44, 1], [-1, 0, 100], [3.0, 1.0, 0.0])
    assert out[0] == 0.125


def test_wide_samples_beyond_shared_cache():
    """samples with more than 128 features (FFM kernels cache 128 per sample in shared memory)"""
    rng = np.random.default_rng(9)
    nf, nfl, k = 400, 150, 2
    m = gpu_model("FFM", nf, nfl, k)
    o = CpuModel("oracle", "FFM", nf, nfl, k)
    st = pkg.synth.random_state(rng, nf, o.row_len)
    m.set_state(st)
    o.set_state(st)
    b = pkg.synth.random_csr(rng, 6, nf, nfl, max_nnz=140, min_nnz=130, oob_frac=0.02)
    got, _ = m.train(**b)
    want, _ = o.train_batch_csr(**b)
    assert_close(got, want, RTOL, 5e-6, "logits")
    assert_state_close(m.get_state(), o.get_state(), rtol=2e-5, atol=2e-6, atol_z=2e-4, name="wide")


@pytest.mark.parametrize("mt,k", [("FFM", 2), ("FM", 3), ("LR", 1), ("FM", 1100)])
def test_sequential_mode_has_no_sample_width_cap(mt, k):
    """the reference trains samples of any width (ffm.cpp:57-70, fm.cpp:40-67): beyond the 96 valid features (or
    1024 FM factors) the sequential kernel keeps in shared memory, a serial path takes over -- same trajectory"""
    rng = np.random.default_rng(21)
    nf, nfl = 500, 160
    m = gpu_model(mt, nf, nfl, k, mode="sequential")
    o = CpuModel("oracle", mt, nf, nfl, k)
    st = pkg.synth.random_state(rng, nf, o.row_len)
    m.set_state(st)
    o.set_state(st)
    wide = mt != "FM" or k < 1000
    b = pkg.synth.random_csr(rng, 5, nf, nfl, max_nnz=150 if wide else 12, min_nnz=100 if wide else 4, oob_frac=0.02)
    got, gl = m.train(**b)
    want, wl = o.train_csr(**b)
    assert_close(got, want, RTOL, 1e-6, "logits")
    assert abs(gl - wl) <= 1e-9 * max(1.0, abs(wl))
    assert_state_close(m.get_state(), o.get_state(), RTOL, name=f"wide-seq-{mt}")


def test_async_pipeline_three_batches_in_flight():
    rng = np.random.default_rng(4)
    nf, nfl, k = 300, 6, 4
    m = gpu_model("FFM", nf, nfl, k)
    o = CpuModel("oracle", "FFM", nf, nfl, k)
    st = pkg.synth.random_state(rng, nf, o.row_len)
    m.set_state(st)
    o.set_state(st)
    pend = []
    for it in range(7):
        b = pkg.synth.random_csr(rng, 100 + it, nf, nfl)
        lg, loss = m.train(**b, sync=False)
        pend.append((lg, loss, o.train_batch_csr(**b)))
    m.sync()
    for lg, loss, (want, wl) in pend[-3:]:
        assert_close(lg, want, RTOL, 2e-6, "logits")
        assert abs(float(loss[0]) - wl) <= 1e-6 * max(1.0, abs(wl))
    assert_state_close(m.get_state(), o.get_state(), rtol=5e-5, atol=5e-6, atol_z=5e-4, name="async")


@pytest.mark.skipif(not have_ref(), reason="prebuilt oracle/_ref did not travel")
def test_against_the_reference_binary_itself():
    """same check as the golden trajectory, against the reference .so executing on the box's CPU"""
    rng = np.random.default_rng(77)
    nf, nfl, k = 120, 7, 4
    m = gpu_model("FFM", nf, nfl, k, mode="sequential")
    r = CpuModel("ref", "FFM", nf, nfl, k)
    st = pkg.synth.random_state(rng, nf, r.row_len)
    m.set_state(st)
    r.set_state(st)
    b = pkg.synth.random_csr(rng, 300, nf, nfl)
    got, _ = m.train(**b)
    want, _ = r.train_csr(**b)
    assert_close(got, want, RTOL, 1e-6, "logits")
    assert_state_close(m.get_state(), r.get_state(), RTOL, name="ref")


# ---------------------------------------------------------------------------------------------
# 6. device AUC (ftrl_eval_auc), Gaussian init (tests/test_utils.cpp:26-38), IEEE flavour of the kernels
# ---------------------------------------------------------------------------------------------
def test_device_auc_matches_rank_sum_with_ties():
    rng = np.random.default_rng(12)
    m = gpu_model("LR", 16, 1, 1)
    for n, levels in ((1, 0), (2, 0), (1000, 0), (100_000, 0), (100_000, 17), (50_000, 1)):
        s = rng.normal(0, 2, n).astype(np.float32)
        if levels:  # heavy ties (cold-start models score every sample the same)
            s = np.round(s * levels / 4).astype(np.float32) / levels
        s[: n // 50] = -s[: n // 50]
        if n > 10:
            s[:5] = [0.0, -0.0, np.float32(1e-45), -np.float32(1e-45), 0.0]
        y = (rng.random(n) < 0.3).astype(np.int32)
        got, want = m.auc(s, y), pkg.synth.auc(y, s)
        if np.isnan(want):
            assert np.isnan(got)
        else:
            assert abs(got - want) <= 1e-12, (n, levels, got, want)
    assert np.isnan(m.auc(np.zeros(4, np.float32), np.ones(4, np.int32)))  # one class only


def test_gaussian_init_distribution_and_reference_shape_test():
    """tests/test_utils.cpp:26-38 (init_weights(100, 10, 4, 0, 0.01): every row has a value in (-0.05, 0.05)) plus
    the moments of the device generator; n = z = 0"""
    m = gpu_model("FFM", 100, 10, 4, init_mean=0.0, init_stddev=0.01)
    st = m.get_state()
    assert st["vec_w"].shape == (100, 40)
    assert all(((row > -0.05) & (row < 0.05)).any() for row in st["vec_w"])
    assert not st["vec_n"].any() and not st["vec_z"].any() and not st["lin_n"].any() and not st["lin_z"].any()
    m.close()
    m = gpu_model("FFM", 20000, 8, 8, init_mean=0.25, init_stddev=0.02, seed=5)
    w = m.vec_w.astype(np.float64).ravel()
    lw = m.lin_w.astype(np.float64)
    n = w.size
    assert abs(w.mean() - 0.25) < 5 * 0.02 / np.sqrt(n) and abs(w.std() - 0.02) < 5 * 0.02 / np.sqrt(2 * n)
    assert abs(lw.mean() - 0.25) < 5 * 0.02 / np.sqrt(lw.size)
    zs = (w - 0.25) / 0.02
    assert abs((zs ** 3).mean()) < 0.02 and abs((zs ** 4).mean() - 3.0) < 0.05   # skewness, kurtosis
    assert 0.6826 - 0.005 < (np.abs(zs) < 1).mean() < 0.6826 + 0.005
    # another seed gives another model; the same seed the same model
    m2 = gpu_model("FFM", 20000, 8, 8, init_mean=0.25, init_stddev=0.02, seed=6)
    m3 = gpu_model("FFM", 20000, 8, 8, init_mean=0.25, init_stddev=0.02, seed=5)
    assert not np.array_equal(m2.lin_w, m.lin_w) and np.array_equal(m3.vec_w, m.vec_w)


@pytest.mark.parametrize("mt,nfl,k,kw", [("FFM", 6, 4, {}), ("FFM", 39, 8, {"max_nnz": 39}), ("FFM", 4, 8, {"dup_field": True}),
                                         ("FM", 1, 16, {"dup_feat": True, "max_nnz": 20}),
                                         ("LR", 1, 1, {"dup_feat": True, "max_nnz": 20})])
def test_minibatch_ieee_flavour(mt, nfl, k, kw, monkeypatch):
    """FTRL_B200_PRECISE=1: the same minibatch kernels instantiated with IEEE sqrt / division"""
    monkeypatch.setenv("FTRL_B200_PRECISE", "1")
    rng = np.random.default_rng(70 + nfl + k)
    nf = 300
    m = gpu_model(mt, nf, nfl, k)
    o = CpuModel("oracle", mt, nf, nfl, k)
    st = pkg.synth.random_state(rng, nf, o.row_len)
    m.set_state(st)
    o.set_state(st)
    for it, nrows in enumerate((5, 300, 2000)):
        b = pkg.synth.random_csr(rng, nrows, nf, nfl, **kw)
        got, gl = m.train(**b)
        want, wl = o.train_batch_csr(**b)
        assert_close(got, want, RTOL, 2e-6, f"logits batch {it}")
        assert_state_close(m.get_state(), o.get_state(), rtol=2e-5, atol=2e-6, atol_z=2e-4, name=f"{mt} precise {it}")
