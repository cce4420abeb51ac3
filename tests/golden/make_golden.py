"""Generates the golden fixtures in this directory from the REFERENCE ITSELF (oracle/_ref, compiled from
/root/reference by oracle/Makefile).  Run in the build container (where /root/reference exists):

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures travel to the GPU box; nothing at test time reads /root/reference.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.cpu_model import CpuModel, scalar_fns  # noqa: E402
import ftrl_ffm_b200 as pkg  # noqa: E402

REF_DATA = "/root/reference/data"


def f32(x):
    return float(np.float32(x))


def parse_text(path, libffm):
    row_ptr, field, feat, val, label = [0], [], [], [], []
    with open(path) as f:
        for line in f:
            toks = line.split()
            if not toks:
                continue
            label.append(1 if int(toks[0]) > 0 else 0)
            for t in toks[1:]:
                p = t.split(":")
                if libffm:
                    fl, ft, v = int(p[0]), int(p[1]), float(p[2])
                else:
                    fl, ft, v = 0, int(p[0]), float(p[1])
                if np.float32(v) != 0:
                    field.append(fl)
                    feat.append(ft)
                    val.append(v)
            row_ptr.append(len(feat))
    return {"row_ptr": np.asarray(row_ptr, np.int64), "field": np.asarray(field, np.int32),
            "feat": np.asarray(feat, np.int32), "val": np.asarray(val, np.float32),
            "label": np.asarray(label, np.int32)}


def scalars():
    fn = scalar_fns("ref")
    m = CpuModel("ref", "LR", 4)
    xs = [-30.0, -2.0, -0.5, 0.0, 1e-3, 1.0, 2.0, 17.5]
    out = {
        "sgn": [[x, f32(fn["sgn"](x))] for x in [1.0, 0.0, -2.0, 1e-30]],
        "sigmoid": [[x, f32(fn["sigmoid"](x))] for x in xs],
        "loss": [[y, x, float(fn["loss"](y, x))] for y in (0, 1) for x in xs],
        "weight": [[n, z, f32(m.weight(n, z))] for n in (0.0, 0.25, 1.0, 7.5)
                   for z in (-300.0, -0.1000001, -0.1, 0.0, 0.05, 0.1, 0.2, 12.0)],
        "hyper": {"w_alpha": 1e-4, "w_beta": 1.0, "w_l1": 0.1, "w_l2": 5.0},
    }
    return out


def appendix_b():
    out = {}
    m = CpuModel("ref", "LR", 8)
    steps = []
    for _ in range(3):
        lg = m.train([0, 0], [3, 5], [1.0, 0.5], 1)
        st = m.get_state()
        steps.append({"logit": f32(lg),
                      "feat3": [f32(st["lin_w"][3]), f32(st["lin_z"][3]), f32(st["lin_n"][3])],
                      "feat5": [f32(st["lin_w"][5]), f32(st["lin_z"][5]), f32(st["lin_n"][5])],
                      "bias": [f32(v) for v in st["bias"]]})
    out["lr_steps"] = steps
    m = CpuModel("ref", "FFM", 8, 3, 2)
    st = m.get_state()
    st["lin_w"] = np.array([0.01 * (i + 1) for i in range(8)], np.float32)
    st["vec_w"] = np.array([[0.1 * (i + 1) - 0.05 * j for j in range(6)] for i in range(8)], np.float32)
    st["bias"] = np.array([0.25, 0, 0], np.float32)
    m.set_state(st)
    out["ffm_forward"] = {"logit": f32(m.predict([0, 1, 2], [1, 4, 6], [1.0, 0.5, 2.0])),
                          "prob": f32(m.predict([0, 1, 2], [1, 4, 6], [1.0, 0.5, 2.0], True))}
    m = CpuModel("ref", "FM", 8, 1, 2)
    st = m.get_state()
    st["lin_w"] = np.array([0.01 * (i + 1) for i in range(8)], np.float32)
    st["vec_w"] = np.array([[0.1 * (i + 1) - 0.05 * j for j in range(2)] for i in range(8)], np.float32)
    st["bias"] = np.array([0.25, 0, 0], np.float32)
    m.set_state(st)
    out["fm_forward"] = {"logit": f32(m.predict([0, 0, 0], [1, 4, 6], [1.0, 0.5, 2.0]))}
    return out


def trajectory(model_type, n_feats, n_fields, k, n_rows, seed, **kw):
    """random live state + a ragged CSR block trained sequentially by the reference"""
    rng = np.random.default_rng(seed)
    m = CpuModel("ref", model_type, n_feats, n_fields, k)
    st0 = pkg.synth.random_state(rng, n_feats, m.row_len)
    m.set_state(st0)
    # the reference FFM self-deadlocks on a repeated id inside one sample: keep ids distinct for it
    b = pkg.synth.random_csr(rng, n_rows, n_feats, n_fields, dup_feat=(model_type != "FFM"), **kw)
    logits, loss = m.train_csr(**b)
    st1 = m.get_state()
    pred, ploss = m.predict_csr(b["row_ptr"], b["field"], b["feat"], b["val"], b["label"])
    out = {"n_feats": n_feats, "n_fields": n_fields, "k": k, "logits": logits, "loss": loss,
           "pred": pred, "pred_loss": ploss}
    out.update({"b_" + key: v for key, v in b.items()})
    out.update({"s0_" + key: v for key, v in st0.items()})
    out.update({"s1_" + key: v for key, v in st1.items()})
    return out


def cfg1():
    ffm = parse_text(os.path.join(REF_DATA, "libffm_data.txt"), True)
    svm = parse_text(os.path.join(REF_DATA, "libsvm_data.txt"), False)
    assert np.array_equal(ffm["feat"], svm["feat"]) and np.array_equal(ffm["label"], svm["label"])
    out = {"row_ptr": ffm["row_ptr"], "field": ffm["field"].astype(np.int8), "feat": ffm["feat"],
           "val": ffm["val"], "label": ffm["label"].astype(np.int8)}
    for mt in ("LR", "FM", "FFM"):
        data = ffm if mt == "FFM" else svm
        m = CpuModel("ref", mt, 10000, 8, 16)
        st = m.get_state()  # fast_init: w = 0 (never-touched weights are invisible to these samples? no:
        # untouched latent slices are never read either, so metrics do not depend on the init)
        tr, ev, au = [], [], []
        for _ in range(5):
            _, ls = m.train_csr(**data)
            tr.append(ls / 10000)
            pred, pl = m.predict_csr(data["row_ptr"], data["field"], data["feat"], data["val"], data["label"])
            ev.append(pl / 10000)
            au.append(pkg.synth.auc(data["label"], pred))
        out[mt + "_train_loss"] = np.asarray(tr)
        out[mt + "_eval_loss"] = np.asarray(ev)
        out[mt + "_auc"] = np.asarray(au)
        st = m.get_state()
        out[mt + "_final_lin_z"] = st["lin_z"]
        out[mt + "_final_lin_n"] = st["lin_n"]
        out[mt + "_final_bias"] = st["bias"]
    return out


def model_files():
    """files written by the reference's own save_* (lr.cpp:26-31, ffm.cpp:138-146, :161-174)"""
    rng = np.random.default_rng(7)
    out = {}
    m = CpuModel("ref", "LR", 50)
    st = {"bias": np.array([0.125, 0, 0], np.float32), "lin_w": rng.normal(0, 0.02, 50).astype(np.float32)}
    m.set_state(st)
    path = os.path.join(HERE, "ref_lr.zst")
    m._fn("save_compressed")(m.h, path.encode(), 10)
    out["lr_bias"], out["lr_lin_w"] = st["bias"][:1], st["lin_w"]
    m = CpuModel("ref", "FFM", 50, 4, 4)
    st = {"bias": np.array([-0.75, 0, 0], np.float32), "lin_w": rng.normal(0, 0.02, 50).astype(np.float32),
          "vec_w": rng.normal(0, 0.02, (50, 16)).astype(np.float32)}
    st["vec_w"][3, 5] = 0.0
    st["vec_w"][4, 1] = 1.5e-7
    st["vec_w"][5, 2] = 123456.0
    m.set_state(st)
    m._fn("save_compressed")(m.h, os.path.join(HERE, "ref_ffm.zst").encode(), 10)
    m._fn("save_text")(m.h, os.path.join(HERE, "ref_ffm.txt").encode())
    out["ffm_bias"], out["ffm_lin_w"], out["ffm_vec_w"] = st["bias"][:1], st["lin_w"], st["vec_w"]
    np.savez_compressed(os.path.join(HERE, "model_files.npz"), **out)


def main():
    with open(os.path.join(HERE, "scalars.json"), "w") as f:
        json.dump(scalars(), f, indent=1)
    with open(os.path.join(HERE, "appendix_b.json"), "w") as f:
        json.dump(appendix_b(), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "traj_lr.npz"), **trajectory("LR", 60, 1, 1, 400, 11, max_nnz=9))
    np.savez_compressed(os.path.join(HERE, "traj_fm.npz"), **trajectory("FM", 60, 1, 8, 300, 12, max_nnz=9))
    np.savez_compressed(os.path.join(HERE, "traj_fm_k5.npz"), **trajectory("FM", 40, 1, 5, 120, 15, max_nnz=7))
    np.savez_compressed(os.path.join(HERE, "traj_ffm.npz"), **trajectory("FFM", 80, 6, 4, 300, 13, max_nnz=6))
    np.savez_compressed(os.path.join(HERE, "traj_ffm_dupfield.npz"),
                        **trajectory("FFM", 80, 4, 8, 200, 14, max_nnz=7, dup_field=True))
    np.savez_compressed(os.path.join(HERE, "traj_ffm_k3.npz"), **trajectory("FFM", 50, 5, 3, 150, 16, max_nnz=5))
    np.savez_compressed(os.path.join(HERE, "cfg1.npz"), **cfg1())
    model_files()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
