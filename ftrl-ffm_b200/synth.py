"""Synthetic Criteo-shaped minibatches: every sample has exactly one feature per
field; field f draws its id from the disjoint range [f*per, (f+1)*per), per = n_feats // n_fields (this
also avoids the reference's self-deadlock on a repeated id inside one sample, src/model/ffm.cpp:78);
rank ~ Zipf(s) folded by `mod per` and scattered by a multiplicative permutation (or uniform ids, the
cache-hostile case); the first n_numeric fields carry round(U(0,1),4)+1e-4, the rest 1; labels are
Bernoulli(ctr) or come from a planted logistic model.  Deterministic for a given seed."""
from __future__ import annotations

import math

import numpy as np


def _scatter_mult(per: int) -> int:
    m = 1000003
    while math.gcd(m, per) != 1:
        m += 2
    return m


def criteo_batch(n_rows: int, n_fields: int = 39, n_feats: int = 1_000_000, seed: int = 42, dist: str = "zipf",
                 zipf_s: float = 1.2, n_numeric: int = 13, ctr: float = 0.25, planted: bool = False):
    """Returns a dict of CSR arrays: row_ptr int64[n+1], field/feat int32[nnz], val float32[nnz],
    label int32[n]."""
    rng = np.random.default_rng(seed)
    per = n_feats // n_fields
    if per < 1:
        raise ValueError("n_feats < n_fields")
    if dist == "zipf":
        rank = rng.zipf(zipf_s, size=(n_rows, n_fields)).astype(np.uint64) - 1
        local = (rank % np.uint64(per)) * np.uint64(_scatter_mult(per)) % np.uint64(per)
    elif dist == "uniform":
        local = rng.integers(0, per, size=(n_rows, n_fields)).astype(np.uint64)
    else:
        raise ValueError(dist)
    feat = (local.astype(np.int64) + np.arange(n_fields, dtype=np.int64)[None, :] * per).astype(np.int32)
    val = np.ones((n_rows, n_fields), np.float32)
    nn = min(n_numeric, n_fields)
    if nn:
        val[:, :nn] = (np.round(rng.random((n_rows, nn)), 4) + 1e-4).astype(np.float32)
    if planted:
        # planted logistic model on a hash of the ids: gives a learnable signal for quality parity runs
        wv = ((feat.astype(np.int64) * 2654435761) % 1000) / 1000.0 - 0.5
        logit = (wv * val).sum(axis=1) * 1.5 + math.log(ctr / (1 - ctr))
        p = 1.0 / (1.0 + np.exp(-logit))
        label = (rng.random(n_rows) < p).astype(np.int32)
    else:
        label = (rng.random(n_rows) < ctr).astype(np.int32)
    field = np.broadcast_to(np.arange(n_fields, dtype=np.int32)[None, :], (n_rows, n_fields))
    return {
        "row_ptr": np.arange(0, (n_rows + 1) * n_fields, n_fields, dtype=np.int64),
        "field": np.ascontiguousarray(field).reshape(-1),
        "feat": feat.reshape(-1),
        "val": val.reshape(-1),
        "label": label,
    }


def random_csr(rng, n_rows, n_feats, n_fields, max_nnz=8, min_nnz=0, oob_frac=0.05, dup_field=False,
               dup_feat=False):
    """Small ragged batches for parity tests: random lengths (including empty rows), optional
    out-of-range ids/fields, optional repeated fields / ids inside a sample."""
    row_ptr = [0]
    field, feat, val = [], [], []
    for _ in range(n_rows):
        n = int(rng.integers(min_nnz, max_nnz + 1))
        if dup_field or n > n_fields:
            f = rng.integers(0, n_fields, n)
        else:
            f = rng.permutation(n_fields)[:n]
        if dup_feat or n > n_feats:
            i = rng.integers(0, n_feats, n)
        else:
            i = rng.choice(n_feats, n, replace=False)
        f = f.astype(np.int64)
        i = i.astype(np.int64)
        if n and oob_frac > 0:
            bad = rng.random(n) < oob_frac
            i = np.where(bad, rng.choice([-1, -7, n_feats, n_feats + 5], n), i)
            bad = rng.random(n) < oob_frac
            f = np.where(bad, rng.choice([-1, n_fields, n_fields + 3], n), f)
        v = np.round(rng.normal(1.0, 0.5, n), 3)
        field += list(f)
        feat += list(i)
        val += list(v)
        row_ptr.append(len(feat))
    return {
        "row_ptr": np.asarray(row_ptr, np.int64),
        "field": np.asarray(field, np.int32),
        "feat": np.asarray(feat, np.int32),
        "val": np.asarray(val, np.float32),
        "label": rng.integers(0, 2, n_rows).astype(np.int32),
    }


def random_state(rng, n_feats, row_len, live=True):
    """Non-cold-start model state so latent arithmetic is exercised: |z| above
    l1 for most coordinates, n >= 0.5 so the ffm.cpp:118 term stays finite."""
    st = {
        "bias": np.array([0.0, rng.uniform(0.5, 2.0), rng.normal(0, 50)], np.float32),
        "lin_w": rng.normal(0, 0.02, n_feats).astype(np.float32),
        "lin_n": rng.uniform(0.5, 3.0, n_feats).astype(np.float32),
        "lin_z": (rng.normal(0, 1, n_feats) * (300.0 if live else 0.0)).astype(np.float32),
    }
    if row_len:
        st["vec_w"] = rng.normal(0, 0.02, (n_feats, row_len)).astype(np.float32)
        st["vec_n"] = rng.uniform(0.5, 3.0, (n_feats, row_len)).astype(np.float32)
        st["vec_z"] = (rng.normal(0, 1, (n_feats, row_len)) * (300.0 if live else 0.0)).astype(np.float32)
    return st


def slice_csr(b: dict, r0: int, r1: int) -> dict:
    rp = b["row_ptr"]
    a, e = int(rp[r0]), int(rp[r1])
    return {"row_ptr": (rp[r0:r1 + 1] - a).astype(np.int64), "field": b["field"][a:e], "feat": b["feat"][a:e],
            "val": b["val"][a:e], "label": b["label"][r0:r1]}


def write_text(b: dict, path: str, fmt: str = "libffm") -> None:
    """libffm `label field:feat:val ...` or libsvm `label feat:val ...` (src/data/parser.cpp)."""
    rp = b["row_ptr"]
    with open(path, "w") as f:
        for r in range(len(rp) - 1):
            toks = [str(int(b["label"][r]))]
            for t in range(int(rp[r]), int(rp[r + 1])):
                v = str(np.float32(b["val"][t]))  # shortest text that round-trips in fp32 ("1.0", "0.5489")
                if fmt == "libffm":
                    toks.append(f"{int(b['field'][t])}:{int(b['feat'][t])}:{v}")
                else:
                    toks.append(f"{int(b['feat'][t])}:{v}")
            f.write(" ".join(toks) + "\n")


def auc(labels, scores) -> float:
    """ROC AUC with average ranks for ties (the reference has no AUC code; harness-side metric)."""
    y = np.asarray(labels).astype(bool)
    s = np.asarray(scores, np.float64)
    order = np.argsort(s, kind="mergesort")
    ss = s[order]
    ranks = np.empty(len(s), np.float64)
    i = 0
    n = len(s)
    # average ranks over ties
    boundaries = np.flatnonzero(np.diff(ss)) + 1
    starts = np.concatenate(([0], boundaries))
    ends = np.concatenate((boundaries, [n]))
    avg = (starts + ends - 1) / 2.0 + 1.0
    ranks[order] = np.repeat(avg, ends - starts)
    n_pos = int(y.sum())
    n_neg = n - n_pos
    if n_pos == 0 or n_neg == 0:
        return float("nan")
    return float((ranks[y].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))
