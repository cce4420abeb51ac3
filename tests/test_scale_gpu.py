"""GPU tests at BASELINE.json sizes: size-independent properties of the minibatch path (the oracle only
finishes in seconds on sub-samples, which are checked too)."""
import os

import numpy as np
import pytest

from conftest import assert_close, assert_state_close
from oracle.cpu_model import CpuModel
import ftrl_ffm_b200 as pkg

pytestmark = pytest.mark.gpu

NFL = 39


def rows_snapshot(m, ranges):
    out = []
    for r0, n in ranges:
        for which in (0, 1, 2):
            lin, vec = m.get_rows(which, r0, n)
            out += [lin, vec]
    return out


def make(n_feats, k, env=None, **kw):
    old = {}
    for key, v in (env or {}).items():
        old[key] = os.environ.get(key)
        os.environ[key] = v
    try:
        m = pkg.FtrlModel("FFM", n_feats=n_feats, n_fields=NFL, n_factors=k, **kw)
    finally:
        for key, v in old.items():
            if v is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = v
    m.randomize_state(seed=11)
    return m


@pytest.mark.parametrize("dist", ["zipf", "uniform"])
def test_cfg3_index_pipeline_counts_match_numpy(dist):
    """config 3 shape (39 fields, 1M features, k 4, 64K samples): the sort/segment pipeline's view of the batch
    (distinct rows, rows finalised per sample) must equal a numpy count of the same ids"""
    b = pkg.synth.criteo_batch(65536, NFL, 1_000_000, seed=5, dist=dist)
    m = make(1_000_000, 4)
    _, loss = m.train(**b, want_logits=False)
    st = m.last_batch_stats()
    ids, cnt = np.unique(b["feat"], return_counts=True)
    assert st["nnz_valid"] == len(b["feat"])
    assert st["n_unique"] == len(ids)
    assert st["n_fused_rows"] == int((cnt == 1).sum())
    assert np.isfinite(loss)


def test_cfg3_deterministic_and_tile_equals_generic_path():
    b = pkg.synth.criteo_batch(65536, NFL, 1_000_000, seed=6, dist="zipf")
    ids = np.unique(b["feat"])
    ranges = [(0, 3000), (int(ids[len(ids) // 2]) - 50, 2000), (999_000, 1000)]
    snaps = []
    for env in (None, None, {"FTRL_B200_TILE": "0"}):
        m = make(1_000_000, 4, env)
        logits, loss = m.train(**b)
        snaps.append((logits, loss, rows_snapshot(m, ranges)))
        m.close()
    # same kernels, same input: bit-identical (no atomics on the data path)
    assert np.array_equal(snaps[0][0], snaps[1][0]) and snaps[0][1] == snaps[1][1]
    for a, c in zip(snaps[0][2], snaps[1][2]):
        assert np.array_equal(a, c)
    # tile kernels vs the generic LDG kernels: same semantics, different summation orders
    assert_close(snaps[2][0], snaps[0][0], 1e-5, 2e-6, "logits tile vs generic")
    for a, c in zip(snaps[0][2], snaps[2][2]):
        assert_close(a, c, 2e-5, 2e-4, "state tile vs generic")


def test_cfg4_properties_one_step():
    """config 4 shape (39 fields, 10M features, k 8, 64K samples, tables 37 GB):
    loss consistency, untouched rows unchanged, n never decreases, sample-order invariance"""
    n_feats = 10_000_000
    b = pkg.synth.criteo_batch(65536, NFL, n_feats, seed=7, dist="zipf")
    present = np.zeros(n_feats, bool)
    present[b["feat"]] = True
    m = make(n_feats, 8)
    ranges = [(0, 2000), (5_000_000, 2000)]
    before = rows_snapshot(m, ranges)
    logits, loss = m.train(**b)
    after = rows_snapshot(m, ranges)
    # the fp64 loss sum the library returns is the sum of eval/loss.h over the returned logits
    s = 1.0 / (1.0 + np.exp(-logits.astype(np.float64)))
    want = float(np.sum(-b["label"] * np.log(s) - (1 - b["label"]) * np.log(1 - s)))
    assert abs(loss - want) <= 1e-9 * abs(want)
    i = 0
    for r0, n in ranges:
        touched = present[r0:r0 + n]
        for which in (0, 1, 2):
            lin_b, vec_b, lin_a, vec_a = before[i], before[i + 1], after[i], after[i + 1]
            i += 2
            assert np.array_equal(lin_b[~touched], lin_a[~touched]) and np.array_equal(vec_b[~touched], vec_a[~touched])
            if which == 1:  # n only grows
                assert (lin_a >= lin_b).all() and (vec_a >= vec_b).all()
                assert (lin_a[touched] > lin_b[touched]).any()
    # a permutation of the samples is the same minibatch
    perm = np.random.default_rng(1).permutation(65536)
    bp = {"row_ptr": b["row_ptr"], "field": b["field"].reshape(-1, NFL)[perm].reshape(-1),
          "feat": b["feat"].reshape(-1, NFL)[perm].reshape(-1), "val": b["val"].reshape(-1, NFL)[perm].reshape(-1),
          "label": b["label"][perm]}
    m2 = make(n_feats, 8)
    logits2, loss2 = m2.train(**bp)
    assert_close(logits2, logits[perm], 1e-5, 2e-6, "logits under permutation")
    assert abs(loss2 - loss) <= 1e-9 * abs(loss)
    after2 = rows_snapshot(m2, ranges)
    for a, c in zip(after, after2):
        assert_close(c, a, 2e-5, 2e-4, "state under permutation")


def test_criteo_shape_subsample_against_oracle():
    """F 39, k 8 (the benchmark's sample shape) on a table the oracle can hold: 4096 samples, live state"""
    rng = np.random.default_rng(8)
    nf, k = 39 * 2000, 8
    m = pkg.FtrlModel("FFM", n_feats=nf, n_fields=NFL, n_factors=k)
    o = CpuModel("oracle", "FFM", nf, NFL, k)
    st = pkg.synth.random_state(rng, nf, NFL * k)
    m.set_state(st)
    o.set_state(st)
    for step in range(2):
        b = pkg.synth.criteo_batch(4096, NFL, nf, seed=20 + step, dist="zipf")
        got, gl = m.train(**b)
        want, wl = o.train_batch_csr(**b)
        assert_close(got, want, 1e-5, 2e-6, "logits")
        assert abs(gl - wl) <= 1e-6 * abs(wl)
    # hot rows collect thousands of fp32 contributions per coordinate (fp64 in the oracle)
    assert_state_close(m.get_state(), o.get_state(), rtol=1e-4, atol=1e-5, atol_z=5e-3, name="criteo subsample")
