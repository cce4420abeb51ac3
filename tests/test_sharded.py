"""Feature-sharded multi-GPU path.

CPU part (gloo, world_size 2): the host-side plumbing a launcher needs -- exchanging fixed-size peer blobs
with an all-gather and the shard <-> full-table index mapping.
GPU part: G logical shards inside one process on one GPU must give the same model as a single GPU
(shard-count invariance)."""
import os
import socket

import numpy as np
import pytest

from conftest import assert_close, assert_state_close
import ftrl_ffm_b200 as pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # what bench.py does with the blobs of ftrl_export_peer_blob
    blob = torch.full((pkg.binding.PEER_BLOB_BYTES,), rank + 1, dtype=torch.uint8)
    out = [torch.zeros_like(blob) for _ in range(world)]
    dist.all_gather(out, blob)
    ok = all(int(o[0]) == r + 1 and int(o[-1]) == r + 1 for r, o in enumerate(out))
    # per-rank share of a global batch + shard mapping
    rng = np.random.default_rng(0)
    st = pkg.synth.random_state(rng, 37, 6)
    mine = pkg.shard_state(st, world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: v.tolist() for k, v in mine.items()})
    merged = pkg.merge_states([{k: np.asarray(v, np.float32) for k, v in g.items()} for g in gathered])
    ok = ok and all(np.array_equal(merged[k], st[k]) for k in st)
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and float(t) == float(world)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gloo_world2_blob_exchange_and_shard_mapping():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_state_round_trip():
    rng = np.random.default_rng(1)
    st = pkg.synth.random_state(rng, 41, 8)
    for g in (1, 2, 4, 8):
        parts = [pkg.shard_state(st, g, r) for r in range(g)]
        assert sum(len(p["lin_w"]) for p in parts) == 41
        merged = pkg.merge_states(parts)
        for k in st:
            assert np.array_equal(merged[k], st[k])


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_shard_count_invariance_on_one_gpu(world):
    rng = np.random.default_rng(5)
    nf, nfl, k, B = 3000, 13, 8, 512
    kw = dict(model_type="FFM", n_feats=nf, n_fields=nfl, n_factors=k)
    single = pkg.FtrlModel(**kw)
    sharded = pkg.LogicalShards(world, max_batch_rows=B, max_batch_nnz=B * nfl, **kw)
    st = pkg.synth.random_state(rng, nf, nfl * k)
    single.set_state(st)
    sharded.set_state(st)
    for step in range(3):
        parts = [pkg.synth.criteo_batch(B, nfl, nf, seed=100 * step + r, dist="zipf" if step % 2 else "uniform")
                 for r in range(world)]
        # the single GPU sees the concatenation of all ranks' samples as ONE minibatch
        glob = {"row_ptr": np.concatenate([[0]] + [p["row_ptr"][1:] + i * B * nfl for i, p in enumerate(parts)]),
                "field": np.concatenate([p["field"] for p in parts]), "feat": np.concatenate([p["feat"] for p in parts]),
                "val": np.concatenate([p["val"] for p in parts]), "label": np.concatenate([p["label"] for p in parts])}
        lg1, loss1 = single.train(**glob)
        outs = sharded.train(parts)
        lgG = np.concatenate([o[0] for o in outs])
        assert_close(lgG, lg1, 1e-5, 2e-6, f"logits step {step}")
        assert abs(sum(o[1] for o in outs) - loss1) <= 1e-6 * max(1.0, abs(loss1))
        assert_state_close(sharded.get_state(), single.get_state(), rtol=2e-5, atol=2e-6, atol_z=2e-4,
                           name=f"G={world} step {step}")
    sharded.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("mt,k", [("LR", 1), ("FM", 8), ("FM", 5)])
def test_lr_fm_shard_count_invariance_on_one_gpu(mt, k, world):
    """LR and FM shard like FFM (lr.cpp:9-18, fm.cpp:21-32 train any sample on any model): the owner materialises
    the linear w / latent row of every row the global batch touches and pushes it to the ranks that touch it, the
    ranks reduce their duplicates and send one (sum g, sum g^2) per row back.  Hot rows (Zipf ids: thousands of
    occurrences, several chunks per row) and rows only their owner touches are both in these batches; prediction
    reads every shard."""
    rng = np.random.default_rng(8)
    nf, nfl, B = 3000, 13, 700
    kw = dict(model_type=mt, n_feats=nf, n_fields=nfl, n_factors=k)
    single = pkg.FtrlModel(**kw)
    sharded = pkg.LogicalShards(world, max_batch_rows=B, max_batch_nnz=B * nfl, **kw)
    st = pkg.synth.random_state(rng, nf, 0 if mt == "LR" else k)
    single.set_state(st)
    sharded.set_state(st)
    for step in range(3):
        parts = [pkg.synth.criteo_batch(B, nfl, nf, seed=100 * step + r, dist="zipf" if step % 2 == 0 else "uniform")
                 for r in range(world)]
        glob = {"row_ptr": np.concatenate([[0]] + [p["row_ptr"][1:] + i * B * nfl for i, p in enumerate(parts)]),
                "field": np.concatenate([p["field"] for p in parts]), "feat": np.concatenate([p["feat"] for p in parts]),
                "val": np.concatenate([p["val"] for p in parts]), "label": np.concatenate([p["label"] for p in parts])}
        lg1, loss1 = single.train(**glob)
        outs = sharded.train(parts)
        lgG = np.concatenate([o[0] for o in outs])
        assert_close(lgG, lg1, 1e-5, 2e-6, f"logits step {step}")
        assert abs(sum(o[1] for o in outs) - loss1) <= 1e-6 * max(1.0, abs(loss1))
        assert_state_close(sharded.get_state(), single.get_state(), rtol=2e-5, atol=2e-6, atol_z=2e-4,
                           name=f"{mt} G={world} step {step}")
    p = parts[0]
    want, _ = single.predict(p["row_ptr"], p["field"], p["feat"], p["val"], p["label"])
    got, _ = sharded.models[1].predict(p["row_ptr"], p["field"], p["feat"], p["val"], p["label"])
    assert_close(got, want, 1e-5, 2e-6, "predict through the shards")
    sharded.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_generic_fallback_repeated_fields(world):
    """samples that repeat a field (the reference trains any sample: ffm.cpp:90-136): when ANY rank's batch has one,
    every rank takes the generic kernels for that step (device-side flag, no host round trip) and the owners fuse
    nothing; the field masks stay exact, so untouched slices keep their stored w.  Steps with and without such
    samples alternate."""
    rng = np.random.default_rng(12)
    nf, nfl, k, B = 400, 6, 4, 160
    kw = dict(model_type="FFM", n_feats=nf, n_fields=nfl, n_factors=k)
    single = pkg.FtrlModel(**kw)
    sharded = pkg.LogicalShards(world, max_batch_rows=B, max_batch_nnz=B * 8, **kw)
    st = pkg.synth.random_state(rng, nf, nfl * k)
    single.set_state(st)
    sharded.set_state(st)
    for step in range(4):
        dup = step % 2 == 0
        # only ONE rank's share repeats fields in the "dup" steps: the flag has to travel
        parts = [pkg.synth.random_csr(rng, B, nf, nfl, max_nnz=8 if (dup and r == world - 1) else nfl, min_nnz=1,
                                      dup_field=dup and r == world - 1, dup_feat=dup and r == 0) for r in range(world)]
        off = np.cumsum([0] + [int(p["row_ptr"][-1]) for p in parts])
        glob = {"row_ptr": np.concatenate([[0]] + [p["row_ptr"][1:] + off[i] for i, p in enumerate(parts)]),
                "field": np.concatenate([p["field"] for p in parts]), "feat": np.concatenate([p["feat"] for p in parts]),
                "val": np.concatenate([p["val"] for p in parts]), "label": np.concatenate([p["label"] for p in parts])}
        lg1, loss1 = single.train(**glob)
        outs = sharded.train(parts)
        lgG = np.concatenate([o[0] for o in outs])
        assert_close(lgG, lg1, 1e-5, 2e-6, f"logits step {step}")
        assert abs(sum(o[1] for o in outs) - loss1) <= 1e-6 * max(1.0, abs(loss1))
        assert_state_close(sharded.get_state(), single.get_state(), rtol=2e-5, atol=2e-6, atol_z=2e-4,
                           name=f"generic G={world} step {step}")
    sharded.close()


def _ragged(b, rng, nf, drop=0.3, oob=0.02):
    """criteo-shaped batch -> samples with a random SUBSET of the fields (still distinct), a few out-of-range ids"""
    keep = rng.random(len(b["feat"])) >= drop
    rows = np.repeat(np.arange(len(b["label"])), np.diff(b["row_ptr"]))
    cnt = np.bincount(rows[keep], minlength=len(b["label"]))
    feat = b["feat"][keep].copy()
    bad = rng.random(len(feat)) < oob
    feat[bad] = nf + 7            # dropped by the validity mask (ffm.cpp:30-36)
    return {"row_ptr": np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64), "field": b["field"][keep].copy(),
            "feat": feat, "val": b["val"][keep].copy(), "label": b["label"].copy()}


@pytest.mark.gpu
def test_sharded_ragged_samples_and_field_masks():
    """samples that carry different subsets of the fields on different ranks: the owner must materialise w for the
    UNION of the slices the ranks touch, and only for those (cold slices keep their initial w)"""
    rng = np.random.default_rng(11)
    world, nf, nfl, k, B = 2, 600, 9, 4, 256
    kw = dict(model_type="FFM", n_feats=nf, n_fields=nfl, n_factors=k)
    single = pkg.FtrlModel(**kw)
    sharded = pkg.LogicalShards(world, max_batch_rows=B, max_batch_nnz=B * nfl, **kw)
    st = pkg.synth.random_state(rng, nf, nfl * k)
    single.set_state(st)
    sharded.set_state(st)
    for step in range(3):
        parts = [_ragged(pkg.synth.criteo_batch(B, nfl, nf, seed=50 * step + r, dist="zipf"), rng, nf)
                 for r in range(world)]
        off = np.cumsum([0] + [len(p["feat"]) for p in parts])
        glob = {"row_ptr": np.concatenate([[0]] + [p["row_ptr"][1:] + off[i] for i, p in enumerate(parts)]),
                "field": np.concatenate([p["field"] for p in parts]), "feat": np.concatenate([p["feat"] for p in parts]),
                "val": np.concatenate([p["val"] for p in parts]), "label": np.concatenate([p["label"] for p in parts])}
        lg1, loss1 = single.train(**glob)
        outs = sharded.train(parts)
        assert_close(np.concatenate([o[0] for o in outs]), lg1, 1e-5, 2e-6, f"logits step {step}")
        assert_state_close(sharded.get_state(), single.get_state(), rtol=2e-5, atol=2e-6, atol_z=2e-4,
                           name=f"ragged step {step}")
    sharded.close()


@pytest.mark.gpu
def test_sharded_rank_with_an_empty_share():
    """the last step of an epoch may leave some ranks without samples: the collective step must still complete
    and equal one GPU training on the samples that exist"""
    rng = np.random.default_rng(12)
    world, nf, nfl, k, B = 2, 500, 7, 4, 128
    kw = dict(model_type="FFM", n_feats=nf, n_fields=nfl, n_factors=k)
    single = pkg.FtrlModel(**kw)
    sharded = pkg.LogicalShards(world, max_batch_rows=B, max_batch_nnz=B * nfl, **kw)
    st = pkg.synth.random_state(rng, nf, nfl * k)
    single.set_state(st)
    sharded.set_state(st)
    empty = {"row_ptr": np.zeros(1, np.int64), "field": np.zeros(0, np.int32), "feat": np.zeros(0, np.int32),
             "val": np.zeros(0, np.float32), "label": np.zeros(0, np.int32)}
    for step in range(2):
        full = pkg.synth.criteo_batch(B, nfl, nf, seed=70 + step, dist="zipf")
        parts = [full, empty] if step == 0 else [empty, full]
        lg1, loss1 = single.train(**full)
        outs = sharded.train(parts)
        lg = outs[0][0] if step == 0 else outs[1][0]
        assert_close(lg, lg1, 1e-5, 2e-6, f"logits step {step}")
        assert abs(sum(o[1] for o in outs) - loss1) <= 1e-6 * max(1.0, abs(loss1))
        assert_state_close(sharded.get_state(), single.get_state(), rtol=2e-5, atol=2e-6, atol_z=2e-4,
                           name=f"empty share step {step}")
    sharded.close()


@pytest.mark.gpu
def test_sharded_model_file_is_the_single_model_file(tmp_path):
    """save on ONE rank of a sharded run (peer reads), load on EVERY rank: the file is the reference's single-model
    layout (ffm.cpp:138-159), byte-identical to what one GPU holding the same weights writes, and the reference's
    own load_compressed_model reads it back"""
    from oracle.cpu_model import CpuModel, have_ref
    rng = np.random.default_rng(13)
    world, nf, nfl, k, B = 4, 997, 6, 4, 128   # 997 rows: the shards differ in size
    kw = dict(model_type="FFM", n_feats=nf, n_fields=nfl, n_factors=k)
    single = pkg.FtrlModel(**kw)
    sharded = pkg.LogicalShards(world, max_batch_rows=B, max_batch_nnz=B * nfl, **kw)
    st = pkg.synth.random_state(rng, nf, nfl * k)
    single.set_state(st)
    sharded.set_state(st)
    parts = [pkg.synth.criteo_batch(B, nfl, nf, seed=90 + r, dist="zipf") for r in range(world)]
    sharded.train(parts)
    f_sh, f_one = str(tmp_path / "sharded.zst"), str(tmp_path / "single.zst")
    sharded.models[1].save_compressed_model(f_sh)         # any one rank
    full = sharded.get_state()
    single.set_state({"bias": full["bias"], "lin_w": full["lin_w"], "vec_w": full["vec_w"]})
    single.save_compressed_model(f_one)
    assert open(f_sh, "rb").read() == open(f_one, "rb").read()
    # load on every rank of a fresh sharded run
    again = pkg.LogicalShards(world, max_batch_rows=B, max_batch_nnz=B * nfl, **kw)
    for m in again.models:
        m.load_compressed_model(f_sh)
    got = again.get_state()
    assert np.array_equal(got["lin_w"], full["lin_w"]) and np.array_equal(got["vec_w"], full["vec_w"])
    assert got["bias"][0] == full["bias"][0]
    # text form
    f_txt = str(tmp_path / "sharded.txt")
    sharded.models[0].save_model(f_txt)
    for m in again.models:
        m.set_state({"lin_w": np.zeros(m.n_local, np.float32)})
        m.load_model(f_txt)
    assert_close(again.get_state()["lin_w"], full["lin_w"], 1e-5, 1e-7, "text lin_w")
    if have_ref():
        r = CpuModel("ref", "FFM", nf, nfl, k)
        r._fn("load_compressed")(r.h, f_sh.encode())
        rs = r.get_state()
        assert np.array_equal(rs["lin_w"], full["lin_w"]) and np.array_equal(rs["vec_w"], full["vec_w"])
    sharded.close()
    again.close()
