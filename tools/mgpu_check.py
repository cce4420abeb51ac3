"""Multi-process / multi-GPU check of the feature-sharded path (run under torchrun, one rank per GPU):
the model trained by G ranks on G shares of a global minibatch must equal the model one GPU trains on the
whole minibatch.  Prints `MGPU_CHECK OK` on rank 0."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import ftrl_ffm_b200 as pkg


def big(rank, world, local, n_feats):
    """size-independent properties of one sharded step at BASELINE.json configs[4] scale (100M features, k 8, 64K
    samples per GPU; the tables do not fit one GPU, so there is no single-GPU model to compare with): the loss the
    library returns is the sum of eval/loss.h over the logits it returns; rows no sample touches keep every bit;
    n never decreases; touched rows move.  Prints `MGPU_BIG OK` on rank 0."""
    nfl, k, B = 39, 8, 65536
    m = pkg.FtrlModel("FFM", n_feats=n_feats, n_fields=nfl, n_factors=k, device=local, rank=rank, world_size=world,
                      max_batch_rows=B, max_batch_nnz=B * nfl)
    blob = torch.frombuffer(bytearray(m.export_peer_blob()), dtype=torch.uint8).cuda()
    blobs = [torch.zeros_like(blob) for _ in range(world)]
    dist.all_gather(blobs, blob)
    m.attach_peers([t.cpu().numpy().tobytes() for t in blobs])
    m.randomize_state(seed=11)
    parts = [pkg.synth.criteo_batch(B, nfl, n_feats, seed=500 + r, dist="zipf") for r in range(world)]
    touched_ids = np.unique(np.concatenate([p["feat"] for p in parts]))
    mine = touched_ids[touched_ids % world == rank] // world          # local rows some sample touches
    per = n_feats // nfl
    # local row ranges: the hot head of field 0, somewhere in the middle, the cold tail of the last field
    ranges = [(0, 2048), (int(mine[len(mine) // 2]) - 1000, 2048), (m.n_local - 2048, 2048)]
    ranges = [(max(0, min(r0, m.n_local - n)), n) for r0, n in ranges]

    def snap():
        return [[m.get_rows(which, r0, n) for which in (0, 1, 2)] for r0, n in ranges]
    before = snap()
    dist.barrier()
    logits, loss = m.train(**parts[rank])
    dist.barrier()
    after = snap()
    ok = bool(np.isfinite(logits).all())
    sg = 1.0 / (1.0 + np.exp(-logits.astype(np.float64)))
    y = parts[rank]["label"]
    want = float(np.sum(-y * np.log(sg) - (1 - y) * np.log(1 - sg)))
    ok = ok and abs(loss - want) <= 1e-9 * abs(want)
    moved = 0
    for (r0, n), bf, af in zip(ranges, before, after):
        hit = np.isin(np.arange(r0, r0 + n), mine)
        for which in (0, 1, 2):
            (lb, vb), (la, va) = bf[which], af[which]
            ok = ok and np.array_equal(lb[~hit], la[~hit]) and np.array_equal(vb[~hit], va[~hit])
            if which == 1:
                ok = ok and bool((la >= lb).all()) and bool((va >= vb).all())
                moved += int((la[hit] > lb[hit]).sum())
    ok = ok and moved > 0
    t = torch.tensor([1.0 if ok else 0.0, float(moved)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    print(f"rank {rank}: n_local {m.n_local} loss/sample {loss / B:.6f} (recomputed {want / B:.6f}) "
          f"touched local rows {len(mine)} checked-and-moved {moved} ok {ok}", flush=True)
    if rank == 0:
        print("MGPU_BIG OK" if t[0].item() == 1.0 else "MGPU_BIG FAILED", flush=True)
    m.close()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if len(sys.argv) > 1 and sys.argv[1] == "--big":
        big(rank, world, local, int(sys.argv[2]) if len(sys.argv) > 2 else 12_500_000 * world)
        dist.destroy_process_group()
        return
    nf, nfl, k, B = 20000, 39, 8, 2048
    kw = dict(model_type="FFM", n_feats=nf, n_fields=nfl, n_factors=k)
    m = pkg.FtrlModel(device=local, rank=rank, world_size=world, max_batch_rows=B, max_batch_nnz=B * nfl, **kw)
    blob = torch.frombuffer(bytearray(m.export_peer_blob()), dtype=torch.uint8).cuda()
    blobs = [torch.zeros_like(blob) for _ in range(world)]
    dist.all_gather(blobs, blob)
    m.attach_peers([t.cpu().numpy().tobytes() for t in blobs])
    dist.barrier()
    rng = np.random.default_rng(3)
    st = pkg.synth.random_state(rng, nf, nfl * k)          # same on every rank (same seed)
    m.set_state(pkg.shard_state(st, world, rank))
    single = pkg.FtrlModel(device=local, **kw) if rank == 0 else None
    if single:
        single.set_state(st)
    dist.barrier()
    ok = True
    for step in range(4):
        parts = [pkg.synth.criteo_batch(B, nfl, nf, seed=10 * step + r, dist="zipf" if step % 2 else "uniform")
                 for r in range(world)]
        lg, loss = m.train(**parts[rank])
        if single:
            glob = {"row_ptr": np.concatenate([[0]] + [p["row_ptr"][1:] + i * B * nfl for i, p in enumerate(parts)]),
                    "field": np.concatenate([p["field"] for p in parts]),
                    "feat": np.concatenate([p["feat"] for p in parts]),
                    "val": np.concatenate([p["val"] for p in parts]),
                    "label": np.concatenate([p["label"] for p in parts])}
            lg1, loss1 = single.train(**glob)
            err = np.max(np.abs(lg - lg1[:B]) / np.maximum(1.0, np.abs(lg1[:B])))
            ok = ok and err <= 1e-5
            print(f"step {step}: rank-0 logits max rel err vs single GPU {err:.2e}", flush=True)
        dist.barrier()
    mine = m.get_state()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        full = pkg.merge_states(gathered)
        ref = single.get_state()
        for key in ref:
            d = np.max(np.abs(full[key].astype(np.float64) - ref[key]) / np.maximum(1.0, np.abs(ref[key])))
            print(f"state {key}: max rel diff {d:.2e}", flush=True)
            ok = ok and d <= 5e-5
        print("MGPU_CHECK OK" if ok else "MGPU_CHECK FAILED", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
