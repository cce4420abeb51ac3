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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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
