"""G processes (multiprocessing spawn, no torch / NCCL), one GPU each, blobs exchanged through pipes -- isolates the
cost of the cross-process (CUDA IPC) mapping in the sharded step from anything torch.distributed adds"""
import multiprocessing as mp
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(rank, G, q_out, q_in, q_res):
    import numpy as np
    import ftrl_ffm_b200 as pkg
    nfl, nf, k, B = 39, int(os.environ.get('NF', 10_000_000)), 8, 65536
    m = pkg.FtrlModel("FFM", n_feats=nf, n_fields=nfl, n_factors=k, device=rank, rank=rank, world_size=G,
                      max_batch_rows=B, max_batch_nnz=B * nfl)
    q_out.put((rank, m.export_peer_blob()))
    blobs = q_in.get()
    m.attach_peers(blobs)
    m.randomize_state(seed=7)
    batches = [pkg.synth.criteo_batch(B, nfl, nf, seed=42 + 1000 * rank + i, dist=os.environ.get('DIST','zipf')) for i in range(3)]
    q_out.put((rank, "ready")); q_in.get()
    m.profile_enable(True)
    for i in range(2):
        m.train(**batches[i % 3])
    m.profile_reset()
    t0 = time.perf_counter()
    n = 6
    for i in range(n):
        m.train(**batches[i % 3])
    dt = time.perf_counter() - t0
    q_res.put((rank, dt / n, {k_: round(v["ms"] / n, 3) for k_, v in m.profile().items() if v["ms"] > 0}))


if __name__ == "__main__":
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    ctx = mp.get_context("spawn")
    q_out, q_res = ctx.Queue(), ctx.Queue()
    q_ins = [ctx.Queue() for _ in range(G)]
    ps = [ctx.Process(target=worker, args=(r, G, q_out, q_ins[r], q_res)) for r in range(G)]
    for p in ps:
        p.start()
    blobs = dict(q_out.get() for _ in range(G))
    for q in q_ins:
        q.put([blobs[r] for r in range(G)])
    for _ in range(G):
        q_out.get()
    for q in q_ins:
        q.put("go")
    for _ in range(G):
        print(q_res.get())
    for p in ps:
        p.join()
