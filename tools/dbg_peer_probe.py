"""two processes, one GPU each: scattered row reads / writes against the PEER shard through the library's own
mappings (CUDA IPC), both ranks at the same time -- bisects the mapping from the tile kernel's access pattern"""
import ctypes as C
import multiprocessing as mp
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(rank, G, q_out, q_in, q_res):
    import ftrl_ffm_b200 as pkg
    nfl, nf, k, B = 39, int(os.environ.get("NF", 10_000_000)), 8, 65536
    m = pkg.FtrlModel("FFM", n_feats=nf, n_fields=nfl, n_factors=k, device=rank, rank=rank, world_size=G,
                      max_batch_rows=B, max_batch_nnz=B * nfl)
    q_out.put((rank, m.export_peer_blob()))
    m.attach_peers(q_in.get())
    fn = m.lib.ftrl_dbg_peer_traffic
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
    n_rows = B * nfl // 2
    res = {}
    for name, q, flags, grid in [("local rd", rank, 1, 1184), ("peer rd", 1 - rank, 1, 1184), ("peer wr", 1 - rank, 2, 1184),
                                 ("peer rd+wr", 1 - rank, 3, 1184), ("peer rd staging", 1 - rank, 4, 1184),
                                 ("peer rd tab sequential", 1 - rank, 16, 1184), ("peer wr tab", 1 - rank, 8, 1184)]:
        ms = C.c_float()
        fn(m.h, q, flags, n_rows, grid, 1, C.byref(ms))  # warm
        q_out.put((rank, "ready")); q_in.get()
        rc = fn(m.h, q, flags, n_rows, grid, 5, C.byref(ms))
        byt = n_rows * 1248 * (1 if flags != 3 else 2)
        res[name] = f"{ms.value:.2f} ms {byt / ms.value / 1e6:.0f} GB/s"
    q_res.put((rank, res))


if __name__ == "__main__":
    G = 2
    ctx = mp.get_context("spawn")
    q_out, q_res = ctx.Queue(), ctx.Queue()
    q_ins = [ctx.Queue() for _ in range(G)]
    ps = [ctx.Process(target=worker, args=(r, G, q_out, q_ins[r], q_res)) for r in range(G)]
    for p in ps:
        p.start()
    blobs = dict(q_out.get() for _ in range(G))
    for q in q_ins:
        q.put([blobs[r] for r in range(G)])
    for _ in range(7):
        for _ in range(G):
            q_out.get()
        for q in q_ins:
            q.put("go")
    for _ in range(G):
        print(q_res.get())
    for p in ps:
        p.join()
