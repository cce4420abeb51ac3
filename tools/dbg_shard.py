import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, "/root/repo")
import numpy as np
import ftrl_ffm_b200 as pkg
nf, nfl, k, B = 3000, 13, 8, 512
kw = dict(model_type="FFM", n_feats=nf, n_fields=nfl, n_factors=k)
sh = pkg.LogicalShards(2, max_batch_rows=B, max_batch_nnz=B * nfl, **kw)
parts = [pkg.synth.criteo_batch(B, nfl, nf, seed=r) for r in range(2)]
t0 = time.time()
p0 = sh.models[0].train(**parts[0], sync=False); print("rank0 enqueued", time.time() - t0, flush=True)
p1 = sh.models[1].train(**parts[1], sync=False); print("rank1 enqueued", time.time() - t0, flush=True)
for r, m in enumerate(sh.models):
    try:
        m.sync(); print("rank", r, "synced", time.time() - t0, flush=True)
    except Exception as e:
        print("rank", r, "ERR", e, time.time() - t0, flush=True)
