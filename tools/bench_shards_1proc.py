"""single process, G devices: LogicalShards over real GPUs (plain peer pointers, no CUDA IPC) -- to separate the
cost of the IPC mapping from the cost of the exchange itself"""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ftrl_ffm_b200 as pkg

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nfl, nf, k, B = 39, int(os.environ.get('NF', 10_000_000)), 8, 65536
sh = pkg.LogicalShards(G, devices=list(range(G)), model_type="FFM", n_feats=nf, n_fields=nfl, n_factors=k,
                       max_batch_rows=B, max_batch_nnz=B * nfl)
for m in sh.models:
    m.randomize_state(seed=7)
batches = [[pkg.synth.criteo_batch(B, nfl, nf, seed=42 + 1000 * r + i, dist=os.environ.get('DIST','zipf')) for r in range(G)] for i in range(3)]
for m in sh.models:
    m.profile_enable(True)
for i in range(2):
    sh.train(batches[i % 3])
for m in sh.models:
    m.profile_reset()
t0 = time.perf_counter()
n = 6
for i in range(n):
    sh.train(batches[i % 3])
dt = time.perf_counter() - t0
print(f"G={G} single process: {G * B * n / dt:.3e} samples/s, {1e3 * dt / n:.2f} ms/step (includes host packing of pageable CSR)")
print({k_: round(v["ms"] / n, 3) for k_, v in sh.models[0].profile().items() if v["ms"] > 0})
