#!/usr/bin/env python
"""Text front end, end to end: Criteo-shaped libffm TEXT -> parser threads -> pinned CSR -> GPU, through
the C++ `main` (run ON the GPU box).  Reports samples/s of `main --online true` (streams + parses the text every epoch;
parsing of block i+1 overlaps the training of block i) at 1 thread and at all cores, and of `--csr_cache true` (binary
image from the second run on).  usage: tools/bench_frontend.py [n_lines] [out.json]"""
import json, os, re, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import ftrl_ffm_b200 as pkg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "frontend.json")
path = "/tmp/frontend.ffm"
b = pkg.synth.criteo_batch(n, 39, 10_000_000, seed=42)
# vectorised writer (synth.write_text formats token by token)
cols = [np.char.add(np.char.add(np.char.add(str(f) + ":", b["feat"][f::39].astype(str)), ":"),
                    np.char.mod("%.6g", b["val"][f::39])) for f in range(39)]
lines = b["label"].astype(str)
for c in cols:
    lines = np.char.add(np.char.add(lines, " "), c)
open(path, "w").write("\n".join(lines.tolist()) + "\n")
size = os.path.getsize(path)
main = os.path.join(ROOT, "ftrl-ffm_b200", "main")
res = {"lines": n, "bytes_per_line": size / n, "cores": os.cpu_count(), "runs": []}


def run(threads, extra, label):
    cmd = [main, "--train_data", path, "--model_type", "FFM", "--n_fields", "39", "--n_feats", "10000000", "--n_factors", "8",
           "--n_epochs", "3", "--n_threads", str(threads), "--batch_size", "65536", "--seed", "1"] + extra
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    wall = time.time() - t0
    ep = [float(x) for x in re.findall(r"train time: ([0-9.]+)s", r.stdout)]
    rec = {"what": label, "threads": threads, "rc": r.returncode, "epoch_s": ep, "wall_s": wall,
           "samples_per_s_last_epoch": n / ep[-1] if ep else None, "text_GBps_last_epoch": size / ep[-1] / 1e9 if ep else None}
    res["runs"].append(rec)
    print(json.dumps(rec), flush=True)


run(1, ["--online", "true"], "stream + parse every epoch, 1 parser thread")
run(os.cpu_count(), ["--online", "true"], "stream + parse every epoch, all cores")
run(os.cpu_count(), ["--online", "true", "--csr_cache", "true"], "binary CSR image (written by this run)")
run(os.cpu_count(), ["--online", "true", "--csr_cache", "true"], "binary CSR image (reused)")
json.dump(res, open(out, "w"), indent=1)
