#!/usr/bin/env python
"""bench.py -- FTRL-FFM training throughput on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU path (oracle/_ref)

Workload (config.workload): BASELINE.json configs[3] -- synthetic Criteo-shaped libffm data, FFM, 39 fields,
10M features, k = 8, minibatch 64K samples per GPU; one "step" = forward + FTRL update of one minibatch.
  value  : samples/s with the CSR minibatches already resident in HBM (CUDA-event time, max over ranks)
  e2e    : samples/s through ftrl_train_batch() with PINNED HOST buffers: H2D copies of the CSR and the
           D2H read of the loss are inside the timed region (3 batches in flight)
  roofline: forward+update kernels (k_row_touch + k_row_materialise + k_ffm_tile +
           k_ffm_staged_rows + k_ffm_combine; LR/FM: k_lrfm_sample + k_lrfm_rows + k_lrfm_combine), algorithmic
           bytes of SURVEY.md 8(d) / DESIGN.md divided by their CUDA-event time, against MEASURED_PEAKS.json;
           `traffic` = DRAM bytes of the same kernels from the ncu capture in profiles/traffic.json, used only
           when that capture was made from the kernel sources the loaded library was built from
  cpu_baseline: the reference's train() (oracle/_ref, all host cores) on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, n_fields, n_feats, k, batch)
    "cfg4": ("FFM", 39, 10_000_000, 8, 65536),   # BASELINE.json configs[3]: the metric's configuration
    "cfg3": ("FFM", 39, 1_000_000, 4, 65536),    # configs[2]
    "cfg2-fm": ("FM", 39, 1_000_000, 16, 65536),  # configs[1]
    "cfg2-lr": ("LR", 39, 1_000_000, 1, 65536),
    # BASELINE.json configs[4]: 100M features, k 8: 374 GB of w/z/n, feature-sharded over 8 GPUs (46.8 GB each)
    "cfg5": ("FFM", 39, 100_000_000, 8, 65536),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--dist", default="zipf", choices=["zipf", "uniform"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--n-distinct", type=int, default=6, help="distinct synthetic minibatches cycled through")
    ap.add_argument("--cpu-samples", type=int, default=0, help="samples of the bounded CPU-baseline run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the extra uniform-id roofline measurement")
    return ap.parse_args()


def kernel_source_sha16():
    """identifies the kernel sources a profile was captured from (profiles/traffic.json carries the same stamp)"""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "ftrl-ffm_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h")):
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """lines before the mark (warm-up) are dropped when later ones exist"""
        self.n_mark = len(self.lines)

    def since_mark(self):
        return len(self.lines) - getattr(self, "n_mark", 0)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[getattr(self, "n_mark", 0):]:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_batches(wl, dist, batch, n_distinct, rank):
    import ftrl_ffm_b200 as pkg
    model, n_fields, n_feats, k, b0 = WORKLOADS[wl]
    batch = batch or b0
    return [pkg.synth.criteo_batch(batch, n_fields, n_feats, seed=42 + 1000 * rank + i, dist=dist)
            for i in range(n_distinct)], (model, n_fields, n_feats, k, batch)


CPU_BASELINE_SECONDS = 10.0


def workload_name(wl, model, n_fields, n_feats, k, B):
    """config.workload -- identical on the GPU arm and on the `--impl reference` arm"""
    tag = {"cfg4": " (BASELINE.json configs[3])", "cfg5": " (BASELINE.json configs[4])", "cfg3": " (BASELINE.json configs[2])",
           "cfg2-fm": " (BASELINE.json configs[1])", "cfg2-lr": " (BASELINE.json configs[1])"}.get(wl, "")
    return f"{wl}: {model} n_fields={n_fields} n_feats={n_feats} k={k} batch={B}/GPU{tag}"


def alg_bytes(model, F, k, U, nnz, B):
    """20 B per touched coordinate (read z,n; write z,n,w) + CSR bytes"""
    lat = U * (F - 1) * k if model == "FFM" else U * k if model == "FM" else 0
    return 20.0 * (lat + U) + 12.0 * nnz + 8.0 * B


def cpu_baseline(wl, dist, n_samples, threads=None):
    """the reference's own train() on this box's host cores, bounded sample of the same workload"""
    from oracle.cpu_model import CpuModel, have_ref
    import ftrl_ffm_b200 as pkg
    model, n_fields, n_feats, k, _ = WORKLOADS[wl]
    cores = threads or os.cpu_count() or 1
    fold = min(n_feats, 1_000_000)  # the reference's per-weight mutex + vector rows make 10M rows impractical
    data = pkg.synth.criteo_batch(n_samples, n_fields, n_feats, seed=42, dist=dist)
    if fold != n_feats:
        per_big, per_small = n_feats // n_fields, fold // n_fields
        local = data["feat"].astype(np.int64) - data["field"].astype(np.int64) * per_big
        data["feat"] = (data["field"].astype(np.int64) * per_small + local % per_small).astype(np.int32)
    kind = "reference" if have_ref() else "port"
    if kind == "reference":
        m = CpuModel("ref", model, fold, n_fields, k, fast_init=True)
        m.stage_csr(**data)
        m.train_staged(n_threads=cores)  # warm-up pass (materialises w, touches pages)
        secs, passes = 0.0, 0
        while secs < CPU_BASELINE_SECONDS and passes < 1000:  # about 10 s of CPU work (training passes over the sample)
            dt, _ = m.train_staged(n_threads=cores)
            secs += dt
            passes += 1
    else:
        m = CpuModel("oracle", model, fold, n_fields, k)
        m.train_csr(**data)
        secs, passes = 0.0, 0
        while secs < CPU_BASELINE_SECONDS and passes < 1000:
            t0 = time.perf_counter()
            m.train_csr(**data)
            secs += time.perf_counter() - t0
            passes += 1
        cores = 1
    one = None
    if kind == "reference" and cores > 1:
        n1 = max(256, n_samples // 8)
        sub = pkg.synth.slice_csr(data, 0, n1)
        m1 = CpuModel("ref", model, fold, n_fields, k, fast_init=True)
        m1.stage_csr(**sub)
        m1.train_staged(n_threads=1)
        s1, p1 = 0.0, 0
        while s1 < 2.0 and p1 < 1000:
            dt, _ = m1.train_staged(n_threads=1)
            s1 += dt
            p1 += 1
        one = p1 * n1 / s1
        m1.close()
    sample = (f"{passes} training passes over {n_samples} samples of the step-0 distribution, {model} F={n_fields} "
              f"k={k}, ids folded into {fold} rows, after one warm-up pass; reference train() from {cores} threads "
              "with the chunking of ftrl_offline.cpp:63-103 (constructor replaced by a zero fill, see "
              "oracle/ref_shim.cpp)")
    return {"value": passes * n_samples / secs, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample,
            "seconds": secs, "value_1_thread": one}


def run_reference(args, rank, world):
    if rank != 0:
        return
    model, n_fields, n_feats, k, b0 = WORKLOADS[args.workload]
    n = args.cpu_samples or (4096 if model == "FFM" else 65536)
    from oracle.cpu_model import CpuModel, have_ref
    import ftrl_ffm_b200 as pkg
    cores = os.cpu_count() or 1
    fold = min(n_feats, 1_000_000)
    kind = "reference" if have_ref() else "port"
    datas = []
    for i in range(min(args.n_distinct, args.steps + args.warmup)):
        d = pkg.synth.criteo_batch(n, n_fields, n_feats, seed=42 + i, dist=args.dist)
        if fold != n_feats:
            per_big, per_small = n_feats // n_fields, fold // n_fields
            local = d["feat"].astype(np.int64) - d["field"].astype(np.int64) * per_big
            d["feat"] = (d["field"].astype(np.int64) * per_small + local % per_small).astype(np.int32)
        datas.append(d)
    m = CpuModel("ref" if kind == "reference" else "oracle", model, fold, n_fields, k)
    total = 0.0
    for s in range(args.warmup + args.steps):
        d = datas[s % len(datas)]
        if kind == "reference":
            m.stage_csr(**d)
            secs, _ = m.train_staged(n_threads=cores)
        else:
            t0 = time.perf_counter()
            m.train_csr(**d)
            secs = time.perf_counter() - t0
        if s >= args.warmup:
            total += secs
    value = n * args.steps / total
    used = cores if kind == "reference" else 1
    sample = (f"each step = {n} samples (bounded sample of the 64K-sample minibatch), ids folded into {fold} rows, "
              f"reference train() from {used} host threads")
    line = {
        "impl": "reference", "metric": "FTRL-FFM train samples/sec", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, model, n_fields, n_feats, k, b0),
                   "ids": f"{args.dist}" + (" s=1.2" if args.dist == "zipf" else ""),
                   "samples_per_step": n, "note": "CPU arm: each step is a bounded sample of the minibatch"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import ftrl_ffm_b200 as pkg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this trainer has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batches, (model, n_fields, n_feats, k, B) = make_batches(args.workload, args.dist, args.batch, args.n_distinct, rank)
    nnz = int(batches[0]["row_ptr"][-1])
    # the synthetic minibatches are resident in HBM and never rewritten: the library may read the ids of batch i+1
    # while batch i trains (ftrl_config.reserved[0], include/ftrl_b200.h)
    m = pkg.FtrlModel(model, n_feats=n_feats, n_fields=n_fields, n_factors=k, device=local_rank,
                      max_batch_rows=B, max_batch_nnz=nnz, rank=rank, world_size=world, stable_device_inputs=True)
    if world > 1:
        # feature-sharded tables: every rank maps every peer's shard (CUDA IPC); the blobs travel by all-gather
        blob = torch.frombuffer(bytearray(m.export_peer_blob()), dtype=torch.uint8).cuda()
        blobs = [torch.zeros_like(blob) for _ in range(world)]
        dist.all_gather(blobs, blob)
        m.attach_peers([bytes(t.cpu().numpy().tobytes()) for t in blobs])
        dist.barrier()
    m.randomize_state(seed=7)  # live latent state (cold-start latents stay exactly 0 in the reference)
    stream = torch.cuda.current_stream()
    m.set_stream(stream.cuda_stream)

    # ---- inputs resident in HBM ----
    dev = []
    for b in batches:
        dev.append({key: torch.from_numpy(np.ascontiguousarray(v)).cuda() for key, v in b.items()})
    d_loss = torch.zeros(1, dtype=torch.float64, device="cuda")

    def step_dev(i):
        d = dev[i % len(dev)]
        m.train_device(B, nnz, d["row_ptr"].data_ptr(), d["field"].data_ptr(), d["feat"].data_ptr(),
                       d["val"].data_ptr(), d["label"].data_ptr(), 0, d_loss.data_ptr())

    # per-batch U (distinct rows) for the algorithmic byte count, outside the timed region
    U = []
    for i in range(len(dev)):
        step_dev(i)
        st = m.last_batch_stats()
        u = st["n_unique"]
        if world > 1:  # owner-side counts: sum over ranks = distinct rows of the global minibatch
            t = torch.tensor([u], dtype=torch.int64, device="cuda")
            dist.all_reduce(t)
            u = int(t.item())
        U.append(u)
    launches_per_step = st["kernel_launches"]
    fused_rows = st["n_fused_rows"]
    # the sampler starts before the warm-up (same load as the timed steps): nvidia-smi needs ~100 ms to deliver
    # its first line and K steps can be shorter than that
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        step_dev(i)
    barrier()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step_dev(i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    missing = 0 if sampler.since_mark() else 1
    if world > 1:
        t = torch.tensor([ms, float(missing)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, missing = float(t[0].item()), int(t[1].item())
    if missing:
        # no nvidia-smi line fell inside the timed region on some rank: keep the same load running (untimed,
        # the same number of steps on every rank -- the sharded step is collective) for about 0.4 s
        for j in range(min(400, max(8, int(400.0 / max(ms / args.steps, 0.05))))):
            step_dev(j)
        barrier()
    clocks = sampler.stop()
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)

    # ---- per-kernel time of the same steps (CUDA events on the launching stream, inside the library) ----
    m.profile_enable(True)
    m.profile_reset()
    for i in range(args.steps):
        step_dev(i)
    torch.cuda.synchronize()
    prof = m.profile()
    m.profile_enable(False)
    # forward + FTRL update kernels (sharded runs: + the row pull, the owner-side materialise and apply)
    hot = ["materialise", "sample", "rows", "combine", "generic", "exchange", "pull", "apply"]
    hot_ms = sum(prof[p]["ms"] for p in hot if p in prof) / args.steps
    Ubar = float(np.mean([U[i % len(U)] for i in range(args.steps)]))
    bytes_alg = alg_bytes(model, n_fields, k, Ubar, nnz * world, B * world) / world  # per GPU
    peak, peak_src = peaks()
    achieved = bytes_alg / (hot_ms / 1e3) / 1e9
    traffic, traffic_note = None, "no ncu capture for this workload"
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            tj = json.load(open(tp))
            if tj.get("_kernel_source_sha16") != kernel_source_sha16():
                traffic_note = "profiles/traffic.json was captured from other kernel sources: not reported"
            else:
                traffic = tj.get(f"{args.workload}-{args.dist}")
                traffic_note = tj.get("_source") if traffic else traffic_note
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
                "dram_gbs_from_traffic": (traffic / (hot_ms / 1e3) / 1e9) if traffic else None,
                "kernels": ("k_row_touch + k_row_materialise + k_ffm_tile + k_ffm_staged_rows + k_ffm_combine "
                            "(forward + FTRL update)" if world == 1 else
                            "owner select/sort/k_owner_materialise (pushes w to the row caches) + k_ffm_tile + k_ffm_staged_rows + "
                            "k_ffm_combine + k_owner_apply (forward + FTRL update + row exchange)") if model == "FFM"
                else "k_lrfm_sample + k_lrfm_rows + k_lrfm_combine",
                "alg_bytes_per_step": bytes_alg, "kernel_ms_per_step": hot_ms,
                "phase_ms_per_step": {p: v["ms"] / args.steps for p, v in prof.items() if v["ms"] > 0},
                "U_over_nnz": Ubar / nnz}

    # ---- the cache-hostile case SURVEY.md 8(d) asks to report next to it: uniform ids (U ~ nnz) ----
    roofline_uniform = None
    if args.dist == "zipf" and not args.no_secondary and world == 1:
        ub = [pkg.synth.criteo_batch(B, n_fields, n_feats, seed=4242 + i, dist="uniform") for i in range(3)]
        udev = [{key: torch.from_numpy(np.ascontiguousarray(v)).cuda() for key, v in b_.items()} for b_ in ub]

        def step_u(i):
            d = udev[i % len(udev)]
            m.train_device(B, nnz, d["row_ptr"].data_ptr(), d["field"].data_ptr(), d["feat"].data_ptr(),
                           d["val"].data_ptr(), d["label"].data_ptr(), 0, d_loss.data_ptr())
        Uu = []
        for i in range(len(udev)):
            step_u(i)
            Uu.append(m.last_batch_stats()["n_unique"])
        nsteps_u = 6
        m.profile_enable(True)
        m.profile_reset()
        torch.cuda.synchronize()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record(stream)
        for i in range(nsteps_u):
            step_u(i)
        u1.record(stream)
        torch.cuda.synchronize()
        uprof = m.profile()
        m.profile_enable(False)
        u_hot = sum(uprof[p]["ms"] for p in hot if p in uprof) / nsteps_u
        u_bytes = alg_bytes(model, n_fields, k, float(np.mean([Uu[i % len(Uu)] for i in range(nsteps_u)])), nnz, B)
        u_ach = u_bytes / (u_hot / 1e3) / 1e9
        roofline_uniform = {"achieved": u_ach, "peak": peak, "unit": "GB/s", "frac": u_ach / peak,
                            "value_samples_per_s": B * nsteps_u / (u0.elapsed_time(u1) / 1e3),
                            "kernel_ms_per_step": u_hot, "alg_bytes_per_step": u_bytes,
                            "U_over_nnz": float(np.mean(Uu)) / nnz}
        del udev

    # ---- end to end through the host-pointer C ABI: pinned host CSR, H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        pinned = []
        for b in batches:
            pb = {}
            for key, v in b.items():
                t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
                pb[key] = t
            pinned.append(pb)
        losses = torch.zeros(args.steps + args.warmup, dtype=torch.float64).pin_memory()
        lib = m.lib

        def step_host(i, slot):
            p = pinned[i % len(pinned)]
            rc = lib.ftrl_train_batch(m.h, B, p["row_ptr"].data_ptr(), p["field"].data_ptr(), p["feat"].data_ptr(),
                                      p["val"].data_ptr(), p["label"].data_ptr(), None,
                                      losses.data_ptr() + 8 * slot)
            if rc != 0:
                raise RuntimeError(lib.ftrl_last_error(m.h))

        for i in range(args.warmup):
            step_host(i, i)
        m.sync()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_host(i, args.warmup + i)
        m.sync()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = 8 * (B + 1) + 12 * nnz + 4 * B
        e2e = {"value": world * B * args.steps / dt, "unit": "samples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 8, "ms_per_step": 1e3 * dt / args.steps,
               "mean_loss_last_step": float(losses[-1].item()) / B}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            n_cpu = args.cpu_samples or (16384 if model == "FFM" else 262144)
            cpu = cpu_baseline(args.workload, args.dist, n_cpu)
        except Exception as ex:  # the bench line must still be printed
            cpu = {"value": None, "unit": "samples/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}

    if rank == 0:
        line = {
            "metric": "FTRL-FFM train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, model, n_fields, n_feats, k, B),
                       "ids": f"{args.dist}" + (" s=1.2" if args.dist == "zipf" else ""),
                       "state": "randomized live z/n (ftrl_randomize_state)", "mode": "minibatch",
                       "l2": "inputs larger than L2 (rows touched per step >> 126 MB), no explicit flush",
                       "parallelism": "single GPU" if world == 1 else
                       f"{world} GPUs: tables sharded by feature id (feat mod {world}), samples split across ranks, "
                       f"duplicates reduced where the samples live, one w plane in / one gradient sum out per distinct (row, rank) over NVLink peer memory, global batch {B * world}",
                       "distinct_batches": len(batches), "fused_rows_per_step": fused_rows},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step) * args.steps,
            "roofline": roofline, "roofline_uniform_ids": roofline_uniform, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
