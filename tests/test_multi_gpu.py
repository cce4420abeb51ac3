"""Real multi-GPU runs of the feature-sharded path (self-skipping below 2 devices): one process per GPU under
torchrun with peers mapped over CUDA IPC (tools/mgpu_check.py), and the C++ `main --n_gpus N` (several handles in
one process).  The single-GPU box of the driver's `-m gpu` run skips these; `gpurun --gpus 2|4|8` runs them."""
import os
import subprocess
import sys

import numpy as np
import pytest

import ftrl_ffm_b200 as pkg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_devices():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_torchrun_sharded_equals_single_gpu(world):
    if n_devices() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
                          os.path.join(ROOT, "tools", "mgpu_check.py")], capture_output=True, text=True, env=env, timeout=900)
    print(out.stdout[-3000:], out.stderr[-2000:])
    assert out.returncode == 0 and "MGPU_CHECK OK" in out.stdout


@pytest.mark.parametrize("world", [2, 4])
def test_main_n_gpus_trains_saves_and_matches_single_gpu(world, tmp_path):
    """`main --n_gpus N`: same data, same seed: the epoch lines and the saved model of an N-GPU run equal the
    single-GPU run's (up to fp32 re-association of the per-rank partial sums)"""
    if n_devices() < world:
        pytest.skip(f"needs {world} GPUs")
    main = os.path.join(ROOT, "ftrl-ffm_b200", "main")
    nfl, nf = 10, 3000
    data = pkg.synth.criteo_batch(6000, nfl, nf, seed=3, dist="zipf", planted=True)
    path = str(tmp_path / "train.ffm")
    pkg.synth.write_text(data, path)
    outs, models = [], []
    for g in (1, world):
        mp = str(tmp_path / f"model_{g}.zst")
        r = subprocess.run([main, "--train_data", path, "--eval_data", path, "--model_type", "FFM", "--n_fields", str(nfl),
                            "--n_feats", str(nf), "--n_factors", "4", "--n_epochs", "2", "--online", "false",
                            "--batch_size", "512", "--seed", "9", "--n_gpus", str(g), "--model_path", mp, "--auc", "true"],
                           capture_output=True, text=True, timeout=600)
        print(r.stdout, r.stderr)
        assert r.returncode == 0
        outs.append([ln for ln in r.stdout.splitlines() if ln.startswith("epoch")])
        m = pkg.FtrlModel("FFM", n_feats=nf, n_fields=nfl, n_factors=4)
        m.load_compressed_model(mp)
        models.append((m.lin_w, m.vec_w))
        m.close()

    def nums(lines):
        import re
        return np.array([[float(x) for x in re.findall(r"(?:loss|auc): ([0-9.]+)", ln)] for ln in lines if "loss" in ln or "auc" in ln], dtype=object)
    a, b = nums(outs[0]), nums(outs[1])
    assert len(a) == len(b) and len(a) >= 4
    for x, y in zip(a, b):
        assert np.allclose(np.array(x, float), np.array(y, float), atol=2e-4)
    assert np.allclose(models[0][0], models[1][0], rtol=1e-4, atol=1e-6)
    assert np.allclose(models[0][1], models[1][1], rtol=1e-4, atol=1e-6)
