"""ftrl-ffm_b200: B200-native FTRL trainer for LR / FM / FFM -- the hot path of massquantity/Ftrl-FFM
(per-sample forward + FTRL z/n/w update) as hand-written sm_100a CUDA behind a C ABI.

  csrc/     CUDA kernels + the C ABI (include/ftrl_b200.h) -> libftrl_b200.so
  host/     C++17 host side: `main` CLI drop-in, libsvm/libffm parser -> pinned CSR, epoch drivers
  binding   ctypes mirror of the reference's model interface (used by tests/ and bench.py)
  synth     synthetic Criteo-shaped data

Import through the alias module at the repo root: `import ftrl_ffm_b200`.
"""
from .binding import (ABI_SYMBOLS, LIB_PATH, MODE_BATCH, MODE_SEQUENTIAL, BatchStats, Config, FtrlError,  # noqa: F401
                      FtrlModel, LogicalShards, load_library, merge_states, shard_state)
from .build import build_host, build_library  # noqa: F401
from . import synth  # noqa: F401
