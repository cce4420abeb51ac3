"""ctypes binding of the C ABI in include/ftrl_b200.h, plus `FtrlModel`: the Python-side mirror of
the reference's model interface (ftrl::FtrlModel / LR / FM / FFM, src/include/model/*.h) operating on
CSR minibatches instead of one feat_vec per call.

There is no CPU fallback: if libftrl_b200.so is missing, or no CUDA device is usable, loading or
creating a model raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# FTRL_B200_LIB: load another build of the same library (A/B timing of two builds in one GPU session)
LIB_PATH = os.environ.get("FTRL_B200_LIB") or os.path.join(HERE, "libftrl_b200.so")

MODEL_TYPES = {"LR": 0, "FM": 1, "FFM": 2}
MODE_BATCH, MODE_SEQUENTIAL = 0, 1
PEER_BLOB_BYTES = 1024


class FtrlError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"ftrl_b200 status {status}: {msg}")
        self.status = status


class Config(C.Structure):
    """struct ftrl_config"""
    _fields_ = [
        ("model_type", C.c_int32), ("n_feats", C.c_int32), ("n_fields", C.c_int32), ("n_factors", C.c_int32),
        ("init_mean", C.c_float), ("init_stddev", C.c_float), ("w_alpha", C.c_float), ("w_beta", C.c_float),
        ("w_l1", C.c_float), ("w_l2", C.c_float), ("mode", C.c_int32), ("device", C.c_int32),
        ("seed", C.c_uint64), ("max_batch_rows", C.c_int64), ("max_batch_nnz", C.c_int64),
        ("rank", C.c_int32), ("world_size", C.c_int32), ("reserved", C.c_int32 * 8),
    ]


class BatchStats(C.Structure):
    """struct ftrl_batch_stats"""
    _fields_ = [(n, C.c_int64) for n in
                ("n_rows", "nnz_valid", "n_unique", "n_fused_rows", "n_segmented_rows", "n_chunks",
                 "kernel_launches")] + [("reserved", C.c_int64 * 5)]


# every symbol include/ftrl_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "ftrl_config_default", "ftrl_abi_version", "ftrl_create", "ftrl_destroy", "ftrl_last_error",
    "ftrl_train_batch", "ftrl_train_batch_device", "ftrl_predict_batch", "ftrl_predict_batch_device",
    "ftrl_sync", "ftrl_get_weights", "ftrl_set_weights", "ftrl_get_state", "ftrl_set_state",
    "ftrl_get_rows", "ftrl_set_rows", "ftrl_row_len", "ftrl_has_zero_weights", "ftrl_save_model",
    "ftrl_load_model", "ftrl_save_model_text", "ftrl_load_model_text", "ftrl_set_stream",
    "ftrl_profile_enable", "ftrl_profile_reset", "ftrl_profile_read", "ftrl_last_batch_stats",
    "ftrl_randomize_state", "ftrl_alloc_pinned", "ftrl_free_pinned",
    "ftrl_export_peer_blob", "ftrl_attach_peers", "ftrl_eval_auc", "ftrl_eval_auc_device",
]

_lib = None


def load_library(path: str | None = None):
    """dlopen libftrl_b200.so and declare prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(p)
    vp, i64, i32p, f32p, i64p, f64p = C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    lib.ftrl_config_default.argtypes = [C.POINTER(Config)]
    lib.ftrl_config_default.restype = None
    lib.ftrl_abi_version.restype = C.c_int
    lib.ftrl_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.ftrl_destroy.argtypes = [vp]
    lib.ftrl_destroy.restype = None
    lib.ftrl_last_error.argtypes = [vp]
    lib.ftrl_last_error.restype = C.c_char_p
    lib.ftrl_train_batch.argtypes = [vp, i64, i64p, i32p, i32p, f32p, i32p, f32p, f64p]
    lib.ftrl_train_batch_device.argtypes = [vp, i64, i64, i64p, i32p, i32p, f32p, i32p, f32p, f64p]
    lib.ftrl_predict_batch.argtypes = [vp, i64, i64p, i32p, i32p, f32p, i32p, C.c_int, f32p, f64p]
    lib.ftrl_predict_batch_device.argtypes = [vp, i64, i64, i64p, i32p, i32p, f32p, i32p, C.c_int, f32p, f64p]
    lib.ftrl_sync.argtypes = [vp]
    for n in ("ftrl_get_weights", "ftrl_set_weights"):
        getattr(lib, n).argtypes = [vp, f32p, f32p, f32p]
    for n in ("ftrl_get_state", "ftrl_set_state"):
        getattr(lib, n).argtypes = [vp, C.c_int, f32p, f32p, f32p]
    for n in ("ftrl_get_rows", "ftrl_set_rows"):
        getattr(lib, n).argtypes = [vp, C.c_int, i64, i64, f32p, f32p]
    lib.ftrl_row_len.argtypes = [vp]
    lib.ftrl_row_len.restype = i64
    lib.ftrl_has_zero_weights.argtypes = [vp, C.POINTER(C.c_int)]
    lib.ftrl_save_model.argtypes = [vp, C.c_char_p, C.c_int]
    for n in ("ftrl_load_model", "ftrl_save_model_text", "ftrl_load_model_text"):
        getattr(lib, n).argtypes = [vp, C.c_char_p]
    lib.ftrl_set_stream.argtypes = [vp, vp]
    lib.ftrl_profile_enable.argtypes = [vp, C.c_int]
    lib.ftrl_profile_reset.argtypes = [vp]
    lib.ftrl_profile_read.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.ftrl_last_batch_stats.argtypes = [vp, C.POINTER(BatchStats)]
    lib.ftrl_randomize_state.argtypes = [vp, C.c_uint64, C.c_float, C.c_float, C.c_float]
    lib.ftrl_alloc_pinned.argtypes = [C.c_size_t]
    lib.ftrl_alloc_pinned.restype = vp
    lib.ftrl_free_pinned.argtypes = [vp]
    lib.ftrl_free_pinned.restype = None
    for n in ("ftrl_eval_auc", "ftrl_eval_auc_device"):
        getattr(lib, n).argtypes = [vp, i64, f32p, i32p, C.POINTER(C.c_double)]
    lib.ftrl_export_peer_blob.argtypes = [vp, vp]
    lib.ftrl_attach_peers.argtypes = [vp, vp]
    if path is None:
        _lib = lib
    return lib


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _csr(row_ptr, field, feat, val, label):
    rp = np.ascontiguousarray(row_ptr, np.int64)
    fe = np.ascontiguousarray(feat, np.int32)
    fi = np.zeros(len(fe), np.int32) if field is None else np.ascontiguousarray(field, np.int32)
    va = np.ascontiguousarray(val, np.float32)
    la = None if label is None else np.ascontiguousarray(label, np.int32)
    return rp, fi, fe, va, la


class FtrlModel:
    """LR / FM / FFM with FTRL on one B200, behind the C ABI.

    Mirrors the reference's model objects: `train` / `predict` (src/include/model/ftrl_model.h:18-19)
    take a CSR minibatch; `bias`, `lin_w`, `vec_w` are readable/writable like the reference's public
    members; `get_state`/`set_state` expose n and z (protected in the reference).
    """

    def __init__(self, model_type="FFM", n_feats=10000, n_fields=8, n_factors=16, init_mean=0.0,
                 init_stddev=0.02, w_alpha=1e-4, w_beta=1.0, w_l1=0.1, w_l2=5.0, mode="batch",
                 device=0, seed=42, max_batch_rows=0, max_batch_nnz=0, rank=0, world_size=1, stable_device_inputs=False):
        self.lib = load_library()
        mt = str(model_type).upper()  # cmd_option.cpp:70 upper-cases --model_type
        if mt not in MODEL_TYPES:
            raise ValueError(f"Invalid model_type: {model_type}, expect `LR`, `FM` or `FFM`.")
        cfg = Config()
        self.lib.ftrl_config_default(C.byref(cfg))
        cfg.model_type = MODEL_TYPES[mt]
        cfg.n_feats, cfg.n_fields, cfg.n_factors = int(n_feats), int(n_fields), int(n_factors)
        cfg.init_mean, cfg.init_stddev = init_mean, init_stddev
        cfg.w_alpha, cfg.w_beta, cfg.w_l1, cfg.w_l2 = w_alpha, w_beta, w_l1, w_l2
        cfg.mode = MODE_SEQUENTIAL if str(mode).lower().startswith("seq") else MODE_BATCH
        cfg.device, cfg.seed = int(device), int(seed)
        cfg.max_batch_rows, cfg.max_batch_nnz = int(max_batch_rows), int(max_batch_nnz)
        cfg.rank, cfg.world_size = int(rank), int(world_size)
        # ftrl_config.reserved[0] bit 0: device-resident CSR arrays are complete before the previous train call was made
        cfg.reserved[0] = 1 if stable_device_inputs else 0
        self.cfg = cfg
        self.model_type = mt
        self.n_feats = int(n_feats)
        h = C.c_void_p()
        rc = self.lib.ftrl_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise FtrlError(rc, (self.lib.ftrl_last_error(None) or b"").decode())
        self.h = h
        self.row_len = int(self.lib.ftrl_row_len(self.h))
        self._keep = []  # host buffers of in-flight async calls

    # -- plumbing ---------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise FtrlError(rc, (self.lib.ftrl_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.ftrl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.ftrl_sync(self.h))
        self._keep.clear()

    # -- hot path, host CSR -------------------------------------------------------
    def train(self, row_ptr, field, feat, val, label, want_logits=True, sync=True):
        """One minibatch through ftrl_train_batch.  Returns (logits | None, loss_sum)."""
        rp, fi, fe, va, la = _csr(row_ptr, field, feat, val, label)
        n = len(rp) - 1
        logits = np.zeros(n, np.float32) if want_logits else None
        loss = np.zeros(1, np.float64)
        self._check(self.lib.ftrl_train_batch(self.h, n, _np_ptr(rp), _np_ptr(fi), _np_ptr(fe), _np_ptr(va),
                                              _np_ptr(la), _np_ptr(logits), _np_ptr(loss)))
        self._keep.append((rp, fi, fe, va, la, logits, loss))
        if sync:
            self.sync()
        return logits, (float(loss[0]) if sync else loss)

    def predict(self, row_ptr, field, feat, val, label=None, output_prob=False):
        """ftrl_predict_batch.  Returns (out, loss_sum | None)."""
        rp, fi, fe, va, la = _csr(row_ptr, field, feat, val, label)
        n = len(rp) - 1
        out = np.zeros(n, np.float32)
        loss = np.zeros(1, np.float64) if la is not None else None
        self._check(self.lib.ftrl_predict_batch(self.h, n, _np_ptr(rp), _np_ptr(fi), _np_ptr(fe), _np_ptr(va),
                                                _np_ptr(la), int(output_prob), _np_ptr(out), _np_ptr(loss)))
        self._keep.append((rp, fi, fe, va, la, out, loss))
        self.sync()
        return out, (float(loss[0]) if loss is not None else None)

    # -- hot path, device-resident CSR (raw device pointers as ints) -----------------
    def train_device(self, n_rows, nnz, row_ptr, field, feat, val, label, logits=0, loss=0):
        self._check(self.lib.ftrl_train_batch_device(self.h, int(n_rows), int(nnz), row_ptr, field, feat, val,
                                                     label, logits or None, loss or None))

    def predict_device(self, n_rows, nnz, row_ptr, field, feat, val, label, output_prob, out, loss=0):
        self._check(self.lib.ftrl_predict_batch_device(self.h, int(n_rows), int(nnz), row_ptr, field, feat, val,
                                                       label or None, int(output_prob), out, loss or None))

    # -- state ------------------------------------------------------------------------
    def _get(self, which):
        b = np.zeros(1, np.float32)
        lin = np.zeros(self.n_local, np.float32)
        vec = np.zeros((self.n_local, self.row_len), np.float32) if self.row_len else None
        if which == 0:
            self._check(self.lib.ftrl_get_weights(self.h, _np_ptr(b), _np_ptr(lin), _np_ptr(vec)))
        else:
            self._check(self.lib.ftrl_get_state(self.h, which, _np_ptr(b), _np_ptr(lin), _np_ptr(vec)))
        return b, lin, vec

    def get_state(self) -> dict:
        """{bias: [w, n, z], lin_w, lin_n, lin_z, vec_w, vec_n, vec_z} (planes in the reference layout)"""
        st = {}
        bias = np.zeros(3, np.float32)
        for which, nm in ((0, "w"), (1, "n"), (2, "z")):
            b, lin, vec = self._get(which)
            bias[which] = b[0]
            st["lin_" + nm] = lin
            if vec is not None:
                st["vec_" + nm] = vec
        st["bias"] = bias
        return st

    def set_state(self, st: dict) -> None:
        for which, nm in ((0, "w"), (1, "n"), (2, "z")):
            b = np.array([st["bias"][which]], np.float32) if "bias" in st else None
            lin = np.ascontiguousarray(st["lin_" + nm], np.float32) if "lin_" + nm in st else None
            vec = None
            if self.row_len and "vec_" + nm in st:
                vec = np.ascontiguousarray(st["vec_" + nm], np.float32).reshape(self.n_local, self.row_len)
            if which == 0:
                self._check(self.lib.ftrl_set_weights(self.h, _np_ptr(b), _np_ptr(lin), _np_ptr(vec)))
            else:
                self._check(self.lib.ftrl_set_state(self.h, which, _np_ptr(b), _np_ptr(lin), _np_ptr(vec)))

    def get_rows(self, which, row0, n_rows):
        """(lin, vec) of local rows [row0, row0 + n_rows); which: 0 = w, 1 = n, 2 = z (ftrl_get_rows)"""
        lin = np.zeros(n_rows, np.float32)
        vec = np.zeros((n_rows, self.row_len), np.float32) if self.row_len else None
        self._check(self.lib.ftrl_get_rows(self.h, int(which), int(row0), int(n_rows), _np_ptr(lin), _np_ptr(vec)))
        return lin, vec

    @property
    def bias(self):
        return float(self._get(0)[0][0])

    @property
    def lin_w(self):
        return self._get(0)[1]

    @property
    def vec_w(self):
        return self._get(0)[2]

    def auc(self, scores, labels) -> float:
        """ROC AUC on the device (ftrl_eval_auc): ties get average ranks"""
        sc = np.ascontiguousarray(scores, np.float32)
        la = np.ascontiguousarray(labels, np.int32)
        out = C.c_double(0.0)
        self._check(self.lib.ftrl_eval_auc(self.h, len(sc), _np_ptr(sc), _np_ptr(la), C.byref(out)))
        return float(out.value)

    def has_zero_weights(self) -> bool:
        out = C.c_int(0)
        self._check(self.lib.ftrl_has_zero_weights(self.h, C.byref(out)))
        return bool(out.value)

    # -- model files --------------------------------------------------------------------
    def save_compressed_model(self, path, compress_level=10):
        self._check(self.lib.ftrl_save_model(self.h, os.fsencode(path), int(compress_level)))

    def load_compressed_model(self, path):
        self._check(self.lib.ftrl_load_model(self.h, os.fsencode(path)))

    def save_model(self, path):
        self._check(self.lib.ftrl_save_model_text(self.h, os.fsencode(path)))

    def load_model(self, path):
        self._check(self.lib.ftrl_load_model_text(self.h, os.fsencode(path)))

    # -- measurement ----------------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        self._check(self.lib.ftrl_set_stream(self.h, C.c_void_p(cuda_stream)))

    def profile_enable(self, on=True):
        self._check(self.lib.ftrl_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._check(self.lib.ftrl_profile_reset(self.h))

    def profile(self) -> dict:
        out = {}
        i = 0
        while True:
            name, ms, n = C.c_char_p(), C.c_double(), C.c_int64()
            if self.lib.ftrl_profile_read(self.h, i, C.byref(name), C.byref(ms), C.byref(n)) != 0:
                break
            out[name.value.decode()] = {"ms": ms.value, "launches": n.value}
            i += 1
        return out

    def randomize_state(self, seed=1, z_scale=300.0, n_lo=0.5, n_hi=3.0):
        self._check(self.lib.ftrl_randomize_state(self.h, int(seed), z_scale, n_lo, n_hi))

    # -- multi-GPU (feature-sharded tables) ------------------------------------------------
    def export_peer_blob(self) -> bytes:
        buf = C.create_string_buffer(PEER_BLOB_BYTES)
        self._check(self.lib.ftrl_export_peer_blob(self.h, buf))
        return buf.raw

    def attach_peers(self, blobs) -> None:
        """blobs: the PEER_BLOB_BYTES-sized blobs of ranks 0..world_size-1, in rank order"""
        raw = b"".join(bytes(b) for b in blobs)
        assert len(raw) == PEER_BLOB_BYTES * self.cfg.world_size
        self._check(self.lib.ftrl_attach_peers(self.h, raw))

    @property
    def n_local(self) -> int:
        """rows of lin / vec held by this rank (all of them on a single GPU)"""
        g, r = max(1, self.cfg.world_size), self.cfg.rank
        return (self.cfg.n_feats - r + g - 1) // g if g > 1 else self.cfg.n_feats

    def last_batch_stats(self) -> dict:
        s = BatchStats()
        self._check(self.lib.ftrl_last_batch_stats(self.h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in BatchStats._fields_ if n != "reserved"}


# ---- feature-sharded helpers -----------------------------------------------------------------------
def shard_state(st: dict, world: int, rank: int) -> dict:
    """the rows rank, rank + world, ... of a full model state (bias is replicated)"""
    return {k: (v if k == "bias" else np.ascontiguousarray(v[rank::world])) for k, v in st.items()}


def merge_states(shards: list) -> dict:
    """inverse of shard_state"""
    world = len(shards)
    out = {"bias": shards[0]["bias"]}
    for k in shards[0]:
        if k == "bias":
            continue
        n = sum(len(s[k]) for s in shards)
        full = np.zeros((n,) + shards[0][k].shape[1:], shards[0][k].dtype)
        for r, s in enumerate(shards):
            full[r::world] = s[k]
        out[k] = full
    return out


class LogicalShards:
    """`world` feature shards inside ONE process (one handle per shard, on the given devices -- by default
    all on device 0).  Used by the tests to check that the sharded exchange is invariant to the number of
    shards without needing several GPUs; real multi-GPU runs use one process per GPU (see bench.py)."""

    def __init__(self, world, devices=None, **kw):
        devices = devices or [0] * world
        self.world = world
        self.models = [FtrlModel(rank=r, world_size=world, device=devices[r], **kw) for r in range(world)]
        blobs = [m.export_peer_blob() for m in self.models]
        for m in self.models:
            m.attach_peers(blobs)

    def set_state(self, st):
        for r, m in enumerate(self.models):
            m.set_state(shard_state(st, self.world, r))

    def get_state(self):
        return merge_states([m.get_state() for m in self.models])

    def train(self, batches):
        """batches[r]: the CSR dict of rank r's share of the global minibatch.  All ranks are enqueued
        before any is synchronised (the device-side barriers need every rank in flight)."""
        pend = [m.train(**b, sync=False) for m, b in zip(self.models, batches)]
        err = None
        for m in self.models:
            try:
                m.sync()
            except FtrlError as e:  # keep draining the other ranks, then report
                err = e
        if err:
            raise err
        return [(lg, float(ls[0])) for lg, ls in pend]

    def close(self):
        for m in self.models:
            m.close()
