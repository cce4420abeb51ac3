"""ctypes front-end for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

backend "oracle" -> oracle/libftrl_oracle.so   (plain-C restatement, oracle/ftrl_oracle.c)
backend "ref"    -> oracle/_ref/libftrl_ref.so (the reference's own classes, oracle/ref_shim.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libftrl_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libftrl_ref.so")
REF_MAIN = os.path.join(HERE, "_ref", "main")

MODEL_TYPES = {"LR": 0, "FM": 1, "FFM": 2}

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """(Re)build the checkers; `ref` needs /root/reference and is skipped without it."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", HERE, "-j8"] + targets, check=True, stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


_libs: dict = {}


def _load(backend: str):
    if backend in _libs:
        return _libs[backend]
    if backend == "oracle":
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        lib = C.CDLL(ORACLE_SO)
        p = "ftrl_oracle_"
    elif backend == "ref":
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        lib = C.CDLL(REF_SO)
        p = "ftrl_ref_"
    else:
        raise ValueError(backend)
    fn = lambda name: getattr(lib, p + name)  # noqa: E731
    fn("create").restype = C.c_void_p
    if backend == "oracle":
        fn("create").argtypes = [C.c_int] * 4 + [C.c_float] * 4
        for name in ("bias",):
            fn(name).restype = C.POINTER(C.c_float)
            fn(name).argtypes = [C.c_void_p]
        for name in ("lin", "vec"):
            fn(name).restype = C.POINTER(C.c_float)
            fn(name).argtypes = [C.c_void_p, C.c_int]
        fn("train_batch_csr").restype = C.c_double
        fn("train_batch_csr").argtypes = [C.c_void_p, C.c_int64, _i64p, _i32p, _i32p, _f32p, _i32p, C.c_void_p]
        fn("weight").argtypes = [C.c_void_p, C.c_float, C.c_float]
    else:
        fn("create").argtypes = [C.c_int] * 4 + [C.c_float] * 4 + [C.c_int]
        fn("get_bias").argtypes = [C.c_void_p, _f32p]
        fn("set_bias").argtypes = [C.c_void_p, _f32p]
        for name in ("get_lin", "set_lin", "get_vec", "set_vec"):
            fn(name).argtypes = [C.c_void_p, C.c_int, _f32p]
        fn("stage_csr").argtypes = [C.c_void_p, C.c_int64, _i64p, _i32p, _i32p, _f32p, _i32p]
        fn("train_staged").restype = C.c_double
        fn("train_staged").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_double)]
        fn("weight").argtypes = [C.c_void_p, C.c_float, C.c_float]
        for name in ("save_compressed",):
            fn(name).argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        for name in ("load_compressed", "save_text", "load_text"):
            fn(name).argtypes = [C.c_void_p, C.c_char_p]
    fn("destroy").argtypes = [C.c_void_p]
    fn("row_len").restype = C.c_int64
    fn("row_len").argtypes = [C.c_void_p]
    fn("train").restype = C.c_float
    fn("train").argtypes = [C.c_void_p, C.c_int, _i32p, _i32p, _f32p, C.c_int]
    fn("predict").restype = C.c_float
    fn("predict").argtypes = [C.c_void_p, C.c_int, _i32p, _i32p, _f32p, C.c_int]
    fn("train_csr").restype = C.c_double
    fn("train_csr").argtypes = [C.c_void_p, C.c_int64, _i64p, _i32p, _i32p, _f32p, _i32p, C.c_void_p]
    fn("predict_csr").restype = C.c_double
    fn("predict_csr").argtypes = [C.c_void_p, C.c_int64, _i64p, _i32p, _i32p, _f32p, C.c_void_p, C.c_int, C.c_void_p]
    fn("loss").restype = C.c_double
    fn("loss").argtypes = [C.c_int, C.c_double]
    fn("sigmoid").restype = C.c_float
    fn("sigmoid").argtypes = [C.c_float]
    fn("sgn").restype = C.c_float
    fn("sgn").argtypes = [C.c_float]
    fn("weight").restype = C.c_float
    _libs[backend] = (lib, p)
    return _libs[backend]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class CpuModel:
    """One LR/FM/FFM model on the CPU, either checker behind the same interface."""

    def __init__(self, backend, model_type, n_feats, n_fields=1, n_factors=1,
                 w_alpha=1e-4, w_beta=1.0, w_l1=0.1, w_l2=5.0, fast_init=True):
        self.backend = backend
        self.lib, self.p = _load(backend)
        self.model_type = model_type.upper()
        self.n_feats, self.n_fields, self.k = int(n_feats), int(n_fields), int(n_factors)
        mt = MODEL_TYPES[self.model_type]
        args = [mt, self.n_feats, self.n_fields, self.k, w_alpha, w_beta, w_l1, w_l2]
        if backend == "ref":
            args.append(1 if fast_init else 0)
        self.h = self._fn("create")(*args)
        if not self.h:
            raise RuntimeError("create failed")
        self.row_len = int(self._fn("row_len")(self.h))

    def _fn(self, name):
        return getattr(self.lib, self.p + name)

    def close(self):
        if getattr(self, "h", None):
            self._fn("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state ---------------------------------------------------------
    def get_state(self) -> dict:
        names = ("w", "n", "z")
        st = {}
        if self.backend == "oracle":
            st["bias"] = np.ctypeslib.as_array(self._fn("bias")(self.h), shape=(3,)).copy()
            for t, nm in enumerate(names):
                st["lin_" + nm] = np.ctypeslib.as_array(self._fn("lin")(self.h, t), shape=(self.n_feats,)).copy()
                if self.row_len:
                    st["vec_" + nm] = np.ctypeslib.as_array(
                        self._fn("vec")(self.h, t), shape=(self.n_feats, self.row_len)).copy()
        else:
            b = np.zeros(3, np.float32)
            self._fn("get_bias")(self.h, b)
            st["bias"] = b
            for t, nm in enumerate(names):
                a = np.zeros(self.n_feats, np.float32)
                self._fn("get_lin")(self.h, t, a)
                st["lin_" + nm] = a
                if self.row_len:
                    v = np.zeros((self.n_feats, self.row_len), np.float32)
                    self._fn("get_vec")(self.h, t, v)
                    st["vec_" + nm] = v
        return st

    def set_state(self, st: dict) -> None:
        names = ("w", "n", "z")
        if self.backend == "oracle":
            if "bias" in st:
                np.ctypeslib.as_array(self._fn("bias")(self.h), shape=(3,))[:] = st["bias"]
            for t, nm in enumerate(names):
                if "lin_" + nm in st:
                    np.ctypeslib.as_array(self._fn("lin")(self.h, t), shape=(self.n_feats,))[:] = st["lin_" + nm]
                if self.row_len and "vec_" + nm in st:
                    np.ctypeslib.as_array(self._fn("vec")(self.h, t), shape=(self.n_feats, self.row_len))[:] = \
                        np.asarray(st["vec_" + nm], np.float32).reshape(self.n_feats, self.row_len)
        else:
            if "bias" in st:
                self._fn("set_bias")(self.h, np.ascontiguousarray(st["bias"], np.float32))
            for t, nm in enumerate(names):
                if "lin_" + nm in st:
                    self._fn("set_lin")(self.h, t, np.ascontiguousarray(st["lin_" + nm], np.float32))
                if self.row_len and "vec_" + nm in st:
                    self._fn("set_vec")(self.h, t, np.ascontiguousarray(st["vec_" + nm], np.float32).reshape(-1))

    # ---- per-sample ----------------------------------------------------
    @staticmethod
    def _sample(field, feat, val):
        return (np.ascontiguousarray(field, np.int32), np.ascontiguousarray(feat, np.int32),
                np.ascontiguousarray(val, np.float32))

    def train(self, field, feat, val, label) -> float:
        f, i, x = self._sample(field, feat, val)
        return float(self._fn("train")(self.h, len(i), f, i, x, int(label)))

    def predict(self, field, feat, val, output_prob=False) -> float:
        f, i, x = self._sample(field, feat, val)
        return float(self._fn("predict")(self.h, len(i), f, i, x, int(output_prob)))

    # ---- CSR blocks ----------------------------------------------------
    @staticmethod
    def _csr(row_ptr, field, feat, val, label):
        return (np.ascontiguousarray(row_ptr, np.int64), np.ascontiguousarray(field, np.int32),
                np.ascontiguousarray(feat, np.int32), np.ascontiguousarray(val, np.float32),
                None if label is None else np.ascontiguousarray(label, np.int32))

    def train_csr(self, row_ptr, field, feat, val, label):
        """sequential reference semantics; returns (logits, loss_sum)"""
        rp, f, i, x, y = self._csr(row_ptr, field, feat, val, label)
        n = len(rp) - 1
        out = np.zeros(n, np.float32)
        ls = self._fn("train_csr")(self.h, n, rp, f, i, x, y, _ptr(out))
        return out, float(ls)

    def predict_csr(self, row_ptr, field, feat, val, label=None, output_prob=False):
        rp, f, i, x, y = self._csr(row_ptr, field, feat, val, label)
        n = len(rp) - 1
        out = np.zeros(n, np.float32)
        ls = self._fn("predict_csr")(self.h, n, rp, f, i, x, _ptr(y), int(output_prob), _ptr(out))
        return out, float(ls)

    def train_batch_csr(self, row_ptr, field, feat, val, label):
        """DERIVED minibatch semantics (oracle backend only); returns (logits, loss_sum)"""
        assert self.backend == "oracle"
        rp, f, i, x, y = self._csr(row_ptr, field, feat, val, label)
        n = len(rp) - 1
        out = np.zeros(n, np.float32)
        ls = self._fn("train_batch_csr")(self.h, n, rp, f, i, x, y, _ptr(out))
        return out, float(ls)

    # ---- reference-only: threaded epoch for the CPU baseline -------------
    def stage_csr(self, row_ptr, field, feat, val, label):
        assert self.backend == "ref"
        rp, f, i, x, y = self._csr(row_ptr, field, feat, val, label)
        self._fn("stage_csr")(self.h, len(rp) - 1, rp, f, i, x, y)

    def train_staged(self, n_threads=1, shuffle=False, seed=0):
        """returns (seconds, mean_loss) for one pass over the staged samples"""
        assert self.backend == "ref"
        loss = C.c_double(0.0)
        secs = self._fn("train_staged")(self.h, int(n_threads), int(shuffle), int(seed), C.byref(loss))
        return float(secs), float(loss.value)

    # ---- scalars ---------------------------------------------------------
    def weight(self, n, z) -> float:
        return float(self._fn("weight")(self.h, float(n), float(z)))


def scalar_fns(backend: str):
    lib, p = _load(backend)
    return {k: getattr(lib, p + k) for k in ("loss", "sigmoid", "sgn")}
