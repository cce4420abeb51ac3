"""CPU tests of the C++ host parser (ftrl-ffm_b200/host/parser.hpp) against the reference's own
Parser classes (through oracle/_ref when present) and the reference's unit test tests/test_data.cpp."""
import ctypes as C
import os

import numpy as np
import pytest

import ftrl_ffm_b200 as pkg
from ftrl_ffm_b200.build import PARSER_LIB, build_host
from oracle.cpu_model import REF_SO, have_ref

# tests/common.h:12-24 of the reference
REF_TEST_SAMPLES = (
    "0 0:1:1 1:13:1 2:21:1 3:31:1\n"
    "1 0:4:1 1:11:1 2:23:1 3:32:1\n"
    "1 0:2:1 1:13:1 2:25:1 3:34:1\n"
    "0 0:1:1 1:14:1 2:21:1 3:32:1\n"
    "0 0:2:1 1:15:1 2:22:1 3:34:1\n"
    "1 0:4:1 1:11:1 2:21:1 3:35:1\n"
    "1 0:5:1 1:12:1 2:23:1 3:31:1\n"
    "1 0:5:1 1:12:1 2:25:1 3:38:1\n"
    "0 0:2:1 1:11:1 2:24:1 3:37:1\n"
    "1 0:1:1 1:15:1 2:22:1 3:35:1")


def parser_lib():
    if not os.path.exists(PARSER_LIB):
        build_host()
    lib = C.CDLL(PARSER_LIB)
    lib.host_parse_text.restype = C.c_int64
    lib.host_parse_text.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int]
    lib.host_parse_nnz.restype = C.c_int64
    lib.host_parse_fetch.argtypes = [C.c_void_p] * 5
    return lib


def parse(text, libffm, n_threads=1):
    lib = parser_lib()
    raw = text.encode()
    n = lib.host_parse_text(raw, len(raw), int(libffm), n_threads)
    nnz = lib.host_parse_nnz()
    rp = np.zeros(n + 1, np.int64)
    fi, fe, va, la = np.zeros(nnz, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz, np.float32), np.zeros(n, np.int32)
    lib.host_parse_fetch(*(a.ctypes.data_as(C.c_void_p) for a in (rp, fi, fe, va, la)))
    return {"row_ptr": rp, "field": fi, "feat": fe, "val": va, "label": la}


def test_reference_test_data_cpp():
    """tests/test_data.cpp:9-18: 10 samples, data[0].y == 0, x[0] == (0,1,1), x[3] == (3,31,1); partitioned
    multi-threaded loading keeps file order"""
    for nt in (1, 4):
        d = parse(REF_TEST_SAMPLES, True, nt)
        assert len(d["label"]) == 10 and d["label"][0] == 0
        assert (d["field"][0], d["feat"][0], d["val"][0]) == (0, 1, 1.0)
        assert (d["field"][3], d["feat"][3], d["val"][3]) == (3, 31, 1.0)
        assert list(d["label"]) == [0, 1, 1, 0, 0, 1, 1, 1, 0, 1]


def test_parsing_rules():
    # label > 0 -> 1 else 0; value 0 dropped; libsvm field forced to 0; '+', exponents, trailing blanks, CRLF
    d = parse("3 5:0.5 7:0 9:1e-2 \r\n-1 2:+2.5\n0 4:1\n\n", False)
    assert list(d["label"]) == [1, 0, 0]
    assert list(d["row_ptr"]) == [0, 2, 3, 4]
    assert list(d["feat"]) == [5, 9, 2, 4] and list(d["field"]) == [0, 0, 0, 0]
    np.testing.assert_array_equal(d["val"], np.array([0.5, 1e-2, 2.5, 1.0], np.float32))
    d = parse("1 0:3:1.5 12:400:-2\n", True)
    assert list(d["field"]) == [0, 12] and list(d["feat"]) == [3, 400] and list(d["val"]) == [1.5, -2.0]


def test_threads_give_identical_csr():
    rng = np.random.default_rng(0)
    b = pkg.synth.random_csr(rng, 5000, 1000, 9, max_nnz=9, oob_frac=0.0)
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".ffm", delete=False) as f:
        path = f.name
    try:
        pkg.synth.write_text(b, path, "libffm")
        text = open(path).read()
    finally:
        os.remove(path)
    one, many = parse(text, True, 1), parse(text, True, 7)
    for k in one:
        np.testing.assert_array_equal(one[k], many[k])
    keep = b["val"] != 0
    np.testing.assert_array_equal(one["feat"], b["feat"][keep])
    np.testing.assert_array_equal(one["field"], b["field"][keep])
    np.testing.assert_array_equal(one["val"], b["val"][keep])
    np.testing.assert_array_equal(one["label"], b["label"])


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_against_reference_parser_line_by_line():
    ref = C.CDLL(REF_SO)
    ref.ftrl_ref_parse_line.argtypes = [C.c_int, C.c_char_p, C.c_int] + [C.c_void_p] * 3 + [C.POINTER(C.c_int)]
    rng = np.random.default_rng(1)
    for libffm in (True, False):
        lines = []
        for _ in range(300):
            toks = [str(int(rng.integers(-2, 3)))]
            for _ in range(int(rng.integers(0, 7))):
                v = rng.choice(["1", "0", "0.25", "-3.5", "1e-3", "12.", ".5", "0.0", "7e2"])
                ft = int(rng.integers(0, 100000))
                toks.append(f"{int(rng.integers(0, 40))}:{ft}:{v}" if libffm else f"{ft}:{v}")
            lines.append(" ".join(toks))
        got = parse("\n".join(lines) + "\n", libffm, 3)
        assert len(got["label"]) == len(lines)
        fi, fe, va = np.zeros(64, np.int32), np.zeros(64, np.int32), np.zeros(64, np.float32)
        for r, line in enumerate(lines):
            lab = C.c_int(0)
            n = ref.ftrl_ref_parse_line(int(libffm), line.encode(), 64, fi.ctypes.data, fe.ctypes.data,
                                        va.ctypes.data, C.byref(lab))
            a, e = got["row_ptr"][r], got["row_ptr"][r + 1]
            assert n == e - a, line
            assert lab.value == got["label"][r]
            np.testing.assert_array_equal(got["field"][a:e], fi[:n])
            np.testing.assert_array_equal(got["feat"][a:e], fe[:n])
            np.testing.assert_array_equal(got["val"][a:e], va[:n])


def test_csr_cache_round_trip_and_invalidation(tmp_path):
    """--csr_cache: the binary image next to a data file reproduces the parsed CSR exactly and is ignored once
    the text changes (size / mtime fingerprint) or is asked for in the other format"""
    import os
    lib = parser_lib()
    lib.host_load_file.restype = C.c_int64
    lib.host_load_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    rng = np.random.default_rng(3)
    b = pkg.synth.random_csr(rng, 700, 500, 7, max_nnz=7, oob_frac=0.0)
    path = str(tmp_path / "d.ffm")
    pkg.synth.write_text(b, path, "libffm")

    def load(use_cache, libffm=1):
        hit = C.c_int(-1)
        n = lib.host_load_file(path.encode(), libffm, 3, use_cache, C.byref(hit))
        nnz = lib.host_parse_nnz()
        rp = np.zeros(n + 1, np.int64)
        fi, fe, va, la = np.zeros(nnz, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz, np.float32), np.zeros(n, np.int32)
        lib.host_parse_fetch(*(a.ctypes.data_as(C.c_void_p) for a in (rp, fi, fe, va, la)))
        return hit.value, {"row_ptr": rp, "field": fi, "feat": fe, "val": va, "label": la}

    hit0, d0 = load(0)
    assert hit0 == 0 and not os.path.exists(path + ".csr")
    hit1, d1 = load(1)            # parses and writes the image
    assert hit1 == 0 and os.path.exists(path + ".csr")
    hit2, d2 = load(1)            # served from the image
    assert hit2 == 1
    for k in d0:
        np.testing.assert_array_equal(d0[k], d1[k])
        np.testing.assert_array_equal(d0[k], d2[k])
    with open(path, "a") as f:    # the text changes: the image is stale
        f.write("1 0:1:1.0\n")
    hit3, d3 = load(1)
    assert hit3 == 0 and len(d3["label"]) == len(d0["label"]) + 1
    hit4, _ = load(1)
    assert hit4 == 1
    with open(path + ".csr", "r+b") as f:   # truncated image: ignored, rebuilt
        f.truncate(100)
    hit5, d5 = load(1)
    assert hit5 == 0 and len(d5["label"]) == len(d3["label"])


def test_decimal_fast_path_is_correctly_rounded():
    """short decimals take a one-division fast path; it must return the float std::stof returns (the correctly
    rounded fp32), including values whose nearest double and nearest float disagree in the last bit"""
    rng = np.random.default_rng(9)
    toks = []
    for _ in range(20000):
        nd = int(rng.integers(1, 8))
        ip = int(rng.integers(0, 200)) if rng.random() < 0.5 else 0
        frac = "".join(str(int(d)) for d in rng.integers(0, 10, nd))
        toks.append(f"{ip}.{frac}")
    toks += ["0.1", "0.3", "16777215.0", "16777216.5", "0.0000001", "1.0000001", "9.9999999", "0.5489", "123.456"]
    text = "".join(f"1 {i + 1}:{t}\n" for i, t in enumerate(toks))
    d = parse(text, False, 2)
    want = np.array([np.float32(t) for t in toks], np.float32)
    keep = want != 0
    np.testing.assert_array_equal(d["val"], want[keep])
    assert list(d["feat"]) == [i + 1 for i in range(len(toks)) if keep[i]]


def _fetch(lib, n):
    nnz = lib.host_parse_nnz()
    rp = np.zeros(n + 1, np.int64)
    fi, fe, va, la = np.zeros(nnz, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz, np.float32), np.zeros(n, np.int32)
    lib.host_parse_fetch(*(a.ctypes.data_as(C.c_void_p) for a in (rp, fi, fe, va, la)))
    return {"row_ptr": rp, "field": fi, "feat": fe, "val": va, "label": la}


@pytest.mark.parametrize("trailing_newline", [True, False])
@pytest.mark.parametrize("threads,block", [(1, 1 << 16), (4, 1 << 16), (7, 70_000), (3, 1 << 22)])
def test_streamed_blocks_equal_whole_file_parse(tmp_path, threads, block, trailing_newline):
    """the streaming front end of `main` (host::TextBlockReader: blocks of complete lines, parallel pread into a
    reused buffer, persistent parser threads, per-thread parts reused) yields the same CSR, in file order, as one
    parse of the whole text -- for blocks much smaller than the file, a last line without newline, and lines of very
    different lengths (one longer than a block's slice per thread)"""
    rng = np.random.default_rng(threads * 1000 + block % 97)
    b = pkg.synth.criteo_batch(3000, 39, 100_000, seed=5)
    path = str(tmp_path / "s.ffm")
    pkg.synth.write_text(b, path)
    text = open(path).read()
    lines = text.rstrip("\n").split("\n")
    # a few short, empty and very long lines in between
    long_line = "1 " + " ".join(f"{i % 39}:{i}:1.5" for i in range(6000))
    for pos in sorted(rng.integers(0, len(lines), 6).tolist(), reverse=True):
        lines.insert(pos, rng.choice(["0 3:17:1", "", long_line]))
    text = "\n".join(lines) + ("\n" if trailing_newline else "")
    open(path, "w").write(text)
    want = parse(text, True, 1)
    lib = parser_lib()
    lib.host_stream_file.restype = C.c_int64
    lib.host_stream_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_int)]
    nb = C.c_int(0)
    n = lib.host_stream_file(path.encode(), 1, threads, block, C.byref(nb))
    assert n == len(want["label"])
    got = _fetch(lib, n)
    for key in want:
        assert np.array_equal(got[key], want[key]), key
    assert nb.value >= max(1, len(text) // (block + (block >> 5) + 4096))
    assert lib.host_stream_file(str(tmp_path / "missing").encode(), 1, 2, block, None) == -1
