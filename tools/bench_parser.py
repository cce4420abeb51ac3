"""Text front-end throughput: host/parser.hpp (from_chars, multi-threaded) against the
reference's own Parser (oracle/_ref, one thread) on Criteo-shaped libffm text.  CPU only."""
import ctypes as C
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ftrl_ffm_b200 as pkg

n_lines = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
b = pkg.synth.criteo_batch(n_lines, 39, 10_000_000, seed=42)
path = "/tmp/bench_parser.ffm"
pkg.synth.write_text(b, path, "libffm")
text = open(path, "rb").read()
print(f"{n_lines} lines, {len(text) / n_lines:.0f} B/line, {len(text) / 1e6:.1f} MB")

lib = C.CDLL(os.path.join(os.path.dirname(pkg.binding.LIB_PATH), "libftrl_host_parser.so"))
lib.host_parse_text.restype = C.c_int64
lib.host_parse_text.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int]
for th in (1, 2, 4, 8, os.cpu_count() or 1):
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        n = lib.host_parse_text(text, len(text), 1, th)
        best = min(best, time.perf_counter() - t0)
    assert n == n_lines
    print(f"host/parser.hpp  threads={th:3d}: {n_lines / best / 1e6:6.2f} M lines/s  {len(text) / best / 1e9:5.2f} GB/s")

try:
    from oracle.cpu_model import have_ref, REF_SO
    ref = C.CDLL(REF_SO) if have_ref() else None
except Exception:
    ref = None
if ref is not None and hasattr(ref, "ftrl_ref_parse_line"):
    lines = text.split(b"\n")[:20000]
    fn = ref.ftrl_ref_parse_line
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    fld = (C.c_int32 * 256)(); ft = (C.c_int32 * 256)(); val = (C.c_float * 256)(); lab = C.c_int()
    t0 = time.perf_counter()
    for ln in lines:
        fn(1, ln, 256, fld, ft, val, C.byref(lab))
    dt = time.perf_counter() - t0
    print(f"reference Parser threads=  1: {len(lines) / dt / 1e6:6.3f} M lines/s (includes the ctypes call per line)")
