"""Builds libftrl_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc, and the C++17
host programs under host/.  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libftrl_b200.so")
MAIN = os.path.join(HERE, "main")
PARSER_LIB = os.path.join(HERE, "libftrl_host_parser.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(d: str) -> list[str]:
    out = []
    for root, _, files in os.walk(d):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".cpp", ".hpp"))]
    return out


def build_library(force: bool = False, verbose: bool = False) -> str:
    inc = os.path.join(os.path.dirname(HERE), "include", "ftrl_b200.h")
    srcs = _sources(CSRC) + [inc]
    if not force and _newer(LIB, srcs):
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB, os.path.join(CSRC, "ftrl_b200.cu"), "-l:libzstd.so.1"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


def build_host(force: bool = False) -> str | None:
    """C++17 drop-in `main` (CLI, libsvm/libffm parser -> pinned CSR, epoch drivers)."""
    src = os.path.join(HOST, "main.cpp")
    if not os.path.exists(src):
        return None
    srcs = _sources(HOST) + [os.path.join(os.path.dirname(HERE), "include", "ftrl_b200.h")]
    if not force and _newer(MAIN, srcs + [LIB]):
        return MAIN
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
    inc = ["-I", os.path.join(os.path.dirname(HERE), "include")]
    cmd = [cxx, "-O3", "-std=c++17", "-pthread"] + inc + ["-o", MAIN, os.path.join(HOST, "main.cpp"),
                                                          "-L", HERE, "-l:libftrl_b200.so", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True, cwd=HOST)
    # the parser alone as a tiny C library: lets the CPU test-suite check it against the reference's Parser
    subprocess.run([cxx, "-O3", "-std=c++17", "-pthread", "-fPIC", "-shared"] + inc +
                   ["-o", PARSER_LIB, os.path.join(HOST, "parser_capi.cpp")], check=True, cwd=HOST)
    return MAIN


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_host(force=True))
