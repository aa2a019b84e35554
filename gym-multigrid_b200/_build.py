"""Build the CUDA shared library in-tree with nvcc for sm_100a (no torch, no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libmultigrid_b200.so")
SOURCES = ["collect_kernels.cu", "collect_rollout_kernels.cu", "map_kernels.cu", "view_kernels.cu", "wildfire_kernels.cu", "generic_kernels.cu", "render_kernels.cu", "policy_kernels.cu", "astar_host.cu", "host_transport.cpp", "mg_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "0",
              "-shared", "-Xcompiler", "-fPIC"]


def _deps():
    inc = os.path.join(os.path.dirname(PKG_DIR), "include", "multigrid_b200.h")
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [inc]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("MG_NVCC_EXTRA", "").split(), "-o", LIB_PATH, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH
