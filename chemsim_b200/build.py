"""Builds libchemsim_lbm.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m chemsim_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU
box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libchemsim_lbm.so")
# one translation unit per collision operator for the fused step kernels (compiled in parallel)
SOURCES = ["step_bgk.cu", "step_trt.cu", "step_regularized.cu", "step_kbc.cu", "kernels.cu", "lattice.cu"]
HEADERS = ["d2q9.cuh", "consts.hpp", "kernels.cuh", "step_decl.cuh", "step_impl.cuh", "step2_impl.cuh", "nccl_dyn.h",
           os.path.join("..", "..", "include", "chemsim_lbm.h")]
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                      # parity build: never contract a*b+c (d2q9.cuh also uses *_rn intrinsics)
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
    "-Xptxas", "-v",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-shared"]


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_and_link(out: str, defines: list[str], log_path: str, tag: str) -> None:
    """nvcc -c every translation unit in parallel (one process each), then link the shared library."""
    from concurrent.futures import ThreadPoolExecutor
    env = dict(os.environ)
    env["PATH"] = "/usr/bin:" + env.get("PATH", "")     # plain system g++ as nvcc's host compiler
    objdir = os.path.join(OBJ_DIR, tag)
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src: str):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-ccbin", "g++", "-c", "-o", obj,
               os.path.join(CSRC, src)]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        return obj, cmd, res

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    log = ""
    for obj, cmd, res in results:
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
    failed = [r for r in results if r[2].returncode != 0]
    if not failed:
        cmd = [nvcc(), *LINK_FLAGS, "-ccbin", "g++", "-o", out, *[r[0] for r in results], "-ldl"]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        if res.returncode != 0:
            failed = [(out, cmd, res)]
    with open(log_path, "w") as fh:
        fh.write(log)
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(r[2].stdout + r[2].stderr for r in failed))


def build_variant(name: str, defines: list[str]) -> str:
    """Experimental build with extra -D switches -> chemsim_b200/libchemsim_lbm_<name>.so
    (select at run time with CHEMSIM_LBM_LIB=<path>; used by tools/variants.py)."""
    out = os.path.join(HERE, f"libchemsim_lbm_{name}.so")
    _compile_and_link(out, defines, os.path.join(HERE, f"build_{name}.log"), name)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    _compile_and_link(LIB, [], os.path.join(HERE, "build.log"), "default")
    if verbose:
        with open(os.path.join(HERE, "build.log")) as fh:
            print(fh.read())
    return LIB


HARNESS = os.path.join(HERE, "cpp", "main_rs_harness")


def build_harness(force: bool = False) -> str:
    """The C++ host mirror's driver (cpp/main_rs_harness.cpp), linked against the C-ABI library."""
    build()
    src = [os.path.join(HERE, "cpp", n) for n in ("main_rs_harness.cpp", "lbm.hpp")]
    if not force and os.path.exists(HARNESS) and all(os.path.getmtime(s) < os.path.getmtime(HARNESS) for s in src + [LIB]):
        return HARNESS
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", HARNESS, src[0], "-L" + HERE,
           "-lchemsim_lbm", "-Wl,-rpath,$ORIGIN/.."]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return HARNESS


MULTI_HARNESS = os.path.join(HERE, "cpp", "multi_gpu_harness")


def build_multi_harness(force: bool = False) -> str:
    """Single-process multi-GPU driver (cpp/lbm_multi.hpp + cpp/multi_gpu_harness.cpp)."""
    build()
    src = [os.path.join(HERE, "cpp", n) for n in ("multi_gpu_harness.cpp", "lbm_multi.hpp", "lbm.hpp")]
    if not force and os.path.exists(MULTI_HARNESS) and all(
            os.path.getmtime(s) < os.path.getmtime(MULTI_HARNESS) for s in src + [LIB]):
        return MULTI_HARNESS
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-o", MULTI_HARNESS, src[0],
           "-L" + HERE, "-lchemsim_lbm", "-Wl,-rpath,$ORIGIN/.."]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return MULTI_HARNESS


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)
    print(build_harness(force="--force" in sys.argv))
    print(build_multi_harness(force="--force" in sys.argv))
