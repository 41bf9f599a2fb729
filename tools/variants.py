#!/usr/bin/env python
"""Builds experimental variants of the step kernel (compile-time switches) for an A/B sweep on the GPU.
    python tools/variants.py build        (here, on CPU)
    python tools/variants.py run          (under gpurun: bench each variant, print GLUPS)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {
    "base": [],
    "bulk": ["CHEMSIM_EXPERIMENT_BULK"],     # TMA bulk-copy loads (step_bulk_experiment.cuh); run with CHEMSIM_LBM_BULK=1 CHEMSIM_LBM_STEP2=0
    "nohpair": ["CHEMSIM_STEP2_HPAIR=0"],    # f32 phase A on two cells a block-width apart (scalar loads/stores) instead of adjacent pairs
    "nostcs2": ["CHEMSIM_STEP2_STORE_CS=0"], # two-step phase B with write-back stores instead of streaming stores
    "s2ty10": ["CHEMSIM_STEP2_TY=10"],       # 10-row tiles (320 threads, 3 blocks/SM, rim +22 %)
    "s2ty16": ["CHEMSIM_STEP2_TY=16"],       # two-step kernel: 16-row tiles (512 threads, 2 blocks/SM, rim +14 %); default is 8
    "s2ty32": ["CHEMSIM_STEP2_TY=32"],       # 32-row tiles (1024 threads, 1 block/SM, rim +8 %)
    "scalar": ["CHEMSIM_PACKED_STEP2=0"],    # two-step kernels without the packed f32 additions (the r02 build before F32x2)
    "pm": ["CHEMSIM_PACKED_MUL=1"],          # packed multiplications too (fma.rn.f32x2 with an opaque -0 addend, d2q9.cuh)
    "vp": ["CHEMSIM_PACKED_VEC=0x1e"],       # single-step vector kernels packed too, 64-register cap (spills)
    "vp3": ["CHEMSIM_PACKED_VEC=0x1e", "CHEMSIM_STEP_MIN_BLOCKS=3", "CHEMSIM_KBC_MIN_BLOCKS=3"],   # ... at 80 registers
    "vp2": ["CHEMSIM_PACKED_VEC=0x1e", "CHEMSIM_STEP_MIN_BLOCKS=2", "CHEMSIM_KBC_MIN_BLOCKS=2"],   # ... at 128 registers
    "kbc3": ["CHEMSIM_KBC_MIN_BLOCKS=3"],    # KBC capped at 80 registers (3 blocks/SM)
    "kbc4": ["CHEMSIM_KBC_MIN_BLOCKS=4"],    # KBC capped at 64 registers (4 blocks/SM)
    "mb5": ["CHEMSIM_STEP_MIN_BLOCKS=5"],
    "mb6": ["CHEMSIM_STEP_MIN_BLOCKS=6"],
    "t128": ["CHEMSIM_STEP_THREADS=128"],
    "t512": ["CHEMSIM_STEP_THREADS=512"],
    "stcs": ["CHEMSIM_STORE_CS"],
    "noalloc": ["CHEMSIM_LOAD_NOALLOC"],
    "stcs_noalloc": ["CHEMSIM_STORE_CS", "CHEMSIM_LOAD_NOALLOC"],
    "t128_mb10": ["CHEMSIM_STEP_THREADS=128", "CHEMSIM_STEP_MIN_BLOCKS=10"],
}

if sys.argv[1] == "build":
    from chemsim_b200 import build
    only = sys.argv[2:] or list(VARIANTS)
    for name, defs in VARIANTS.items():
        if name not in only:
            continue
        print(name, build.build_variant(name, defs))
else:
    for name in VARIANTS:
        lib = os.path.join(ROOT, "chemsim_b200", f"libchemsim_lbm_{name}.so")
        for dtype in ("f32", "f64"):
            env = dict(os.environ, CHEMSIM_LBM_LIB=lib)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu", "--dtype", dtype,
                                  "--steps", "500", "--warmup", "20"], capture_output=True, text=True, env=env)
            try:
                d = json.loads(out.stdout.strip().splitlines()[-1])
                print(f"{name:14s} {dtype} {d['value']:.2f} GLUPS frac {d['roofline']['frac']:.4f}", flush=True)
            except Exception:
                print(name, dtype, "FAILED", out.stderr[-300:], flush=True)
