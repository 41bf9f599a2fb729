// build.rs — compiles the CUDA side (../chemsim_b200/csrc) for sm_100a with nvcc and links it.
// No Triton, no multi-backend dispatch, no CPU fallback: if nvcc is missing the build fails.
use std::{env, path::PathBuf, process::Command};

// one translation unit per collision operator for the fused step kernels (chemsim_b200/build.py compiles them in parallel)
const SOURCES: [&str; 6] = ["step_bgk.cu", "step_trt.cu", "step_regularized.cu", "step_kbc.cu", "kernels.cu", "lattice.cu"];
const HEADERS: [&str; 7] = ["d2q9.cuh", "consts.hpp", "kernels.cuh", "step_decl.cuh", "step_impl.cuh", "step2_impl.cuh", "nccl_dyn.h"];

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("..").join("chemsim_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libchemsim_lbm.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let status = Command::new(&nvcc)
        .args(&[
            "-gencode", "arch=compute_100a,code=sm_100a",
            "-O3", "-std=c++17", "-lineinfo",
            "-fmad=false",                                  // parity build: never contract a*b+c
            "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
            "-cudart", "static", "-shared", "-o",
        ])
        .arg(&lib)
        .args(SOURCES.iter().map(|f| csrc.join(f)))
        .arg("-ldl")
        .status()
        .expect("nvcc not found (set NVCC=/path/to/nvcc)");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=chemsim_lbm");
    println!("cargo:rustc-env=CHEMSIM_LBM_LIB_DIR={}", out.display());
    for f in SOURCES.iter().chain(HEADERS.iter()) {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", manifest.join("../include/chemsim_lbm.h").display());
}
