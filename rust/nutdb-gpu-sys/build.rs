// UNCOMPILED (no Rust toolchain in the build image) -- see ../README.md
//
// Compiles the kernels and the C ABI with nvcc for sm_100a -- that architecture only: no other -gencode, no CPU
// fallback (nutdb_gpu_ctx_create returns NULL without a B200).  NUTDB_GPU_ROOT points at the checkout that holds
// nutdb_b200/csrc and include/ (default: two levels up from this crate).
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(std::env::var("OUT_DIR").unwrap());
    let root = std::env::var("NUTDB_GPU_ROOT")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env!("CARGO_MANIFEST_DIR")).join("../.."));
    let csrc = root.join("nutdb_b200/csrc");
    // the parser bytecode (parse_program.h) is generated from the grammar description
    let st = Command::new("python3").arg(csrc.join("gen_parse_program.py")).status().expect("python3");
    assert!(st.success(), "gen_parse_program.py failed");
    let nvcc = std::env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let lib = out.join("libnutdb_gpu.so");
    let st = Command::new(nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"])
        .args(["-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("nutdb_gpu.cu"))
        .arg(csrc.join("hydrate.cpp"))
        .arg(csrc.join("dispatch.cpp"))
        .status()
        .expect("nvcc not found: the GPU parser needs the CUDA 12.9 toolkit");
    assert!(st.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=nutdb_gpu");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/nutdb_gpu.h").display());
}
