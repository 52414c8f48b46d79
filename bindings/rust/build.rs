// build.rs — compiles the CUDA engine for sm_100a and links it into the `sameold` crate.
// NOT compiled in this repository's CI (no Rust toolchain in the build image); kept as the reference-side glue a
// maintainer would add next to crates/sameold/Cargo.toml (which today has no build script and no FFI).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("SAME_B200_ROOT").unwrap_or_else(|_| "../..".into()));
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libsame_b200.so");
    let csrc = root.join("sameold_b200/csrc");
    let status = Command::new("nvcc")
        .args([
            "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
            // bit-exactness against the Rust f32 arithmetic: no FMA contraction, IEEE div/sqrt, subnormals kept
            "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
            "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-shared", "-o",
        ])
        .arg(&lib)
        .arg(csrc.join("same_kernels.cu"))
        .arg(csrc.join("same_long.cu"))
        .arg(csrc.join("same_engine.cu"))
        .arg(csrc.join("same_multi.cu"))
        .status()
        .expect("nvcc not found: the B200 engine has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=same_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/same_engine.h").display());
}
