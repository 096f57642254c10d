// build.rs -- compiles the CUDA engine and links it into the `quaru` crate.
// UNVERIFIED in this repository's CI: the build image has no Rust toolchain.
// Equivalent to `make -C quaru_b200/csrc` followed by linking libquaru_b200.so.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = root.join("quaru_b200").join("csrc");
    let status = Command::new("make").arg("-C").arg(&csrc).arg("-j8").status().expect("cannot run make");
    assert!(status.success(), "nvcc build of libquaru_b200.so failed (sm_100a, CUDA >= 12.8 required)");
    let lib = root.join("quaru_b200").join("lib");
    println!("cargo:rustc-link-search=native={}", lib.display());
    println!("cargo:rustc-link-lib=dylib=quaru_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib.display());
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include").join("quaru_b200.h").display());
}
