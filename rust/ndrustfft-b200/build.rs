//! Builds libndfft_b200.so with nvcc (sm_100a) from the repository's Makefile and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let repo = manifest.join("../..").canonicalize().expect("repository root");
    let status = Command::new("make")
        .arg("-C")
        .arg(&repo)
        .arg("lib")
        .status()
        .expect("failed to run make (needs nvcc with sm_100a support)");
    assert!(status.success(), "nvcc build of libndfft_b200.so failed");
    let libdir = repo.join("ndrustfft_b200/lib");
    println!("cargo:rustc-link-search=native={}", libdir.display());
    println!("cargo:rustc-link-lib=dylib=ndfft_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", libdir.display());
    println!("cargo:rerun-if-changed={}", repo.join("ndrustfft_b200/csrc").display());
    println!("cargo:rerun-if-changed={}", repo.join("include/ndfft_b200.h").display());
}
