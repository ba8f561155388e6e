// Links libgfi.so (built by `make -C vectordb-from-scratch_b200/csrc`).  GFI_LIB_DIR overrides the default
// location relative to this crate.
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("GFI_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../vectordb-from-scratch_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=gfi");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=GFI_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/gfi.h");
}
