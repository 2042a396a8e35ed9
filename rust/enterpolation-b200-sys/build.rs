//! Tells cargo where `libenterp_b200.so` lives. The library is built from `enterpolation_b200/csrc` (`make`, nvcc
//! for sm_100a); point `ENTERP_B200_LIB_DIR` at the directory that holds it, or leave it unset to use the in-tree
//! location relative to this crate (`../../enterpolation_b200`).
use std::env;
use std::path::PathBuf;

fn main() {
    println!("cargo:rerun-if-env-changed=ENTERP_B200_LIB_DIR");
    let dir = env::var_os("ENTERP_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|| {
        PathBuf::from(env::var_os("CARGO_MANIFEST_DIR").expect("cargo sets CARGO_MANIFEST_DIR"))
            .join("..")
            .join("..")
            .join("enterpolation_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=enterp_b200");
    // let dependants find the header and set an rpath
    println!("cargo:lib_dir={}", dir.display());
    println!("cargo:include={}", dir.join("..").join("include").display());
}
