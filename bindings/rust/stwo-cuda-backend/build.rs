// Points the linker at libstwo_cuda.so.  STWO_CUDA_LIB_DIR = the directory `python __graft_entry__.py` (or `make -C
// stwo-brainfuck_b200/csrc`) leaves the library in; the default is that directory relative to this crate.
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("STWO_CUDA_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../../stwo-brainfuck_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=stwo_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=STWO_CUDA_LIB_DIR");
}
