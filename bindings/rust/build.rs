// Links libb2m.so (built by `make -C marlin_b200/csrc -j`).  B2M_LIB_DIR = directory holding the library.
fn main() {
    let dir = std::env::var("B2M_LIB_DIR").unwrap_or_else(|_| "../../marlin_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=b2m");
    println!("cargo:rerun-if-env-changed=B2M_LIB_DIR");
}
