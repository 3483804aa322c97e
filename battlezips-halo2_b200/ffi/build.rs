// Link against the in-tree CUDA library (battlezips-halo2_b200/lib/libbzhalo2.so).
fn main() {
    let dir = std::env::var("BZHALO2_LIB_DIR").unwrap_or_else(|_| "../lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=bzhalo2");
    println!("cargo:rerun-if-env-changed=BZHALO2_LIB_DIR");
}
