// Links libbpg.so (built in-tree by `python -m bulletproof_gadgets_b200.build`); BPG_LIB_DIR overrides the location.
fn main() {
    let dir = std::env::var("BPG_LIB_DIR").unwrap_or_else(|_| "../bulletproof_gadgets_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=bpg");
    println!("cargo:rerun-if-env-changed=BPG_LIB_DIR");
}
