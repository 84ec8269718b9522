// Links against libb200ida.so (built by `make -C rust_ida_b200/csrc`).  Set B200IDA_LIB_DIR to its directory.
fn main() {
    let dir = std::env::var("B200IDA_LIB_DIR").unwrap_or_else(|_| "../../rust_ida_b200/csrc".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=b200ida");
    println!("cargo:rerun-if-env-changed=B200IDA_LIB_DIR");
}
