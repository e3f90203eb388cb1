// Links libsvk.so (built by `make -C snark_verifier_axiom_b200/csrc`).  SVK_LIB_DIR points at the directory holding it.
fn main() {
    let dir = std::env::var("SVK_LIB_DIR").unwrap_or_else(|_| "../../snark_verifier_axiom_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=svk");
    println!("cargo:rerun-if-env-changed=SVK_LIB_DIR");
}
