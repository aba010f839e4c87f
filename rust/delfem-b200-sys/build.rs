// Links the prebuilt C-ABI library. DELFEM_B200_LIB_DIR points at del-fem_b200/lib (default: ../../del-fem_b200/lib).
fn main() {
    let dir = std::env::var("DELFEM_B200_LIB_DIR").unwrap_or_else(|_| "../../del-fem_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=delfem_b200");
    println!("cargo:rerun-if-env-changed=DELFEM_B200_LIB_DIR");
}
