// UNVERIFIED (no cargo in the build image).  Links libtfhe_b200.so built by `python -c "import __graft_entry__ as g; g.build()"`.
fn main() {
    let dir = std::env::var("TFHE_B200_LIB_DIR").unwrap_or_else(|_| "../../tfhe-research_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=tfhe_b200");
    println!("cargo:rerun-if-env-changed=TFHE_B200_LIB_DIR");
}
