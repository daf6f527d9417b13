// Links the prebuilt C-ABI library (built by `python -c 'import __graft_entry__ as g; g.build()'`).
fn main() {
    let dir = std::env::var("LBFGSB200_LIB_DIR").unwrap_or_else(|_| "../..".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=lbfgsb200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=LBFGSB200_LIB_DIR");
}
