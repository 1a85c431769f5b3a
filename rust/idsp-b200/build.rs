// UNVERIFIED SOURCE (no rustc in the build image).
fn main() {
    // libidsp_b200.so is built by `python -m idsp_b200.build` (nvcc, sm_100a)
    let dir = std::env::var("IDSP_B200_LIB_DIR").unwrap_or_else(|_| "../../idsp_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=idsp_b200");
}
