// Links libhalo2_b200.so (built by `python -m halo2_b200.build`; path via HALO2_B200_LIB_DIR).
fn main() {
    let dir = std::env::var("HALO2_B200_LIB_DIR").expect("set HALO2_B200_LIB_DIR to .../halo2_b200/_lib");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=halo2_b200");
}
