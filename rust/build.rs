// Links the C ABI of include/rt1w.h (raytracing-1w_b200/_build/librt1w.so; the CUDA runtime is linked statically into it).
fn main() {
    let dir = std::env::var("RT1W_LIB_DIR").expect("set RT1W_LIB_DIR to raytracing-1w_b200/_build");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rt1w");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=RT1W_LIB_DIR");
}
