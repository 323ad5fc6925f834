// build.rs of the reference crate once the shim is in: link libipb200.so (the CUDA runtime is linked statically inside it).
fn main() {
    let dir = std::env::var("IPB200_LIB_DIR").expect("set IPB200_LIB_DIR to the directory that holds libipb200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=ipb200");
    println!("cargo:rerun-if-env-changed=IPB200_LIB_DIR");
}
