// build.rs — links the prebuilt libraries (`make host backend` at the repository root) into the
// reference crate. RT_B200_DIR points at rsoderh-raytracing_b200/ (where librt_backend.so and
// librt_host.so are); the rpath makes the crate's binaries find them at run time.
fn main() {
    let dir = std::env::var("RT_B200_DIR").unwrap_or_else(|_| "../../rsoderh-raytracing_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=rt_backend");
    println!("cargo:rustc-link-lib=dylib=rt_host");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=RT_B200_DIR");
    println!("cargo:rerun-if-changed=src/rt_sys.rs");
}
