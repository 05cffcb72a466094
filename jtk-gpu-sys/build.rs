// libjtkgpu.so is built by `python jtk_b200/build.py` (nvcc -gencode arch=compute_100a,code=sm_100a); this script only
// tells cargo where it is.  JTK_GPU_LIB_DIR defaults to ../jtk_b200 (the in-tree build output).
fn main() {
    let dir = std::env::var("JTK_GPU_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{here}/../jtk_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=jtkgpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=JTK_GPU_LIB_DIR");
    println!("cargo:rerun-if-changed=../include/jtk_gpu.h");
}
