// Links the B200 path-tracing library. RT_B200_LIB_DIR = directory holding librt_b200.so.
fn main() {
    let dir = std::env::var("RT_B200_LIB_DIR").expect("set RT_B200_LIB_DIR to the directory of librt_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rt_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
}
