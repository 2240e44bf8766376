// Links libzg_b200.so (built by `make -C 0g-halo2_b200/csrc`); ZG_B200_LIB_DIR = directory holding it.
fn main() {
    let dir = std::env::var("ZG_B200_LIB_DIR").expect("set ZG_B200_LIB_DIR to the directory of libzg_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=zg_b200");
    println!("cargo:rerun-if-env-changed=ZG_B200_LIB_DIR");
}
