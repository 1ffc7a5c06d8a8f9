// Links the in-tree shared library: SBN_LIB_DIR = <repo>/starky-bn254_b200 (where `make` leaves libstarkybn254_b200.so).
fn main() {
    let dir = std::env::var("SBN_LIB_DIR").expect("set SBN_LIB_DIR to the directory that holds libstarkybn254_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=starkybn254_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=SBN_LIB_DIR");
}
