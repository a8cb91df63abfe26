fn main() {
    // LABRADOR_B200_LIB_DIR = directory holding liblabrador_b200.so (built by `python __graft_entry__.py`)
    let dir = std::env::var("LABRADOR_B200_LIB_DIR").unwrap_or_else(|_| "../..".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=labrador_b200");
}
