// Tell cargo where libceno_b200.so lives: CENO_B200_LIB_DIR=<repo>/ceno_b200/lib
fn main() {
    if let Ok(dir) = std::env::var("CENO_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=ceno_b200");
    println!("cargo:rerun-if-env-changed=CENO_B200_LIB_DIR");
}
