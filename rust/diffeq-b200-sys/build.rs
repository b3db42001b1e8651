fn main() {
    // DIFFEQ_B200_LIB_DIR must point at the directory holding libdiffeq_b200.so (repo: diffeq_b200/).
    if let Ok(dir) = std::env::var("DIFFEQ_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=diffeq_b200");
}
