// Links the CUDA library built by `python -m petal_decomposition_b200.build`.
fn main() {
    let dir = std::env::var("PETAL_B200_LIB_DIR").unwrap_or_else(|_| "../petal_decomposition_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=petal_b200");
}
