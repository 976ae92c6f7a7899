// Links against the in-tree C-ABI library built by recast_rs_b200/build.py.
fn main() {
    let dir = std::env::var("RECAST_B200_LIB_DIR").unwrap_or_else(|_| "../../recast_rs_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=recast_b200");
}
