// Links librpt_b200.so.  RPT_B200_LIB_DIR = the directory that holds it (rpt_b200/lib of the rpt-b200 checkout).
fn main() {
    println!("cargo:rerun-if-env-changed=RPT_B200_LIB_DIR");
    if let Ok(dir) = std::env::var("RPT_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=rpt_b200");
}
