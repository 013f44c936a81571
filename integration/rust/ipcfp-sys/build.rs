// Links the prebuilt C-ABI library (make at the repo root produces ipc_filecoin_proofs_b200/libipcfp.so).
fn main() {
    let dir = std::env::var("IPCFP_LIB_DIR").unwrap_or_else(|_| "../../../ipc_filecoin_proofs_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ipcfp");
    println!("cargo:rerun-if-env-changed=IPCFP_LIB_DIR");
}
