// Links libcfft_b200.so.  Set CFFT_B200_LIB_DIR to the directory that holds it
// (concrete_fft_b200/ in this repository after `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    let dir = std::env::var("CFFT_B200_LIB_DIR")
        .unwrap_or_else(|_| concat!(env!("CARGO_MANIFEST_DIR"), "/../../concrete_fft_b200").to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=cfft_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=CFFT_B200_LIB_DIR");
}
