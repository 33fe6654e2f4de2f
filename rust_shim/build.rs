// Links libkolibrie_b200.so. KOLIBRIE_B200_LIB_DIR points at the directory that holds it (kolibrie_b200/ of the kolibrie_b200 repository
// after `python -c 'import __graft_entry__ as g; g.build()'`); the same directory also holds the drop-in libcudajoin.so that Kolibrie's own
// build.rs links for the legacy `cuda` feature (kolibrie/build.rs:75-79).
fn main() {
    let dir = std::env::var("KOLIBRIE_B200_LIB_DIR").unwrap_or_else(|_| "/usr/local/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=kolibrie_b200");
    println!("cargo:rerun-if-env-changed=KOLIBRIE_B200_LIB_DIR");
}
