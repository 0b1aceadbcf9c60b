// Link against the in-tree libvoxb200.so built by csrc/Makefile.
// VOXB200_LIB_DIR overrides the search directory.
fn main() {
    let dir = std::env::var("VOXB200_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../..", manifest)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=voxb200");
    println!("cargo:rerun-if-env-changed=VOXB200_LIB_DIR");
}
