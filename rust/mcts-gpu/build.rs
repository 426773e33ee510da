// Links libmcts_b200.so (built by `python -m mcts_b200.build` with nvcc for sm_100a).
// MCTS_B200_LIB_DIR = the directory holding the library (mcts_b200/_build in this repository).
fn main() {
    let dir = std::env::var("MCTS_B200_LIB_DIR").expect("set MCTS_B200_LIB_DIR to the directory of libmcts_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=mcts_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=MCTS_B200_LIB_DIR");
}
