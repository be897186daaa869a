fn main() {
    // libfdl_b200.so is built by `python -m rs_face_detection_tflite_b200.build` (nvcc, sm_100a)
    let dir = std::env::var("FDL_LIB_DIR").unwrap_or_else(|_| "../rs_face_detection_tflite_b200".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=fdl_b200");
}
