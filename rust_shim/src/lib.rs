//! Drop-in replacement for `rs_face_detection_tfite::face_detection_lite` on B200.
//! Same public names and signatures as the reference (face_detection.rs:117-267, face_landmark.rs:180-306,
//! iris_landmark.rs:115-292, types.rs); the bodies forward to the C ABI of include/fdl.h.
pub mod face_detection_lite {
    pub mod ffi;
    pub mod types;
    pub mod transform;
    pub mod utils;
    pub mod face_detection;
    pub mod face_landmark;
    pub mod iris_landmark;
    pub mod render;
}
