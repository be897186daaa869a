//! Raw bindings to include/fdl.h (the subset the shim needs).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

#[repr(C)] #[derive(Clone, Copy, Debug)]
pub struct fdl_rect { pub x_center: f64, pub y_center: f64, pub width: f64, pub height: f64, pub rotation: f64, pub normalized: i32, pub _pad: i32 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct fdl_detection { pub data: [f32; 16], pub score: f32, pub anchor: i32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct fdl_landmark { pub x: f64, pub y: f64, pub z: f64 }
#[repr(C)]
pub struct fdl_image { pub data: *const u8, pub width: i32, pub height: i32, pub row_stride: i64, pub mem: i32, pub _pad: i32 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct fdl_primitive { pub kind: i32, pub normalized: i32, pub a: f64, pub b: f64, pub c: f64, pub d: f64, pub thickness: f64,
                           pub r: u8, pub g: u8, pub b_: u8, pub alpha: u8, pub _pad: i32 }
pub const FDL_PRIM_POINT: i32 = 0; pub const FDL_PRIM_LINE: i32 = 1; pub const FDL_PRIM_RECT: i32 = 2; pub const FDL_PRIM_FILLED_RECT: i32 = 3;
pub const FDL_MEM_HOST: i32 = 0; pub const FDL_MEM_DEVICE: i32 = 1;
pub enum fdl_frame {}
pub enum fdl_detector {}
pub enum fdl_landmark_model {}
pub enum fdl_iris_model {}

extern "C" {
    pub fn fdl_last_error() -> *const c_char;
    pub fn fdl_detector_create(model: c_int, model_dir: *const c_char, device: c_int, out: *mut *mut fdl_detector) -> c_int;
    pub fn fdl_detector_destroy(d: *mut fdl_detector);
    pub fn fdl_detector_infer(d: *mut fdl_detector, image: *const fdl_image, roi: *const fdl_rect, out: *mut fdl_detection, cap: c_int, n: *mut c_int) -> c_int;
    pub fn fdl_landmark_create(file: *const c_char, device: c_int, out: *mut *mut fdl_landmark_model) -> c_int;
    pub fn fdl_landmark_destroy(m: *mut fdl_landmark_model);
    pub fn fdl_landmark_infer(m: *mut fdl_landmark_model, image: *const fdl_image, roi: *const fdl_rect, out: *mut fdl_landmark, n: *mut c_int, flag: *mut f32) -> c_int;
    pub fn fdl_iris_create(file: *const c_char, device: c_int, out: *mut *mut fdl_iris_model) -> c_int;
    pub fn fdl_iris_destroy(m: *mut fdl_iris_model);
    pub fn fdl_iris_infer(m: *mut fdl_iris_model, image: *const fdl_image, roi: *const fdl_rect, is_right_eye: c_int, contour: *mut fdl_landmark, iris: *mut fdl_landmark) -> c_int;
    pub fn fdl_face_detection_to_roi(device: c_int, det: *const fdl_detection, w: c_int, h: c_int, size_mode: c_int, out: *mut fdl_rect) -> c_int;
    pub fn fdl_iris_roi_from_face_landmarks(device: c_int, lm: *const fdl_landmark, n: c_int, w: c_int, h: c_int, left: *mut fdl_rect, right: *mut fdl_rect) -> c_int;
    pub fn fdl_update_face_landmarks_with_iris_results(device: c_int, face: *const fdl_landmark, n: c_int, left: *const fdl_landmark, n_left: c_int,
                                                       right: *const fdl_landmark, n_right: c_int, refined: *mut fdl_landmark) -> c_int;
    pub fn fdl_jpeg_info(data: *const u8, len: usize, width: *mut c_int, height: *mut c_int, components: *mut c_int) -> c_int;
    pub fn fdl_decode_jpeg(device: c_int, data: *const u8, len: usize, out_rgb: *mut u8, cap: usize, width: *mut c_int, height: *mut c_int) -> c_int;
    pub fn fdl_frame_create(device: c_int, out: *mut *mut fdl_frame) -> c_int;
    pub fn fdl_frame_destroy(f: *mut fdl_frame);
    pub fn fdl_frame_upload(f: *mut fdl_frame, image: *const fdl_image) -> c_int;
    pub fn fdl_frame_upload_jpeg(f: *mut fdl_frame, data: *const u8, len: usize) -> c_int;
    pub fn fdl_frame_image(f: *const fdl_frame, out: *mut fdl_image) -> c_int;
    pub fn fdl_iris_diameter(device: c_int, iris: *const fdl_landmark, n: c_int, w: c_int, h: c_int, out: *mut f64) -> c_int;
    pub fn fdl_render_to_image(device: c_int, image: *const fdl_image, primitives: *const fdl_primitive, n: c_int, out_rgba: *mut u8, cap: usize, out_mem: c_int) -> c_int;
    pub fn fdl_iris_depth(device: c_int, iris: *const fdl_landmark, n: c_int, focal_length_mm: f64, iris_size_px: f64, w: c_int, h: c_int, out: *mut f64) -> c_int;
}

pub const FDL_ERR_CAPACITY: c_int = -5;

pub fn check(rc: c_int) -> Result<(), anyhow::Error> {
    if rc == 0 { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(fdl_last_error()) }.to_string_lossy().into_owned();
    Err(anyhow::Error::msg(msg))
}

/// `&Mat` (8UC3, RGB) -> fdl_image without copying.
pub fn image_of(mat: &opencv::core::Mat) -> Result<fdl_image, anyhow::Error> {
    use opencv::prelude::*;
    let size = mat.size()?;
    Ok(fdl_image { data: mat.data(), width: size.width, height: size.height, row_stride: mat.step1(0)? as i64, mem: 0, _pad: 0 })
}
