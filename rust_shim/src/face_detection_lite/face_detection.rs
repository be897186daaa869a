//! FaceDetection::new / infer (reference face_detection.rs:153, :205) over the C ABI.
use super::{ffi, types::{Detection, Rect}};
use anyhow::Error;
use opencv::core::Mat;
use std::ffi::CString;

pub enum FaceDetectionModel { FrontCamera = 0, BackCamera = 1, Short = 2, Full = 3, FullSparse = 4 }

pub struct FaceDetection { handle: *mut ffi::fdl_detector }
unsafe impl Send for FaceDetection {}

impl FaceDetection {
    pub fn new(model_type: FaceDetectionModel, model_path: Option<String>) -> Result<FaceDetection, Error> {
        let dir = model_path.map(|p| CString::new(p).unwrap());
        let mut h = std::ptr::null_mut();
        ffi::check(unsafe { ffi::fdl_detector_create(model_type as i32, dir.as_ref().map_or(std::ptr::null(), |c| c.as_ptr()), 0, &mut h) })?;
        Ok(FaceDetection { handle: h })
    }
    pub fn infer(&self, image: &Mat, roi: Option<Rect>) -> Result<Vec<Detection>, Error> {
        let img = ffi::image_of(image)?;
        let croi = roi.map(|r| r.to_c());
        let mut out = vec![ffi::fdl_detection { data: [0.0; 16], score: 0.0, anchor: -1 }; 128];
        let mut n = 0;
        ffi::check(unsafe { ffi::fdl_detector_infer(self.handle, &img, croi.as_ref().map_or(std::ptr::null(), |r| r as *const _), out.as_mut_ptr(), 128, &mut n) })?;
        Ok(out[..n as usize].iter().map(|d| Detection { data: d.data, score: d.score }).collect())
    }
}
impl Drop for FaceDetection { fn drop(&mut self) { unsafe { ffi::fdl_detector_destroy(self.handle) } } }
