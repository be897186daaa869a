//! face_detection_to_roi, FaceLandmark::new / infer (reference face_landmark.rs:180, :208, :232) over the C ABI.
use super::{ffi, transform::SizeMode, types::{Detection, Landmark, Rect}, utils::{default_device, Frame}};
use anyhow::Error;
use opencv::core::Mat;
use std::ffi::CString;

pub fn face_detection_to_roi(face_detection: Detection, image_size: (i32, i32), size_mode: Option<SizeMode>) -> Result<Rect, Error> {
    let det = face_detection.to_c()?;
    let mut out = Rect::new(0.0, 0.0, 0.0, 0.0, 0.0, true).to_c();
    ffi::check(unsafe { ffi::fdl_face_detection_to_roi(default_device(), &det, image_size.0, image_size.1, size_mode.map_or(-1, |m| m.to_int()), &mut out) })?;
    Ok(Rect::from_c(&out))
}

pub struct FaceLandmark { handle: *mut ffi::fdl_landmark_model }
unsafe impl Send for FaceLandmark {}

impl FaceLandmark {
    /// `model_path` is the .tflite FILE (None: "./models/face_landmark.tflite"), as in the reference.
    pub fn new(model_path: Option<String>) -> Result<FaceLandmark, Error> {
        let file = model_path.map(|p| CString::new(p).unwrap());
        let mut h = std::ptr::null_mut();
        ffi::check(unsafe { ffi::fdl_landmark_create(file.as_ref().map_or(std::ptr::null(), |c| c.as_ptr()), default_device(), &mut h) })?;
        Ok(FaceLandmark { handle: h })
    }
    pub fn infer(&self, image: &Mat, roi: Option<Rect>) -> Result<Vec<Landmark>, Error> {
        self.infer_image(&ffi::image_of(image)?, roi)
    }
    pub fn infer_frame(&self, frame: &Frame, roi: Option<Rect>) -> Result<Vec<Landmark>, Error> {
        self.infer_image(&frame.image()?, roi)
    }
    fn infer_image(&self, img: &ffi::fdl_image, roi: Option<Rect>) -> Result<Vec<Landmark>, Error> {
        let croi = roi.map(|r| r.to_c());
        let mut out = vec![ffi::fdl_landmark::default(); 468];
        let (mut n, mut flag) = (0, 0f32);
        ffi::check(unsafe { ffi::fdl_landmark_infer(self.handle, img, croi.as_ref().map_or(std::ptr::null(), |r| r as *const _), out.as_mut_ptr(), &mut n, &mut flag) })?;
        Ok(out[..n as usize].iter().map(Landmark::from_c).collect())      // empty when the face flag says "no face" (face_landmark.rs:292-296)
    }
}
impl Drop for FaceLandmark { fn drop(&mut self) { unsafe { ffi::fdl_landmark_destroy(self.handle) } } }
