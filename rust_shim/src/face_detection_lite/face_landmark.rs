//! face_detection_to_roi, FaceLandmark::new / infer (reference face_landmark.rs:180, :208, :232) over the C ABI.
use super::{ffi, types::{Detection, Landmark, Rect, SizeMode}};
use anyhow::Error;
use opencv::core::Mat;
use std::ffi::CString;

pub fn face_detection_to_roi(face_detection: Detection, image_size: (i32, i32), size_mode: Option<SizeMode>) -> Result<Rect, Error> {
    let det = ffi::fdl_detection { data: face_detection.data, score: face_detection.score, anchor: -1 };
    let mut out = Rect { x_center: 0.0, y_center: 0.0, width: 0.0, height: 0.0, rotation: 0.0, normalized: true }.to_c();
    ffi::check(unsafe { ffi::fdl_face_detection_to_roi(0, &det, image_size.0, image_size.1, size_mode.map_or(-1, |m| m as i32), &mut out) })?;
    Ok(Rect::from_c(&out))
}

pub struct FaceLandmark { handle: *mut ffi::fdl_landmark_model }
unsafe impl Send for FaceLandmark {}

impl FaceLandmark {
    pub fn new(model_path: Option<String>) -> Result<FaceLandmark, Error> {
        let file = model_path.map(|p| CString::new(p).unwrap());
        let mut h = std::ptr::null_mut();
        ffi::check(unsafe { ffi::fdl_landmark_create(file.as_ref().map_or(std::ptr::null(), |c| c.as_ptr()), 0, &mut h) })?;
        Ok(FaceLandmark { handle: h })
    }
    pub fn infer(&self, image: &Mat, roi: Option<Rect>) -> Result<Vec<Landmark>, Error> {
        let img = ffi::image_of(image)?;
        let croi = roi.map(|r| r.to_c());
        let mut out = vec![ffi::fdl_landmark::default(); 468];
        let (mut n, mut flag) = (0, 0f32);
        ffi::check(unsafe { ffi::fdl_landmark_infer(self.handle, &img, croi.as_ref().map_or(std::ptr::null(), |r| r as *const _), out.as_mut_ptr(), &mut n, &mut flag) })?;
        Ok(out[..n as usize].iter().map(|l| Landmark { x: l.x, y: l.y, z: l.z }).collect())
    }
}
impl Drop for FaceLandmark { fn drop(&mut self) { unsafe { ffi::fdl_landmark_destroy(self.handle) } } }
