//! FaceDetection::new / infer, FaceDetectionModel, FaceIndex (reference face_detection.rs:89-123, :153, :205) over the C ABI.
use super::{ffi, types::{Detection, Rect}, utils::{default_device, Frame}};
use anyhow::Error;
use opencv::core::Mat;
use std::ffi::CString;

/// Indexes of keypoints returned by the face detection model (face_detection.rs:89-114).
#[repr(i32)]
#[derive(Debug, Copy, Clone, PartialEq, Eq)]
pub enum FaceIndex { LeftEye = 0, RightEye = 1, NoseTip = 2, Mouth = 3, LeftEyeTragion = 4, RightEyeTragion = 5 }

impl TryFrom<i32> for FaceIndex {
    type Error = ();
    fn try_from(v: i32) -> Result<Self, Self::Error> {
        const ALL: [FaceIndex; 6] = [FaceIndex::LeftEye, FaceIndex::RightEye, FaceIndex::NoseTip, FaceIndex::Mouth, FaceIndex::LeftEyeTragion,
                                     FaceIndex::RightEyeTragion];
        ALL.iter().copied().find(|k| *k as i32 == v).ok_or(())
    }
}

/// face_detection.rs:117-123.
#[derive(Debug, Clone, PartialEq, Eq)]
pub enum FaceDetectionModel { FrontCamera = 0, BackCamera = 1, Short = 2, Full = 3, FullSparse = 4 }

pub struct FaceDetection { handle: *mut ffi::fdl_detector }
unsafe impl Send for FaceDetection {}

impl FaceDetection {
    /// `model_path` is the DIRECTORY that holds the .tflite files (None: "./models"), as in the reference.
    pub fn new(model_type: FaceDetectionModel, model_path: Option<String>) -> Result<FaceDetection, Error> {
        let dir = model_path.map(|p| CString::new(p).unwrap());
        let mut h = std::ptr::null_mut();
        ffi::check(unsafe { ffi::fdl_detector_create(model_type as i32, dir.as_ref().map_or(std::ptr::null(), |c| c.as_ptr()), default_device(), &mut h) })?;
        Ok(FaceDetection { handle: h })
    }
    pub fn infer(&self, image: &Mat, roi: Option<Rect>) -> Result<Vec<Detection>, Error> {
        self.infer_image(&ffi::image_of(image)?, roi)
    }
    /// The same on a frame that is already on the device.
    pub fn infer_frame(&self, frame: &Frame, roi: Option<Rect>) -> Result<Vec<Detection>, Error> {
        self.infer_image(&frame.image()?, roi)
    }
    fn infer_image(&self, img: &ffi::fdl_image, roi: Option<Rect>) -> Result<Vec<Detection>, Error> {
        let croi = roi.map(|r| r.to_c());
        let roi_ptr = croi.as_ref().map_or(std::ptr::null(), |r| r as *const _);
        // the reference returns an unbounded Vec: ask again with the reported count if the first buffer was too small
        let mut cap = 128usize;
        loop {
            let mut out = vec![ffi::fdl_detection { data: [0.0; 16], score: 0.0, anchor: -1 }; cap];
            let mut n = 0;
            let rc = unsafe { ffi::fdl_detector_infer(self.handle, img, roi_ptr, out.as_mut_ptr(), cap as i32, &mut n) };
            if rc == ffi::FDL_ERR_CAPACITY && (n as usize) > cap { cap = n as usize; continue; }
            ffi::check(rc)?;
            return Ok(out[..n as usize].iter().map(Detection::from_c).collect());
        }
    }
}
impl Drop for FaceDetection { fn drop(&mut self) { unsafe { ffi::fdl_detector_destroy(self.handle) } } }
