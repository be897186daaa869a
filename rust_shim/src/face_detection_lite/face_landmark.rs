//! face_detection_to_roi, FaceLandmark::new / infer (reference face_landmark.rs:180, :208, :232) over the C ABI.
use super::{ffi, render::{landmarks_to_render_data, Annotation, Color}, transform::SizeMode, types::{Detection, Landmark, Rect}, utils::{default_device, Frame}};
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

/// The face-mesh contours drawn between landmarks (reference face_landmark.rs:35-160; MediaPipe's face_landmarks_to_render_data_calculator).
pub const FACE_LANDMARK_CONNECTIONS: [(i32, i32); 124] = [
    (61, 146), (146, 91), (91, 181), (181, 84), (84, 17), (17, 314), (314, 405), (405, 321), (321, 375), (375, 291), (61, 185), (185, 40), (40, 39), (39, 37), (37, 0), (0, 267),
    (267, 269), (269, 270), (270, 409), (409, 291), (78, 95), (95, 88), (88, 178), (178, 87), (87, 14), (14, 317), (317, 402), (402, 318), (318, 324), (324, 308), (78, 191), (191, 80),
    (80, 81), (81, 82), (82, 13), (13, 312), (312, 311), (311, 310), (310, 415), (415, 308), (33, 7), (7, 163), (163, 144), (144, 145), (145, 153), (153, 154), (154, 155), (155, 133),
    (33, 246), (246, 161), (161, 160), (160, 159), (159, 158), (158, 157), (157, 173), (173, 133), (46, 53), (53, 52), (52, 65), (65, 55), (70, 63), (63, 105), (105, 66), (66, 107),
    (263, 249), (249, 390), (390, 373), (373, 374), (374, 380), (380, 381), (381, 382), (382, 362), (263, 466), (466, 388), (388, 387), (387, 386), (386, 385), (385, 384), (384, 398), (398, 362),
    (276, 283), (283, 282), (282, 295), (295, 285), (300, 293), (293, 334), (334, 296), (296, 336), (10, 338), (338, 297), (297, 332), (332, 284), (284, 251), (251, 389), (389, 356), (356, 454),
    (454, 323), (323, 361), (361, 288), (288, 397), (397, 365), (365, 379), (379, 378), (378, 400), (400, 377), (377, 152), (152, 148), (148, 176), (176, 149), (149, 150), (150, 136), (136, 172),
    (172, 58), (58, 132), (132, 93), (93, 234), (234, 127), (127, 162), (162, 21), (21, 54), (54, 103), (103, 67), (67, 109), (109, 10),
];

pub fn face_landmarks_to_render_data(
    face_landmarks: Vec<Landmark>, landmark_color: Color, connection_color: Color, thickness: Option<f32>, output: Option<Vec<Annotation>>,
) -> Vec<Annotation> {
    landmarks_to_render_data(face_landmarks, FACE_LANDMARK_CONNECTIONS.to_vec(), Some(landmark_color), Some(connection_color),
                             Some(thickness.unwrap_or(2.0)), Some(true), output)
}
