//! iris_roi_from_face_landmarks, IrisLandmark::new / infer, IrisResults, update_face_landmarks_with_iris_results and the iris
//! diameter / depth helpers (reference iris_landmark.rs:115-433) over the C ABI.
use super::{ffi, render::{landmarks_to_render_data, Annotation, AnnotationData, Color, Point, RectOrOval}, types::{Landmark, Rect}, utils::{default_device, Frame}};
use anyhow::Error;
use opencv::core::Mat;
use std::ffi::CString;

/// iris_landmark.rs:104-110.
#[repr(i32)]
#[derive(Debug, Copy, Clone, PartialEq, Eq)]
pub enum IrisIndex { Center = 0, Left = 1, Top = 2, Right = 3, Bottom = 4 }

/// 71 contour points of the eye region + 5 iris points (iris_landmark.rs:115-129).
pub struct IrisResults { contour: Vec<Landmark>, iris: Vec<Landmark> }
impl IrisResults {
    pub fn new(contour: Vec<Landmark>, iris: Vec<Landmark>) -> Self { IrisResults { contour, iris } }
    pub fn eyeball_contour(&self) -> Vec<Landmark> { self.contour[..15].to_vec() }
    pub fn contour(&self) -> &[Landmark] { &self.contour }
    pub fn iris(&self) -> &[Landmark] { &self.iris }
}

pub fn iris_roi_from_face_landmarks(face_landmarks: Vec<Landmark>, image_size: (i32, i32)) -> Result<(Rect, Rect), Error> {
    let lm = to_c(&face_landmarks);
    let zero = Rect::new(0.0, 0.0, 0.0, 0.0, 0.0, true).to_c();
    let (mut l, mut r) = (zero, zero);
    ffi::check(unsafe { ffi::fdl_iris_roi_from_face_landmarks(default_device(), lm.as_ptr(), lm.len() as i32, image_size.0, image_size.1, &mut l, &mut r) })?;
    Ok((Rect::from_c(&l), Rect::from_c(&r)))
}

pub struct IrisLandmark { handle: *mut ffi::fdl_iris_model }
unsafe impl Send for IrisLandmark {}

impl IrisLandmark {
    pub fn new(model_path: Option<String>) -> Result<IrisLandmark, Error> {
        let file = model_path.map(|p| CString::new(p).unwrap());
        let mut h = std::ptr::null_mut();
        ffi::check(unsafe { ffi::fdl_iris_create(file.as_ref().map_or(std::ptr::null(), |c| c.as_ptr()), default_device(), &mut h) })?;
        Ok(IrisLandmark { handle: h })
    }
    pub fn infer(&self, image: &Mat, roi: Option<Rect>, is_right_eye: Option<bool>) -> Result<IrisResults, Error> {
        self.infer_image(&ffi::image_of(image)?, roi, is_right_eye)
    }
    pub fn infer_frame(&self, frame: &Frame, roi: Option<Rect>, is_right_eye: Option<bool>) -> Result<IrisResults, Error> {
        self.infer_image(&frame.image()?, roi, is_right_eye)
    }
    fn infer_image(&self, img: &ffi::fdl_image, roi: Option<Rect>, is_right_eye: Option<bool>) -> Result<IrisResults, Error> {
        let croi = roi.map(|r| r.to_c());
        let mut contour = vec![ffi::fdl_landmark::default(); 71];
        let mut iris = vec![ffi::fdl_landmark::default(); 5];
        ffi::check(unsafe { ffi::fdl_iris_infer(self.handle, img, croi.as_ref().map_or(std::ptr::null(), |r| r as *const _), is_right_eye.unwrap_or(false) as i32,
                                                 contour.as_mut_ptr(), iris.as_mut_ptr()) })?;
        let f = |v: &Vec<ffi::fdl_landmark>| v.iter().map(Landmark::from_c).collect();
        Ok(IrisResults::new(f(&contour), f(&iris)))
    }
}
impl Drop for IrisLandmark { fn drop(&mut self) { unsafe { ffi::fdl_iris_destroy(self.handle) } } }

fn to_c(v: &[Landmark]) -> Vec<ffi::fdl_landmark> { v.iter().map(|l| l.to_c()).collect() }

/// Update face landmarks with iris detection results (reference iris_landmark.rs:380-398).
pub fn update_face_landmarks_with_iris_results(
    face_landmarks: Vec<Landmark>, iris_data_left: IrisResults, iris_data_right: IrisResults,
) -> Result<Vec<Landmark>, Error> {
    let (face, left, right) = (to_c(&face_landmarks), to_c(&iris_data_left.contour), to_c(&iris_data_right.contour));
    let mut out = vec![ffi::fdl_landmark::default(); 468];
    ffi::check(unsafe {
        ffi::fdl_update_face_landmarks_with_iris_results(default_device(), face.as_ptr(), face.len() as i32, left.as_ptr(), left.len() as i32, right.as_ptr(),
                                                         right.len() as i32, out.as_mut_ptr())
    })?;
    Ok(out.iter().map(Landmark::from_c).collect())
}

/// Iris diameter in pixels (reference iris_landmark.rs:401-418; private there).
pub fn get_iris_diameter(iris_landmarks: &Vec<Landmark>, image_size: (i32, i32)) -> Result<f64, Error> {
    let (iris, mut d) = (to_c(iris_landmarks), 0.0f64);
    ffi::check(unsafe { ffi::fdl_iris_diameter(default_device(), iris.as_ptr(), iris.len() as i32, image_size.0, image_size.1, &mut d) })?;
    Ok(d)
}

/// Iris depth in mm from the lens focal length in mm (reference iris_landmark.rs:421-433; private there).
pub fn get_iris_depth(iris_landmarks: Vec<Landmark>, focal_length_mm: f64, iris_size_px: f64, image_size: (i32, i32)) -> Result<f64, Error> {
    let (iris, mut d) = (to_c(&iris_landmarks), 0.0f64);
    ffi::check(unsafe { ffi::fdl_iris_depth(default_device(), iris.as_ptr(), iris.len() as i32, focal_length_mm, iris_size_px, image_size.0, image_size.1, &mut d) })?;
    Ok(d)
}

/// Eye contour connections and the number of contour points drawn (reference iris_landmark.rs:44-62).
pub const EYE_LANDMARK_CONNECTIONS: [(i32, i32); 15] = [
    (0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6), (6, 7), (7, 8), (9, 10), (10, 11), (11, 12), (12, 13), (13, 14), (0, 9), (8, 14),
];
pub const MAX_EYE_LANDMARK: usize = EYE_LANDMARK_CONNECTIONS.len();

pub fn eye_landmarks_to_render_data(
    eye_contour: Vec<Landmark>, landmark_color: Color, connection_color: Color, thickness: Option<f32>, output: Option<Vec<Annotation>>,
) -> Vec<Annotation> {
    landmarks_to_render_data(eye_contour[0..MAX_EYE_LANDMARK].to_vec(), EYE_LANDMARK_CONNECTIONS.to_vec(), Some(landmark_color),
                             Some(connection_color), Some(thickness.unwrap_or(2.0)), Some(true), output)
}

pub fn iris_landmarks_to_render_data(
    iris_landmarks: Vec<Landmark>, landmark_color: Option<Color>, oval_color: Option<Color>, thickness: Option<f64>,
    image_size: Option<(i32, i32)>, output: Option<Vec<Annotation>>,
) -> Result<Vec<Annotation>, Error> {
    let (width, height) = image_size.unwrap_or((-1, -1));
    let thickness = thickness.unwrap_or(1.0);
    let mut out = output.unwrap_or_default();
    if let Some(c) = oval_color {
        if width < 2 || height < 2 { return Err(Error::msg("oval_color requires a valid image_size arg")); }
        let radius = get_iris_diameter(&iris_landmarks, (width, height))? / 2.0;
        let (rh, rv) = (radius / width as f64, radius / height as f64);
        let ctr = iris_landmarks[IrisIndex::Center as usize];
        out.push(Annotation::new(vec![AnnotationData::RectOrOval(RectOrOval::new(ctr.x - rh, ctr.y - rv, ctr.x + rh, ctr.y + rv, true))], true, thickness, c));
    }
    if let Some(c) = landmark_color {
        out.push(Annotation::new(iris_landmarks.iter().map(|l| AnnotationData::Point(Point::new(l.x, l.y))).collect(), true, thickness, c));
    }
    Ok(out)
}
