//! The public enum of the reference's transform.rs (:15-40).  The functions of that file (image_to_tensor, project_landmarks,
//! bbox_to_roi, ...) run on the device behind FaceDetection / FaceLandmark / IrisLandmark and the two ROI free functions.
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum SizeMode {
    Default = 0,
    /// Make square using `max(width, height)`.
    SquareLong = 1,
    /// Make square using `min(width, height)`.
    SquareShort = 2,
}

impl From<i32> for SizeMode {
    fn from(value: i32) -> Self {
        match value {
            1 => SizeMode::SquareLong,
            2 => SizeMode::SquareShort,
            _ => SizeMode::Default,
        }
    }
}

impl SizeMode {
    pub fn to_int(self) -> i32 {
        self as i32
    }
}
