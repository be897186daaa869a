//! types.rs of the reference, reduced to the fields the hot path exchanges (types.rs:24-37, :99-106, :176-187, :189-207).
use super::ffi;

#[derive(Debug, Clone, Copy)]
pub struct Rect { pub x_center: f64, pub y_center: f64, pub width: f64, pub height: f64, pub rotation: f64, pub normalized: bool }
impl Rect {
    pub(crate) fn to_c(&self) -> ffi::fdl_rect {
        ffi::fdl_rect { x_center: self.x_center, y_center: self.y_center, width: self.width, height: self.height, rotation: self.rotation,
                        normalized: self.normalized as i32, _pad: 0 }
    }
    pub(crate) fn from_c(c: &ffi::fdl_rect) -> Rect {
        Rect { x_center: c.x_center, y_center: c.y_center, width: c.width, height: c.height, rotation: c.rotation, normalized: c.normalized != 0 }
    }
}
#[derive(Debug, Clone, Copy)]
pub struct BBox { pub xmin: f64, pub ymin: f64, pub xmax: f64, pub ymax: f64 }
#[derive(Debug, Clone, Copy)]
pub struct Landmark { pub x: f64, pub y: f64, pub z: f64 }
/// `data` is the reference's Array2<f32>[8,2] flattened row-major.
#[derive(Debug, Clone)]
pub struct Detection { pub data: [f32; 16], pub score: f32 }
impl Detection {
    pub fn bbox(&self) -> BBox { BBox { xmin: self.data[0] as f64, ymin: self.data[1] as f64, xmax: self.data[2] as f64, ymax: self.data[3] as f64 } }
    pub fn keypoint(&self, k: usize) -> (f32, f32) { (self.data[2 * (k + 2)], self.data[2 * (k + 2) + 1]) }
}
pub enum SizeMode { Default = 0, SquareLong = 1, SquareShort = 2 }
