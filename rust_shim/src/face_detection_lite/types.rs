//! The value types of the reference's public API (types.rs:5-246), with every public method it has, so that user code written
//! against the reference compiles unchanged.  Pure host arithmetic (a few flops per call): nothing here is on the hot path; the
//! `to_c` / `from_c` conversions feed the C ABI of include/fdl.h.
use super::ffi;
use ndarray::{Array2, ArrayD};

/// types.rs:5-22.  The shim's `infer` methods never build one (image_to_tensor runs on the device); kept for signature parity.
#[derive(Debug, Clone)]
pub struct ImageTensor {
    pub tensor_data: ArrayD<f32>,
    pub padding: (f64, f64, f64, f64),
    pub original_size: (i32, i32),
}

impl ImageTensor {
    pub fn new(tensor_data: ArrayD<f32>, padding: (f64, f64, f64, f64), original_size: (i32, i32)) -> Self {
        ImageTensor { tensor_data, padding, original_size }
    }
}

/// types.rs:24-97: centre / size / rotation (radians, clockwise); `normalized` = relative to the image size.
#[derive(Debug, Clone, Copy)]
pub struct Rect {
    pub x_center: f64,
    pub y_center: f64,
    pub width: f64,
    pub height: f64,
    pub rotation: f64,
    pub normalized: bool,
}

impl Rect {
    pub fn new(x_center: f64, y_center: f64, width: f64, height: f64, rotation: f64, normalized: bool) -> Self {
        Rect { x_center, y_center, width, height, rotation, normalized }
    }

    /// (width, height); truncated to whole pixels for an absolute rectangle (types.rs:52-59).
    pub fn size(&self) -> (f64, f64) {
        if self.normalized {
            (self.width, self.height)
        } else {
            (self.width as i32 as f64, self.height as i32 as f64)
        }
    }

    /// Normalised <-> absolute for an image of `size` (types.rs:62-77); the rotation is carried over as is.
    pub fn scaled(&self, size: (f64, f64), normalize: bool) -> Rect {
        if normalize == self.normalized {
            return *self;
        }
        let (fx, fy) = if normalize { (size.0.recip(), size.1.recip()) } else { size };
        Rect::new(self.x_center * fx, self.y_center * fy, self.width * fx, self.height * fy, self.rotation, normalize)
    }

    /// Corners TL, TR, BR, BL, rotated about the centre (types.rs:80-96).
    pub fn points(&self) -> Vec<(f64, f64)> {
        let (hw, hh) = (self.width / 2.0, self.height / 2.0);
        let (cx, cy) = (self.x_center, self.y_center);
        let corners = [(cx - hw, cy - hh), (cx + hw, cy - hh), (cx + hw, cy + hh), (cx - hw, cy + hh)];
        if self.rotation == 0.0 {
            return corners.to_vec();
        }
        let (sin, cos) = (self.rotation.sin(), self.rotation.cos());
        corners
            .iter()
            .map(|&(px, py)| {
                let (dx, dy) = (px - cx, py - cy);
                (cx + dx * cos - dy * sin, cy + dx * sin + dy * cos)
            })
            .collect()
    }

    pub(crate) fn to_c(&self) -> ffi::fdl_rect {
        ffi::fdl_rect { x_center: self.x_center, y_center: self.y_center, width: self.width, height: self.height, rotation: self.rotation,
                        normalized: self.normalized as i32, _pad: 0 }
    }
    pub(crate) fn from_c(c: &ffi::fdl_rect) -> Rect {
        Rect::new(c.x_center, c.y_center, c.width, c.height, c.rotation, c.normalized != 0)
    }
}

/// types.rs:99-174.
#[derive(Debug, Clone, Copy)]
pub struct BBox {
    pub xmin: f64,
    pub ymin: f64,
    pub xmax: f64,
    pub ymax: f64,
}

impl BBox {
    pub fn new(xmin: f64, ymin: f64, xmax: f64, ymax: f64) -> Self {
        BBox { xmin, ymin, xmax, ymax }
    }
    pub fn as_tuple(&self) -> (f64, f64, f64, f64) {
        (self.xmin, self.ymin, self.xmax, self.ymax)
    }
    pub fn width(&self) -> f64 {
        self.xmax - self.xmin
    }
    pub fn height(&self) -> f64 {
        self.ymax - self.ymin
    }
    pub fn empty(&self) -> bool {
        self.width() <= 0.0 || self.height() <= 0.0
    }
    /// The reference's (odd) test, kept as it is (types.rs:134-136): xmin >= -1, xmax < 2, ymin >= -1.
    pub fn normalized(&self) -> bool {
        self.xmin >= -1.0 && self.xmax < 2.0 && self.ymin >= -1.0
    }
    pub fn area(&self) -> f64 {
        if self.empty() { 0.0 } else { self.width() * self.height() }
    }
    pub fn intersect(&self, other: &BBox) -> Option<BBox> {
        let b = BBox::new(self.xmin.max(other.xmin), self.ymin.max(other.ymin), self.xmax.min(other.xmax), self.ymax.min(other.ymax));
        if b.xmin < b.xmax && b.ymin < b.ymax { Some(b) } else { None }
    }
    pub fn scale(&self, size: (f64, f64)) -> BBox {
        BBox::new(self.xmin * size.0, self.ymin * size.1, self.xmax * size.0, self.ymax * size.1)
    }
    pub fn absolute(&self, size: (i32, i32)) -> BBox {
        if self.normalized() { self.scale((size.0 as f64, size.1 as f64)) } else { *self }
    }
}

/// types.rs:176-187.
#[derive(Debug, Clone, Copy)]
pub struct Landmark {
    pub x: f64,
    pub y: f64,
    pub z: f64,
}

impl Landmark {
    pub fn new(x: f64, y: f64, z: f64) -> Self {
        Landmark { x, y, z }
    }
    pub(crate) fn to_c(&self) -> ffi::fdl_landmark {
        ffi::fdl_landmark { x: self.x, y: self.y, z: self.z }
    }
    pub(crate) fn from_c(c: &ffi::fdl_landmark) -> Landmark {
        Landmark::new(c.x, c.y, c.z)
    }
}

/// types.rs:189-246: `data` is [rows, 2]: row 0 = (xmin, ymin), row 1 = (xmax, ymax), rows 2.. = keypoints.
#[derive(Debug, Clone)]
pub struct Detection {
    pub data: Array2<f32>,
    pub score: f32,
}

impl Detection {
    /// Panics like the reference when fewer than four values are given (types.rs:197).
    pub fn new(data: Vec<f32>, score: f32) -> Self {
        assert!(data.len() >= 4, "Data must contain at least four elements for the bounding box");
        let rows = data.len() / 2;
        let data = Array2::from_shape_vec((rows, 2), data).expect("an even number of values");
        Detection { data, score }
    }
    pub fn keypoint_count(&self) -> usize {
        self.data.nrows() - 2
    }
    pub fn keypoint(&self, key: usize) -> (f32, f32) {
        (self.data[[key + 2, 0]], self.data[[key + 2, 1]])
    }
    pub fn bbox(&self) -> BBox {
        BBox::new(self.data[[0, 0]] as f64, self.data[[0, 1]] as f64, self.data[[1, 0]] as f64, self.data[[1, 1]] as f64)
    }
    pub fn scaled(&self, factor: f32) -> Detection {
        Detection { data: &self.data * factor, score: self.score }
    }
    /// x by the width, y by the height, in f32 as the reference multiplies (types.rs:237-245).
    pub fn scaled_by_image_size(&self, image_size: (i32, i32)) -> Detection {
        let (sx, sy) = (image_size.0 as f32, image_size.1 as f32);
        let mut data = self.data.clone();
        for mut row in data.rows_mut() {
            row[0] *= sx;
            row[1] *= sy;
        }
        Detection { data, score: self.score }
    }

    /// The C record of the ABI holds the 8 x 2 layout of the face detectors.
    pub(crate) fn to_c(&self) -> Result<ffi::fdl_detection, anyhow::Error> {
        if self.data.nrows() != 8 {
            return Err(anyhow::Error::msg("a face detection has 8 rows (box + 6 keypoints)"));
        }
        let mut data = [0f32; 16];
        for (i, v) in self.data.iter().enumerate() {
            data[i] = *v;
        }
        Ok(ffi::fdl_detection { data, score: self.score, anchor: -1 })
    }
    pub(crate) fn from_c(c: &ffi::fdl_detection) -> Detection {
        Detection::new(c.data.to_vec(), c.score)
    }
}
