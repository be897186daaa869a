//! Drawing (reference render.rs:7-479, face_landmark.rs:324, iris_landmark.rs:312/:330): the same public types and functions;
//! `render_to_image` hands a flat primitive list to `fdl_render_to_image`, which paints it on the device with imageproc's rules
//! (Bresenham lines, 1-px hollow rectangles, clipped filled rectangles, later primitives over earlier ones, no blending).
use super::{ffi, types::{Detection, Landmark}, utils::default_device};
use anyhow::Error;
use image::{DynamicImage, GenericImageView, RgbaImage};

#[derive(Debug, Clone, Copy)]
pub struct Color { pub r: i32, pub g: i32, pub b: i32, pub a: Option<i32> }
impl Color {
    pub fn new(r: Option<i32>, g: Option<i32>, b: Option<i32>, a: Option<i32>) -> Self { Self { r: r.unwrap_or(0), g: g.unwrap_or(0), b: b.unwrap_or(0), a } }
    pub fn as_tuple(&self) -> (i32, i32, i32, Option<i32>) { (self.r, self.g, self.b, self.a) }
}
pub struct Colors;
impl Colors {
    pub const BLACK: Color = Color { r: 0, g: 0, b: 0, a: None };
    pub const RED: Color = Color { r: 255, g: 0, b: 0, a: None };
    pub const GREEN: Color = Color { r: 0, g: 255, b: 0, a: None };
    pub const BLUE: Color = Color { r: 0, g: 0, b: 255, a: None };
    pub const PINK: Color = Color { r: 255, g: 0, b: 255, a: None };
    pub const WHITE: Color = Color { r: 255, g: 255, b: 255, a: None };
}

#[derive(Debug, Clone, Copy)]
pub struct Point { pub x: f64, pub y: f64 }
impl Point {
    pub fn new(x: f64, y: f64) -> Self { Self { x, y } }
    pub fn as_tuple(&self) -> (f64, f64) { (self.x, self.y) }
    pub fn scaled(&self, factor: (f64, f64)) -> Self { Point { x: self.x * factor.0, y: self.y * factor.1 } }
}
#[derive(Debug, Clone, Copy)]
pub struct RectOrOval { pub left: f64, pub top: f64, pub right: f64, pub bottom: f64, pub oval: bool }
impl RectOrOval {
    pub fn new(left: f64, top: f64, right: f64, bottom: f64, oval: bool) -> Self { Self { left, top, right, bottom, oval } }
    pub fn as_tuple(&self) -> (f64, f64, f64, f64) { (self.left, self.top, self.right, self.bottom) }
    pub fn scaled(&self, f: (f64, f64)) -> Self { RectOrOval { left: self.left * f.0, top: self.top * f.1, right: self.right * f.0, bottom: self.bottom * f.1, oval: self.oval } }
}
#[derive(Debug, Clone, Copy)]
pub struct FilledRectOrOval { pub rect: RectOrOval, pub fill: Color }
impl FilledRectOrOval {
    pub fn new(rect: RectOrOval, fill: Color) -> Self { Self { rect, fill } }
    pub fn scaled(&self, factor: (f64, f64)) -> Self { FilledRectOrOval { rect: self.rect.scaled(factor), fill: self.fill } }
}
#[derive(Debug, Clone, Copy)]
pub struct Line { x_start: f64, y_start: f64, x_end: f64, y_end: f64, dashed: bool }
impl Line {
    pub fn new(x_start: f64, y_start: f64, x_end: f64, y_end: f64, dashed: bool) -> Self { Self { x_start, y_start, x_end, y_end, dashed } }
    pub fn as_tuple(&self) -> (f64, f64, f64, f64) { (self.x_start, self.y_start, self.x_end, self.y_end) }
    pub fn scaled(&self, f: (f64, f64)) -> Self { Line { x_start: self.x_start * f.0, y_start: self.y_start * f.1, x_end: self.x_end * f.0, y_end: self.y_end * f.1, dashed: self.dashed } }
}
#[derive(Debug, Clone, Copy)]
pub enum AnnotationData { Point(Point), RectOrOval(RectOrOval), FilledRectOrOval(FilledRectOrOval), Line(Line) }
impl AnnotationData {
    pub fn scaled(&self, factor: (f64, f64)) -> Self {
        match self {
            AnnotationData::Point(p) => AnnotationData::Point(p.scaled(factor)),
            AnnotationData::RectOrOval(r) => AnnotationData::RectOrOval(r.scaled(factor)),
            AnnotationData::FilledRectOrOval(r) => AnnotationData::FilledRectOrOval(r.scaled(factor)),
            AnnotationData::Line(l) => AnnotationData::Line(l.scaled(factor)),
        }
    }
}
#[derive(Debug, Clone)]
pub struct Annotation { data: Vec<AnnotationData>, normalized_positions: bool, thickness: f64, color: Color }
impl Annotation {
    pub fn new(data: Vec<AnnotationData>, normalized_positions: bool, thickness: f64, color: Color) -> Self { Self { data, normalized_positions, thickness, color } }
    pub fn scaled(&self, factor: (f64, f64)) -> Result<Self, Error> {
        if !self.normalized_positions { return Err(Error::msg("position data must be normalized")); }
        Ok(Annotation { data: self.data.iter().map(|d| d.scaled(factor)).collect(), normalized_positions: false, thickness: self.thickness, color: self.color })
    }
}

fn merged(output: Option<Vec<Annotation>>, mut new: Vec<Annotation>) -> Vec<Annotation> {
    match output { Some(mut o) => { o.append(&mut new); o } None => new }
}

pub fn detections_to_render_data(
    detections: Vec<Detection>, bounds_color: Option<Color>, keypoint_color: Option<Color>, line_width: i32, point_width: i32,
    normalized_positions: bool, output: Option<Vec<Annotation>>,
) -> Vec<Annotation> {
    let mut new = Vec::new();
    if let (Some(c), true) = (bounds_color, line_width > 0) {
        let rects = detections.iter().map(|d| { let b = d.bbox(); AnnotationData::RectOrOval(RectOrOval::new(b.xmin, b.ymin, b.xmax, b.ymax, false)) }).collect();
        new.push(Annotation::new(rects, normalized_positions, line_width as f64, c));
    }
    if let (Some(c), true) = (keypoint_color, point_width > 0) {
        let pts = detections.iter().flat_map(|d| d.data.rows().into_iter().map(|r| AnnotationData::Point(Point::new(r[0] as f64, r[1] as f64))).collect::<Vec<_>>()).collect();
        new.push(Annotation::new(pts, normalized_positions, point_width as f64, c));
    }
    merged(output, new)
}

pub fn landmarks_to_render_data(
    landmarks: Vec<Landmark>, landmark_connections: Vec<(i32, i32)>, landmark_color: Option<Color>, connection_color: Option<Color>,
    thickness: Option<f32>, normalized_positions: Option<bool>, output: Option<Vec<Annotation>>,
) -> Vec<Annotation> {
    let (lc, cc) = (landmark_color.unwrap_or(Colors::RED), connection_color.unwrap_or(Colors::RED));
    let (th, norm) = (thickness.unwrap_or(1.0) as f64, normalized_positions.unwrap_or(true));
    let lines = landmark_connections.iter().map(|&(a, b)| {
        let (p, q) = (&landmarks[a as usize], &landmarks[b as usize]);
        AnnotationData::Line(Line::new(p.x, p.y, q.x, q.y, false))
    }).collect();
    let points = landmarks.iter().map(|l| AnnotationData::Point(Point::new(l.x, l.y))).collect();
    merged(output, vec![Annotation::new(lines, norm, th, cc), Annotation::new(points, norm, th, lc)])
}

pub fn render_to_image(annotations: &Vec<Annotation>, image: &DynamicImage, _blend_mode: Option<bool>) -> DynamicImage {
    let (width, height) = image.dimensions();
    let rgb = image.to_rgb8();
    let mut prims = Vec::new();
    for a in annotations {
        for d in &a.data {
            let (kind, v, col) = match d {
                AnnotationData::Point(p) => (ffi::FDL_PRIM_POINT, [p.x, p.y, 0.0, 0.0], a.color),
                AnnotationData::Line(l) => (ffi::FDL_PRIM_LINE, [l.x_start, l.y_start, l.x_end, l.y_end], a.color),
                AnnotationData::RectOrOval(r) => (ffi::FDL_PRIM_RECT, [r.left, r.top, r.right, r.bottom], a.color),
                AnnotationData::FilledRectOrOval(f) => (ffi::FDL_PRIM_FILLED_RECT, [f.rect.left, f.rect.top, f.rect.right, f.rect.bottom], f.fill),
            };
            prims.push(ffi::fdl_primitive { kind, normalized: a.normalized_positions as i32, a: v[0], b: v[1], c: v[2], d: v[3], thickness: a.thickness,
                                            r: col.r as u8, g: col.g as u8, b_: col.b as u8, alpha: col.a.unwrap_or(255) as u8, _pad: 0 });
        }
    }
    let img = ffi::fdl_image { data: rgb.as_ptr(), width: width as i32, height: height as i32, row_stride: 3 * width as i64, mem: ffi::FDL_MEM_HOST, _pad: 0 };
    let mut out = vec![0u8; 4 * width as usize * height as usize];
    // the reference panics where imageproc does (an empty rectangle): keep that behaviour at this boundary
    ffi::check(unsafe { ffi::fdl_render_to_image(default_device(), &img, prims.as_ptr(), prims.len() as i32, out.as_mut_ptr(), out.len(), ffi::FDL_MEM_HOST) })
        .expect("render_to_image");
    DynamicImage::ImageRgba8(RgbaImage::from_raw(width, height, out).unwrap())
}
