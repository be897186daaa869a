//! `convert_image_to_mat` (reference utils.rs:8-21: imdecode(IMREAD_COLOR) + cvt_color(BGR2RGB)) on the B200: the JPEG is decoded
//! by the device decoder of libfdl_b200.so (bit-exact with OpenCV's libjpeg defaults) and handed back as the 8UC3 RGB `Mat` the
//! reference returns.  `Frame` is the device-resident alternative: decode (or upload) once, pass `frame.image()` to every infer.
use super::ffi;
use anyhow::Error;
use opencv::core::{Mat, Scalar, CV_8UC3};
use opencv::prelude::*;

/// Device used by the free functions and by `new()` of the three models: `FDL_DEVICE` (default 0).
pub fn default_device() -> i32 {
    std::env::var("FDL_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0)
}

pub fn convert_image_to_mat(im_bytes: &[u8]) -> Result<Mat, Error> {
    let (mut w, mut h) = (0i32, 0i32);
    ffi::check(unsafe { ffi::fdl_jpeg_info(im_bytes.as_ptr(), im_bytes.len(), &mut w, &mut h, std::ptr::null_mut()) })?;
    let mut rgb = Mat::new_rows_cols_with_default(h, w, CV_8UC3, Scalar::all(0.0))?;
    let cap = (w as usize) * (h as usize) * 3;
    {
        let bytes = rgb.data_bytes_mut()?;        // freshly allocated: continuous, rows of w * 3 bytes
        ffi::check(unsafe { ffi::fdl_decode_jpeg(default_device(), im_bytes.as_ptr(), im_bytes.len(), bytes.as_mut_ptr(), cap, &mut w, &mut h) })?;
    }
    Ok(rgb)
}

/// A frame staged once on the device (fdl_frame): lib.rs:20-40 hands the same image to four `infer` calls.
pub struct Frame { handle: *mut ffi::fdl_frame }
unsafe impl Send for Frame {}

impl Frame {
    pub fn new() -> Result<Frame, Error> {
        let mut h = std::ptr::null_mut();
        ffi::check(unsafe { ffi::fdl_frame_create(default_device(), &mut h) })?;
        Ok(Frame { handle: h })
    }
    pub fn from_mat(image: &Mat) -> Result<Frame, Error> {
        let f = Frame::new()?;
        let img = ffi::image_of(image)?;
        ffi::check(unsafe { ffi::fdl_frame_upload(f.handle, &img) })?;
        Ok(f)
    }
    pub fn from_jpeg(im_bytes: &[u8]) -> Result<Frame, Error> {
        let f = Frame::new()?;
        ffi::check(unsafe { ffi::fdl_frame_upload_jpeg(f.handle, im_bytes.as_ptr(), im_bytes.len()) })?;
        Ok(f)
    }
    /// (width, height), as `Mat::size()`.
    pub fn size(&self) -> Result<(i32, i32), Error> {
        let img = self.image()?;
        Ok((img.width, img.height))
    }
    pub(crate) fn image(&self) -> Result<ffi::fdl_image, Error> {
        let mut img = ffi::fdl_image { data: std::ptr::null(), width: 0, height: 0, row_stride: 0, mem: 0, _pad: 0 };
        ffi::check(unsafe { ffi::fdl_frame_image(self.handle, &mut img) })?;
        Ok(img)
    }
}
impl Drop for Frame { fn drop(&mut self) { unsafe { ffi::fdl_frame_destroy(self.handle) } } }
