// prepost_kernels.cu -- see prepost_kernels.cuh.  Compiled with -fmad=false.
#include "prepost_kernels.cuh"

#include <atomic>

namespace fdl {

void count_launch();

namespace {

// Opt-in shared memory sizes are a per-device function attribute: remember per device (a process may hold handles on several
// GPUs through the C ABI's `device` argument), set it before the first launch there.  Two racing first callers both set it.
template <typename K>
void opt_in_smem_once(std::atomic<unsigned long long>& done, K kernel, int bytes) {
  int d = 0;
  cudaGetDevice(&d);
  const unsigned long long bit = 1ull << (d & 63);
  if (done.load(std::memory_order_acquire) & bit) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  done.fetch_or(bit, std::memory_order_release);
}

// ------------------------------------------------------------------------------------------------
__global__ void anchors_kernel(SsdOptions opt, float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x, y;
  ssd_anchor(opt, i, &x, &y);
  out[2 * i] = x;
  out[2 * i + 1] = y;
}

// ------------------------------------------------------------------------------------------------
__global__ void i2t_setup_kernel(const fdl_rect* rois, const int* slot_frame, const int* slot_valid, int n, int img_w, int img_h,
                                 int out_w, int out_h, int keep_aspect, double range_min, double range_max, int flip_mode,
                                 I2TParams* params, const int* n_active) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_active) n = min(n, *n_active);
  if (i >= n) return;
  bool flip = flip_mode == 1 || (flip_mode == 2 && (i & 1));
  int frame = slot_frame ? slot_frame[i] : i;
  I2TParams P;
  i2t_setup(rois ? &rois[i] : nullptr, img_w, img_h, out_w, out_h, keep_aspect != 0, range_min, range_max, flip, frame, &P);
  if (slot_valid && !slot_valid[i]) P.valid = 0;
  params[i] = P;
}

// kPxPerThread output pixels (3 channels) per thread and work item (all in flight together).  The source taps are gathered straight from the u8
// frame (L2/texture path); the f32 tensor is written once.  The normalisation goes through a 256-entry table of the
// exact f64 expression (one f64 division per thread and CTA instead of three per pixel).
template <int kPxPerThread>
__global__ void __launch_bounds__(256, 4) i2t_kernel(const uint8_t* __restrict__ frames, long long frame_stride, long long row_stride,
                                                  const I2TParams* __restrict__ params, int n, int out_w, int out_h,
                                                  float* __restrict__ out, long long out_bstride, uint8_t* __restrict__ out_u8,
                                                  const int* n_active) {
  if (n_active) n = min(n, *n_active);
  __shared__ I2TParams P;
  __shared__ float s_lut[256];
  const int px_per_item = 256 * kPxPerThread;
  const int blocks_per_slot = (out_w * out_h + px_per_item - 1) / px_per_item;
  const long long items = (long long)n * blocks_per_slot;
  double lut_min = 0.0, lut_max = 0.0;
  bool lut_built = false;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int slot = (int)(item / blocks_per_slot), blk = (int)(item - (long long)slot * blocks_per_slot);
    __syncthreads();
    {
      const int* src = reinterpret_cast<const int*>(&params[slot]);
      int* dst = reinterpret_cast<int*>(&P);
      for (int i = threadIdx.x; i < (int)(sizeof(I2TParams) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    if (P.valid == 2) continue;                  // another launch owns this slot (see eye_split_kernel)
    if (!lut_built || lut_min != P.range_min || lut_max != P.range_max) {     // uniform over the CTA
      for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = i2t_normalise(i, P.range_min, P.range_max);
      lut_min = P.range_min; lut_max = P.range_max; lut_built = true;
      __syncthreads();
    }
    const ImgSrc src = img_src(frames + (long long)P.frame * frame_stride, row_stride);
#pragma unroll
    for (int k = 0; k < kPxPerThread; ++k) {
      const int pix = blk * px_per_item + k * 256 + threadIdx.x;
      if (pix >= out_w * out_h) break;
      const int oy = pix / out_w, ox = pix - oy * out_w;
      Px3 p;
      if (P.valid) p = i2t_pixel(P, src, ox, oy);
      else { p.r = p.g = p.b = 0; }
      float* o = out + (long long)slot * out_bstride + (long long)pix * 3;
      o[0] = s_lut[p.r & 255];
      o[1] = s_lut[p.g & 255];
      o[2] = s_lut[p.b & 255];
      if (out_u8) {
        uint8_t* u = out_u8 + ((long long)slot * out_w * out_h + pix) * 3;
        u[0] = (uint8_t)p.r; u[1] = (uint8_t)p.g; u[2] = (uint8_t)p.b;
      }
    }
  }
}

// Row-staged variant for the detector's letterbox (roi = None): one CTA per output row.  When the slot's
// transform is the plain letterbox (identity warp, border, ONE real resize, no flip) the CTA stages the two
// source rows this output row interpolates between in shared memory with fully coalesced 16-byte loads --
// every source byte crosses the memory system (HBM, or PCIe when `frames` is mapped pinned host memory) exactly
// once, and only the rows the 2x2-tap resize touches are read at all (27 % of a 1080p frame at S = 256).
// The arithmetic is the same fixed-point resize as i2t_pixel (bit-exact); any other slot takes the generic
// per-pixel path inside the same kernel.
constexpr int kRowsPerItem = 4;    // output rows per work item of the row-staged letterbox kernel

template <int kMinB>
__global__ void __launch_bounds__(256, kMinB) i2t_rows_kernel(const uint8_t* __restrict__ frames, long long frame_stride, long long row_stride,
                                                       const I2TParams* __restrict__ params, int n, int out_w, int out_h,
                                                       float* __restrict__ out, long long out_bstride, const int* n_active,
                                                       const uint8_t* __restrict__ compact, const int* __restrict__ row_pos, long long compact_fstride,
                                                       int rows_per_item) {
  extern __shared__ __align__(16) uint8_t s_rows[];     // [2 * rows_per_item][row_pad]
  if (n_active) n = min(n, *n_active);
  __shared__ I2TParams P;
  __shared__ float s_lut[256];                          // i2t_normalise for every grey level (exact: same f64 expression)
  __shared__ double s_scale[2];                         // resize scales of the slot: x, y (one division per slot, not per pixel)
  __shared__ int s_simple;
  const int groups = (out_h + rows_per_item - 1) / rows_per_item;
  const long long items = (long long)n * groups;
  int cur_slot = -1;
  double lut_min = 0.0, lut_max = 0.0;                  // the range the LUT was last built for (uniform over the CTA)
  bool lut_built = false;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
  const int slot = (int)(item / groups), oy_first = (int)(item - (long long)slot * groups) * rows_per_item;
  const int nrows = min(rows_per_item, out_h - oy_first);
  __syncthreads();
  if (slot != cur_slot) {
    const int* src = reinterpret_cast<const int*>(&params[slot]);
    int* dst = reinterpret_cast<int*>(&P);
    for (int i = threadIdx.x; i < (int)(sizeof(I2TParams) / 4); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    if (!lut_built || lut_min != P.range_min || lut_max != P.range_max) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = i2t_normalise(i, P.range_min, P.range_max);
      lut_min = P.range_min; lut_max = P.range_max; lut_built = true;
    }
    if (threadIdx.x == 0) {
      const int bw = P.warp_w + 2 * P.pad_h, bh = P.warp_h + 2 * P.pad_v;
      bool simple = P.valid && P.has_r2 && !P.flip && P.warp_w == P.src_w && P.warp_h == P.src_h &&
                    (!P.has_r1 || (bw == P.r1_w && bh == P.r1_h)) && !(P.r1_w == out_w && P.r1_h == out_h);
      if (simple) {
        const double e = 1e-9;   // identity warp: the solve reproduces I up to rounding (exact copy)
        simple = fabs(P.Mi[0] - 1.0) < e && fabs(P.Mi[4] - 1.0) < e && fabs(P.Mi[8] - 1.0) < e && fabs(P.Mi[1]) < e && fabs(P.Mi[2]) < e * P.src_w &&
                 fabs(P.Mi[3]) < e && fabs(P.Mi[5]) < e * P.src_h && fabs(P.Mi[6]) < e && fabs(P.Mi[7]) < e;
      }
      s_simple = simple ? 1 : 0;
      s_scale[0] = (double)P.r1_w / (double)out_w;
      s_scale[1] = (double)P.r1_h / (double)out_h;
    }
    cur_slot = slot;
    __syncthreads();
  }
  const uint8_t* img = frames + (long long)P.frame * frame_stride;
  if (!s_simple) {
    for (int r = 0; r < nrows; ++r) {
      const int oy = oy_first + r;
      float* orow = out + (long long)slot * out_bstride + (long long)oy * out_w * 3;
      for (int ox = threadIdx.x; ox < out_w; ox += blockDim.x) {
        Px3 p;
        if (P.valid) p = i2t_pixel(P, img_src(img, row_stride), ox, oy);
        else { p.r = p.g = p.b = 0; }
        orow[3 * ox] = i2t_normalise(p.r, P.range_min, P.range_max);
        orow[3 * ox + 1] = i2t_normalise(p.g, P.range_min, P.range_max);
        orow[3 * ox + 2] = i2t_normalise(p.b, P.range_min, P.range_max);
      }
    }
    continue;
  }
  const int ph = P.has_r1 ? P.pad_h : 0, pv = P.has_r1 ? P.pad_v : 0;
  const int src_w = P.src_w, src_h = P.src_h;
  const int row_bytes = src_w * 3;
  const int row_pad = (row_bytes + 15) & ~15;
  // vertical taps of the item's rows (every thread: a handful of instructions with the division hoisted); source rows
  // outside the frame are the constant-0 border: -1
  int sy[kRowsPerItem][2], wb[kRowsPerItem][2];
  bool any_row = false;
#pragma unroll
  for (int r = 0; r < kRowsPerItem; ++r) {
    int y0, y1;
    resize_coeff_scaled(oy_first + (r < nrows ? r : 0), s_scale[1], P.r1_h, false, &y0, &y1, &wb[r][0], &wb[r][1]);
    y0 -= pv; y1 -= pv;
    sy[r][0] = (r < nrows && y0 >= 0 && y0 < src_h) ? y0 : -1;
    sy[r][1] = (r < nrows && y1 >= 0 && y1 < src_h) ? y1 : -1;
    any_row = any_row || sy[r][0] >= 0 || sy[r][1] >= 0;
  }
  float* const obase = out + (long long)slot * out_bstride + (long long)oy_first * out_w * 3;
  if (!any_row) {
    // letterbox bars: every tap is border => the whole item is the grey level 0 (rows are contiguous: 16-byte stores)
    const float z = s_lut[0];
    const int nf = nrows * out_w * 3;
    if ((reinterpret_cast<uintptr_t>(obase) & 15) == 0) {
      for (int i = threadIdx.x; i < (nf >> 2); i += blockDim.x) reinterpret_cast<float4*>(obase)[i] = make_float4(z, z, z, z);
      for (int i = (nf & ~3) + threadIdx.x; i < nf; i += blockDim.x) obase[i] = z;
    } else {
      for (int i = threadIdx.x; i < nf; i += blockDim.x) obase[i] = z;
    }
    continue;
  }
  // ---- stage the (up to) 2 * nrows source rows: every load of the item is issued before the barrier ----
#pragma unroll
  for (int r = 0; r < kRowsPerItem; ++r) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (sy[r][h] < 0) continue;
      const uint8_t* g = img + (long long)sy[r][h] * row_stride;
      if (row_pos && row_pos[sy[r][h]] >= 0)   // rows gathered into device memory by the copy engine (anything missing is still read in place)
        g = compact + (long long)P.frame * compact_fstride + (long long)row_pos[sy[r][h]] * row_bytes;
      uint8_t* d = s_rows + (2 * r + h) * row_pad;
      if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        // asynchronous 16-byte copies: nothing waits until every row of the item has been requested
        const int nv = row_bytes >> 4;
        const unsigned sd = (unsigned)__cvta_generic_to_shared(d);
        for (int i = threadIdx.x; i < nv; i += blockDim.x)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sd + 16 * i), "l"(g + 16 * (size_t)i) : "memory");
        for (int i = (nv << 4) + threadIdx.x; i < row_bytes; i += blockDim.x) d[i] = g[i];
      } else {
        for (int i = threadIdx.x; i < row_bytes; i += blockDim.x) d[i] = g[i];
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  for (int ox = threadIdx.x; ox < out_w; ox += blockDim.x) {
    int x0, x1, a0, a1;
    resize_coeff_scaled(ox, s_scale[0], P.r1_w, true, &x0, &x1, &a0, &a1);
    const int sx0 = x0 - ph, sx1 = x1 - ph;
    const bool cx0 = sx0 >= 0 && sx0 < src_w, cx1 = sx1 >= 0 && sx1 < src_w;
    const int o0 = cx0 ? 3 * sx0 : 0, o1 = cx1 ? 3 * sx1 : 0;
#pragma unroll
    for (int r = 0; r < kRowsPerItem; ++r) {
      if (r >= nrows) break;
      const bool in0 = sy[r][0] >= 0, in1 = sy[r][1] >= 0;
      const uint8_t* r0 = s_rows + (2 * r) * row_pad;
      const uint8_t* r1 = r0 + row_pad;
      float* orow = obase + (long long)r * out_w * 3 + 3 * ox;
      int v[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int p00 = (in0 && cx0) ? r0[o0 + c] : 0, p01 = (in0 && cx1) ? r0[o1 + c] : 0;
        const int p10 = (in1 && cx0) ? r1[o0 + c] : 0, p11 = (in1 && cx1) ? r1[o1 + c] : 0;
        v[c] = resize_mix(p00, p01, p10, p11, a0, a1, wb[r][0], wb[r][1]);
      }
      orow[0] = s_lut[v[0] & 255];
      orow[1] = s_lut[v[1] & 255];
      orow[2] = s_lut[v[2] & 255];
    }
  }
  }  // item loop
}

// Tile-staged variant for ROI warps whose frames live in mapped pinned host memory: one 32x32 output tile per
// work item.  The source bounding box of the tile (the rectangle of warp space it samples, pushed through the
// inverse perspective matrix) is copied into shared memory with aligned 16-byte loads -- each source byte crosses
// PCIe once per tile instead of once per tap -- and the pixels are then evaluated by the very same i2t_pixel code
// reading through ImgSrc (anything outside the staged box falls back to a direct load, so staging is purely a
// traffic optimisation and cannot change a result).
constexpr int kTileW = 64, kTileH = 32;   // wide tiles: longer source row segments => fewer partially used 128-byte lines over PCIe
constexpr int kTileSmem = 96 * 1024;

__global__ void __launch_bounds__(256) i2t_tile_kernel(const uint8_t* __restrict__ frames, long long frame_stride, long long row_stride,
                                                       const I2TParams* __restrict__ params, int n, int out_w, int out_h,
                                                       float* __restrict__ out, long long out_bstride, const int* n_active) {
  extern __shared__ __align__(16) uint8_t s_tile[];
  if (n_active) n = min(n, *n_active);
  __shared__ I2TParams P;
  __shared__ int s_box[6];   // x0, y0, x1, y1 (pixels, inclusive; x1 < x0: nothing staged), start byte, pitch
  const int tiles_x = (out_w + kTileW - 1) / kTileW, tiles_y = (out_h + kTileH - 1) / kTileH;
  const long long items = (long long)n * tiles_x * tiles_y;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int slot = (int)(item / (tiles_x * tiles_y));
    const int t = (int)(item - (long long)slot * tiles_x * tiles_y);
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    __syncthreads();
    {
      const int* src = reinterpret_cast<const int*>(&params[slot]);
      int* dst = reinterpret_cast<int*>(&P);
      for (int i = threadIdx.x; i < (int)(sizeof(I2TParams) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    if (P.valid == 2) continue;                  // another launch owns this slot (see eye_split_kernel)
    const uint8_t* img = frames + (long long)P.frame * frame_stride;
    const int ox0 = tx * kTileW, oy0 = ty * kTileH, ox1 = min(ox0 + kTileW, out_w) - 1, oy1 = min(oy0 + kTileH, out_h) - 1;
    if (threadIdx.x == 0) {
      s_box[0] = 0; s_box[1] = 0; s_box[2] = -1; s_box[3] = -1; s_box[4] = 0; s_box[5] = 0;
      bool ok = P.valid != 0;
      // output tile -> rectangle of warp space it samples
      int wx0 = P.flip ? P.out_w - 1 - ox1 : ox0, wx1 = P.flip ? P.out_w - 1 - ox0 : ox1, wy0 = oy0, wy1 = oy1;
      if (ok && P.has_r2 && !(P.r1_w == P.out_w && P.r1_h == P.out_h)) {
        int a, b2, c0, c1;
        resize_coeff(wx0, P.out_w, P.r1_w, true, &a, &b2, &c0, &c1); const int nx0 = a;
        resize_coeff(wx1, P.out_w, P.r1_w, true, &a, &b2, &c0, &c1); const int nx1 = b2;
        resize_coeff(wy0, P.out_h, P.r1_h, false, &a, &b2, &c0, &c1); const int ny0 = a;
        resize_coeff(wy1, P.out_h, P.r1_h, false, &a, &b2, &c0, &c1); const int ny1 = b2;
        wx0 = nx0; wx1 = nx1; wy0 = ny0; wy1 = ny1;
      }
      if (ok && P.has_r2 && P.has_r1) {
        const int bw = P.warp_w + 2 * P.pad_h, bh = P.warp_h + 2 * P.pad_v;
        if (bw == P.r1_w && bh == P.r1_h) {
          wx0 = max(wx0 - P.pad_h, 0); wx1 = min(wx1 - P.pad_h, P.warp_w - 1);
          wy0 = max(wy0 - P.pad_v, 0); wy1 = min(wy1 - P.pad_v, P.warp_h - 1);
          if (wx1 < wx0 || wy1 < wy0) ok = false;   // the tile lies entirely in the letterbox border
        } else {
          ok = false;                               // a real first resize: not staged (direct loads)
        }
      }
      if (ok) {
        double lox = 1e30, loy = 1e30, hix = -1e30, hiy = -1e30;
        for (int c = 0; c < 4; ++c) {
          const double x = (c & 1) ? wx1 : wx0, y = (c & 2) ? wy1 : wy0;
          const double w = P.Mi[6] * x + P.Mi[7] * y + P.Mi[8];
          if (!(w > 1e-6)) { ok = false; break; }
          const double sx = (P.Mi[0] * x + P.Mi[1] * y + P.Mi[2]) / w, sy = (P.Mi[3] * x + P.Mi[4] * y + P.Mi[5]) / w;
          lox = dmin(lox, sx); hix = dmax(hix, sx); loy = dmin(loy, sy); hiy = dmax(hiy, sy);
        }
        if (ok && hix - lox < 4096.0 && hiy - loy < 4096.0 && lox > -1e6 && loy > -1e6) {
          const int x0 = max((int)floor(lox) - 1, 0), y0 = max((int)floor(loy) - 1, 0);
          const int x1 = min((int)ceil(hix) + 2, P.src_w - 1), y1 = min((int)ceil(hiy) + 2, P.src_h - 1);
          if (x1 >= x0 && y1 >= y0) {
            const int sb = (3 * x0) & ~15;
            int eb = (3 * (x1 + 1) + 15) & ~15;
            if (eb > (int)row_stride) eb = (int)row_stride;
            const int pitch = eb - sb;
            if ((long long)pitch * (y1 - y0 + 1) <= kTileSmem && pitch > 0) {
              s_box[0] = x0; s_box[1] = y0; s_box[2] = x1; s_box[3] = y1; s_box[4] = sb; s_box[5] = pitch;
            }
          }
        }
      }
    }
    __syncthreads();
    ImgSrc src = img_src(img, row_stride);
    if (s_box[2] >= s_box[0]) {
      const int y0 = s_box[1], rows = s_box[3] - s_box[1] + 1, sb = s_box[4], pitch = s_box[5];
      const bool vec = ((reinterpret_cast<uintptr_t>(img) | (uintptr_t)row_stride) & 15) == 0 && (pitch & 15) == 0;
      if (vec) {
        const int per_row = pitch >> 4;
        for (int i = threadIdx.x; i < per_row * rows; i += blockDim.x) {
          const int r = i / per_row, c = i - r * per_row;
          reinterpret_cast<uint4*>(s_tile + (long long)r * pitch)[c] = __ldg(reinterpret_cast<const uint4*>(img + (long long)(y0 + r) * row_stride + sb) + c);
        }
      } else {
        for (int i = threadIdx.x; i < pitch * rows; i += blockDim.x) {
          const int r = i / pitch, c = i - r * pitch;
          s_tile[(long long)r * pitch + c] = img[(long long)(y0 + r) * row_stride + sb + c];
        }
      }
      src.tile = s_tile; src.tx0 = s_box[0]; src.ty0 = y0; src.tx1 = s_box[2]; src.ty1 = s_box[3]; src.tpitch = pitch; src.tsb = sb;
      // pixels of the box whose bytes were clipped by the row end are served by the fallback path
      src.tx1 = min(src.tx1, (sb + pitch) / 3 - 1);
    }
    __syncthreads();
    const int tw = ox1 - ox0 + 1, th = oy1 - oy0 + 1;
    for (int i = threadIdx.x; i < tw * th; i += blockDim.x) {
      const int py = i / tw, px = i - py * tw;
      const int ox = ox0 + px, oy = oy0 + py;
      Px3 p;
      if (P.valid) p = i2t_pixel(P, src, ox, oy);
      else { p.r = p.g = p.b = 0; }
      float* o = out + (long long)slot * out_bstride + ((long long)oy * out_w + ox) * 3;
      o[0] = i2t_normalise(p.r, P.range_min, P.range_max);
      o[1] = i2t_normalise(p.g, P.range_min, P.range_max);
      o[2] = i2t_normalise(p.b, P.range_min, P.range_max);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// SSD post-processing: one CTA per frame.
constexpr int kPostThreads = 128;

// Ordered compaction inside a CTA: returns this thread's exclusive rank among the flagged threads
// of the current chunk and the chunk total.  s_warp: kPostThreads/32 ints of scratch.
__device__ __forceinline__ int block_rank(bool flag, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned m = __ballot_sync(0xffffffffu, flag);
  int r = __popc(m & ((1u << lane) - 1u));
  if (lane == 0) s_warp[warp] = __popc(m);
  __syncthreads();
  int off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kPostThreads / 32; ++w) {
    int c = s_warp[w];
    if (w < warp) off += c;
    tot += c;
  }
  __syncthreads();
  *total = tot;
  return off + r;
}

__global__ void __launch_bounds__(kPostThreads) ssd_postprocess_kernel(const SsdPostArgs a) {
  extern __shared__ int s_mem[];
  const int N = a.N, tid = threadIdx.x, b = blockIdx.x;
  int* s_idx = s_mem;                                        // survivor -> anchor index (ascending)
  float* s_score = reinterpret_cast<float*>(s_mem + N);      // survivor -> score
  int* s_rem[2] = {s_mem + 2 * N, s_mem + 3 * N};            // remaining lists (survivor positions, score order)
  int* s_cand = s_mem + 4 * N;                               // candidates of the current cluster
  __shared__ int s_warp[kPostThreads / 32];
  __shared__ float s_top[16];

  const float* reg = a.reg + (long long)b * a.reg_bstride;
  const float* cls = a.cls + (long long)b * a.cls_bstride;

  // 1. get_sigmoid_score + convert_to_detections: survivors in ascending anchor order
  int n = 0;
  for (int base = 0; base < N; base += kPostThreads) {
    int i = base + tid;
    bool keep = false;
    float sc = 0.f;
    if (i < N) {
      sc = ssd_score(cls[i]);
      if (sc > 0.5f) {
        float raw[16], d[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) raw[k] = reg[(long long)i * 16 + k];
        decode_box(raw, a.anchors[2 * i], a.anchors[2 * i + 1], a.scale, d);
        keep = d[2] > d[0] && d[3] > d[1];
      }
    }
    int tot;
    int r = block_rank(keep, s_warp, &tot);
    if (keep) { s_idx[n + r] = i; s_score[n + r] = sc; }
    n += tot;
  }
  __syncthreads();
  if (a.n_surv && tid == 0) a.n_surv[b] = n;
  if (a.surv_anchor) for (int j = tid; j < n && j < a.cap_surv; j += kPostThreads) {
    a.surv_anchor[(long long)b * a.cap_surv + j] = s_idx[j];
    a.surv_cluster[(long long)b * a.cap_surv + j] = -1;
  }

  // 2. stable sort by score, descending (nms.rs:134-137): rank = number of elements that precede
  for (int j = tid; j < n; j += kPostThreads) {
    float sj = s_score[j];
    int rank = 0;
    for (int k = 0; k < n; ++k) {
      float sk = s_score[k];
      rank += (sk > sj || (sk == sj && k < j)) ? 1 : 0;
    }
    s_rem[0][rank] = j;
  }
  __syncthreads();

  // letterbox removal constants (transform.rs:115-142): scales in f64, cast to f32
  double pl, pt, pr, pb;
  if (a.params) { pl = a.params[b].pad[0]; pt = a.params[b].pad[1]; pr = a.params[b].pad[2]; pb = a.params[b].pad[3]; }
  else if (a.padding4) { pl = a.padding4[4 * b]; pt = a.padding4[4 * b + 1]; pr = a.padding4[4 * b + 2]; pb = a.padding4[4 * b + 3]; }
  else { pl = pt = pr = pb = 0.0; }
  const float left = (float)pl, top = (float)pt;
  const float h_scale = (float)(1.0 - (pl + pr)), v_scale = (float)(1.0 - (pt + pb));

  // 3. weighted_non_maximum_suppression (nms.rs:56-124)
  int cur = 0, n_rem = n, n_out = 0;
  const double thr = (double)0.3f;
  while (n_rem > 0) {
    const int* rem = s_rem[cur];
    int* next = s_rem[cur ^ 1];
    const int top_pos = rem[0];
    if (tid == 0) {
      float raw[16];
      int ti = s_idx[top_pos];
      for (int k = 0; k < 16; ++k) raw[k] = reg[(long long)ti * 16 + k];
      decode_box(raw, a.anchors[2 * ti], a.anchors[2 * ti + 1], a.scale, s_top);
    }
    __syncthreads();
    int n_cand = 0, n_next = 0;
    for (int base = 0; base < n_rem; base += kPostThreads) {
      int j = base + tid;
      bool is_cand = false, is_rem = false;
      int pos = 0;
      if (j < n_rem) {
        pos = rem[j];
        int ai = s_idx[pos];
        float raw[16], d[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) raw[k] = reg[(long long)ai * 16 + k];
        decode_box(raw, a.anchors[2 * ai], a.anchors[2 * ai + 1], a.scale, d);
        double sim = overlap_similarity(d, s_top);
        is_cand = sim > thr;
        is_rem = !is_cand;
      }
      int tc, tr;
      int rc = block_rank(is_cand, s_warp, &tc);
      int rr = block_rank(is_rem, s_warp, &tr);
      if (is_cand) s_cand[n_cand + rc] = pos;
      if (is_rem) next[n_next + rr] = pos;
      n_cand += tc;
      n_next += tr;
    }
    __syncthreads();
    // weighted merge: thread k < 16 owns coordinate k and accumulates in candidate order, in f32
    if (tid < 16 && n_out < a.max_out) {
      float v;
      if (n_cand > 0) {
        float w = 0.f, total = 0.f;
        for (int c = 0; c < n_cand; ++c) {
          int pos = s_cand[c];
          int ai = s_idx[pos];
          float sc = s_score[pos];
          float raw[16], d[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) raw[k] = reg[(long long)ai * 16 + k];
          decode_box(raw, a.anchors[2 * ai], a.anchors[2 * ai + 1], a.scale, d);
          total += sc;
          float dk = 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) if (k == tid) dk = d[k];
          w += dk * sc;
        }
        v = w / total;
      } else {
        v = s_top[tid];
      }
      // detection_letterbox_removal
      v = (tid & 1) ? (v - top) / v_scale : (v - left) / h_scale;
      fdl_detection* o = reinterpret_cast<fdl_detection*>(a.det_base + (long long)b * a.det_stride) + n_out;
      o->data[tid] = v;
      if (tid == 0) { o->score = s_score[top_pos]; o->anchor = s_idx[top_pos]; }
    }
    if (a.surv_cluster) for (int c = tid; c < n_cand; c += kPostThreads) {
      int pos = s_cand[c];
      if (pos < a.cap_surv) a.surv_cluster[(long long)b * a.cap_surv + pos] = n_out;
    }
    ++n_out;
    __syncthreads();
    if (n_cand == 0) break;   // "number of indexed scores didn't change" (nms.rs:117-119)
    n_rem = n_next;
    cur ^= 1;
  }
  if (tid == 0) {
    *reinterpret_cast<int32_t*>(a.ndet_base + (long long)b * a.ndet_stride) = min(n_out, a.max_out);
    if (a.n_total) a.n_total[(long long)b * a.n_total_stride] = n_out;
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void face_select_kernel(fdl_frame_result* frames, int B, int max_faces, int* slot_frame, int* slot_face, int* n_faces,
                                   int* n_eyes) {
  // single CTA: exclusive scan of min(n_detections, max_faces) over frames
  __shared__ int s_part[1024];
  const int tid = threadIdx.x, T = blockDim.x;
  const int per = (B + T - 1) / T;
  int lo = tid * per, hi = min(lo + per, B);
  int sum = 0;
  for (int b = lo; b < hi; ++b) sum += min(frames[b].n_detections, max_faces);
  s_part[tid] = sum;
  __syncthreads();
  // Hillis-Steele inclusive scan
  for (int off = 1; off < T; off <<= 1) {
    int v = tid >= off ? s_part[tid - off] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int base = s_part[tid] - sum;
  for (int b = lo; b < hi; ++b) {
    int nf = min(frames[b].n_detections, max_faces);
    frames[b].n_faces = nf;
    for (int f = 0; f < nf; ++f) { slot_frame[base + f] = b; slot_face[base + f] = f; }
    base += nf;
  }
  if (tid == T - 1) { *n_faces = s_part[tid]; *n_eyes = 2 * s_part[tid]; }
}

__global__ void face_roi_kernel(const fdl_frame_result* frames, const int* slot_frame, const int* slot_face, int max_slots,
                                int max_faces, int img_w, int img_h, fdl_rect* rois, int* slot_valid, fdl_face_result* faces,
                                const int* n_faces) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= min(max_slots, *n_faces)) return;
  int fr = slot_frame[s], fa = slot_face[s];
  fdl_rect r;
  bool ok = face_detection_to_roi(frames[fr].detections[fa].data, img_w, img_h, FDL_SIZE_MODE_NONE, &r);
  if (!ok) { r.x_center = r.y_center = r.width = r.height = r.rotation = 0.0; r.normalized = 1; r._pad = 0; }
  rois[s] = r;
  slot_valid[s] = ok ? 1 : 0;
  faces[(long long)fr * max_faces + fa].face_roi = r;
}

__global__ void __launch_bounds__(128) landmark_post_kernel(const float* raw, long long raw_bstride, const float* flag,
                                                            long long flag_bstride, const I2TParams* params, const fdl_rect* rois,
                                                            const int* slot_frame, const int* slot_face, int max_slots, int max_faces,
                                                            int tensor_w, int tensor_h, fdl_face_result* faces, fdl_rect* eye_rois,
                                                            int* eye_frame, int* eye_valid, const int* n_faces, int refine) {
  int s = blockIdx.x;
  if (s >= min(max_slots, *n_faces)) return;
  __shared__ ProjectParams pp;
  __shared__ int s_has;
  const I2TParams& P = params[s];
  fdl_face_result* out = &faces[(long long)slot_frame[s] * max_faces + slot_face[s]];
  const float* r = raw + (long long)s * raw_bstride;
  if (threadIdx.x == 0) {
    float logit = flag[(long long)s * flag_bstride];
    // face_landmark.rs:292-296: sigmoid(flag) <= DETECTION_THRESHOLD -> no landmarks
    int has = P.valid && !(sigmoid_f32(logit) <= 0.5f);
    s_has = has;
    out->face_flag_logit = logit;
    out->has_landmarks = has;
    project_setup(tensor_w, tensor_h, P.src_w, P.src_h, P.pad, &rois[s], false, &pp);
    fdl_rect er[2];
    bool ok[2] = {false, false};
    if (has) {
      const int idx[4] = {33, 133, 362, 263};   // iris_landmark.rs:29-35
      double xy[8];
      for (int k = 0; k < 4; ++k) {
        float q[3];
        project_point(pp, r + 3 * idx[k], q);
        xy[2 * k] = (double)q[0]; xy[2 * k + 1] = (double)q[1];
      }
      ok[0] = eye_roi(xy[0], xy[1], xy[2], xy[3], P.src_w, P.src_h, &er[0]);
      ok[1] = eye_roi(xy[4], xy[5], xy[6], xy[7], P.src_w, P.src_h, &er[1]);
    }
    for (int e = 0; e < 2; ++e) {
      if (!ok[e]) { er[e].x_center = er[e].y_center = er[e].width = er[e].height = er[e].rotation = 0.0; er[e].normalized = 1; er[e]._pad = 0; }
      eye_rois[2 * s + e] = er[e];
      eye_valid[2 * s + e] = ok[e] ? 1 : 0;
      eye_frame[2 * s + e] = P.frame;
      out->eye_roi[e] = er[e];
    }
  }
  __syncthreads();
  if (!s_has) return;
  for (int k = threadIdx.x; k < FDL_NUM_FACE_LANDMARKS; k += blockDim.x) {
    float q[3];
    project_point(pp, r + 3 * k, q);
    out->landmarks[3 * k] = q[0]; out->landmarks[3 * k + 1] = q[1]; out->landmarks[3 * k + 2] = q[2];
    // update_face_landmarks_with_iris_results starts from a clone of the face landmarks (iris_landmark.rs:387); the eye
    // contours are scattered over it by iris_post_kernel
    if (refine) { out->refined_landmarks[3 * k] = q[0]; out->refined_landmarks[3 * k + 1] = q[1]; out->refined_landmarks[3 * k + 2] = q[2]; }
  }
}

__global__ void __launch_bounds__(96) iris_post_kernel(const float* contour, long long contour_bstride, const float* iris,
                                                       long long iris_bstride, const I2TParams* params, const fdl_rect* eye_rois,
                                                       const int* eye_valid, const int* slot_frame, const int* slot_face,
                                                       int max_eye_slots, int max_faces, int tensor_w, int tensor_h,
                                                       fdl_face_result* faces, const int* n_eyes, int refine, double focal_length_mm) {
  int es = blockIdx.x;
  if (es >= min(max_eye_slots, *n_eyes)) return;
  if (!eye_valid[es]) return;
  __shared__ ProjectParams pp;
  const I2TParams& P = params[es];
  if (!P.valid) return;
  int s = es >> 1, e = es & 1;
  fdl_face_result* out = &faces[(long long)slot_frame[s] * max_faces + slot_face[s]];
  if (threadIdx.x == 0) project_setup(tensor_w, tensor_h, P.src_w, P.src_h, P.pad, &eye_rois[es], e == 1, &pp);
  __syncthreads();
  __shared__ float s_iris[FDL_NUM_IRIS * 3];
  int k = threadIdx.x;
  if (k < FDL_NUM_EYE_CONTOUR) {
    float q[3];
    project_point(pp, contour + (long long)es * contour_bstride + 3 * k, q);
    float* dst = out->eye_contour[e] + 3 * k;
    dst[0] = q[0]; dst[1] = q[1]; dst[2] = q[2];
    if (refine) {
      // update_face_landmarks_with_iris_results (iris_landmark.rs:389-396): contour point k replaces face landmark
      // LEFT_/RIGHT_EYE_TO_FACE_LANDMARK_INDEX[k]; the two eyes' index sets are disjoint, so the two eye blocks of a
      // face never write the same landmark
      float* rl = out->refined_landmarks + 3 * eye_to_face_landmark_index(e, k);
      rl[0] = q[0]; rl[1] = q[1]; rl[2] = q[2];
    }
  } else if (k < FDL_NUM_EYE_CONTOUR + FDL_NUM_IRIS) {
    int q = k - FDL_NUM_EYE_CONTOUR;
    float v[3];
    project_point(pp, iris + (long long)es * iris_bstride + 3 * q, v);
    float* dst = out->iris[e] + 3 * q;
    dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2];
    s_iris[3 * q] = v[0]; s_iris[3 * q + 1] = v[1]; s_iris[3 * q + 2] = v[2];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // get_iris_diameter / get_iris_depth (iris_landmark.rs:401-433) in image pixels / millimetres
    const double d = iris_diameter(s_iris, P.src_w, P.src_h);
    out->iris_diameter_px[e] = d;
    out->iris_depth_mm[e] = focal_length_mm > 0.0 ? iris_depth(s_iris, focal_length_mm, d, P.src_w, P.src_h) : 0.0;
  }
}

// Stand-alone forms of the refinement helpers (f64 Landmark values, as the reference's signatures).
__global__ void refine_landmarks_kernel(const double* face, const double* left, int n_left, const double* right, int n_right, double* out) {
  for (int k = threadIdx.x; k < 3 * FDL_NUM_FACE_LANDMARKS; k += blockDim.x) out[k] = face[k];
  __syncthreads();
  for (int k = threadIdx.x; k < n_left; k += blockDim.x) {
    const int t = eye_to_face_landmark_index(0, k);
    out[3 * t] = left[3 * k]; out[3 * t + 1] = left[3 * k + 1]; out[3 * t + 2] = left[3 * k + 2];
  }
  for (int k = threadIdx.x; k < n_right; k += blockDim.x) {
    const int t = eye_to_face_landmark_index(1, k);
    out[3 * t] = right[3 * k]; out[3 * t + 1] = right[3 * k + 1]; out[3 * t + 2] = right[3 * k + 2];
  }
}
__global__ void iris_metrics_kernel(const double* iris, int img_w, int img_h, double focal_length_mm, double iris_size_px, double* out2) {
  out2[0] = iris_diameter(iris, img_w, img_h);
  out2[1] = iris_size_px > 0.0 ? iris_depth(iris, focal_length_mm, iris_size_px, img_w, img_h) : 0.0;
}

// ------------------------------------------------------------------------------------------------
// Zero-copy host frames: ROI staging.  The face warp and the two eye warps of a face sample (almost always) the same
// neighbourhood of the frame.  Instead of letting each warp kernel fetch its taps over PCIe (overlapping tiles re-read, 32 GB/s
// effective), the source rectangle of the face warp, grown by a margin, is copied ONCE with fully coalesced 16-byte loads into
// the lane's device frame buffer AT THE SAME OFFSETS; the warps then run on that buffer with unchanged coordinates (so their
// arithmetic is untouched).  Eye warps whose own source rectangle is not inside the copied one keep reading the host frame.
constexpr int kFillRows = 16;      // frame rows per work item of roi_fill_kernel

__global__ void __launch_bounds__(256) roi_fill_kernel(const uint8_t* __restrict__ host_frames, uint8_t* __restrict__ dev_frames,
                                                       long long frame_stride, long long row_stride, const I2TParams* __restrict__ params,
                                                       int n, const int* n_active, SrcBox* boxes, int frame_h, int fill_margin_pct, int quad_trim,
                                                       const uint8_t* __restrict__ compact, const int* __restrict__ row_pos, long long compact_fstride) {
  if (n_active) n = min(n, *n_active);
  __shared__ SrcBox s_box;
  __shared__ int s_frame, s_margin;
  const int chunks = (frame_h + kFillRows - 1) / kFillRows;
  const long long items = (long long)n * chunks;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int slot = (int)(item / chunks), ch = (int)(item - (long long)slot * chunks);
    __syncthreads();
    if (threadIdx.x == 0) {
      const I2TParams& P = params[slot];
      int m = 0;
      // a small margin (the eye ROIs lie well inside the face ROI; an eye that does not keeps reading the host frame), clipped
      // to the frame.  A 12 % margin cost more PCIe time than it saved: the face rectangles of 1080p frames are large.
      SrcBox b = roi_stage_box(P, fill_margin_pct, quad_trim != 0, &m);
      s_margin = m;
      s_box = b; s_frame = P.frame;
      if (ch == 0) boxes[slot] = b;
    }
    __syncthreads();
    const SrcBox b = s_box;
    if (b.x1 < b.x0 || b.y1 < b.y0) continue;
    const int r0 = max(b.y0, ch * kFillRows), r1 = min(b.y1, ch * kFillRows + kFillRows - 1);
    if (r1 < r0) continue;
    for (int r = r0 + warp; r <= r1; r += 8) {
      int x0, x1;
      if (!roi_row_span(b, r, s_margin, &x0, &x1)) continue;     // only the part of the row the rotated ROI reaches
      const int sb = (3 * x0) & ~15;
      int eb = (3 * (x1 + 1) + 15) & ~15;
      if (eb > (int)row_stride) eb = (int)row_stride;
      const int nq = (eb - sb) >> 4;               // 16-byte pieces of this row
      const long long base = (long long)s_frame * frame_stride + sb;
      const uint4* src = reinterpret_cast<const uint4*>(host_frames + base + (long long)r * row_stride);
      if (row_pos) {       // rows the copy engine already gathered for the letterbox are on the device: they do not cross PCIe again
        const int cp = row_pos[r];
        if (cp >= 0) src = reinterpret_cast<const uint4*>(compact + (long long)s_frame * compact_fstride + (long long)cp * row_stride + sb);
      }
      uint4* dst = reinterpret_cast<uint4*>(dev_frames + base + (long long)r * row_stride);
      for (int q = lane; q < nq; q += 32) dst[q] = __ldg(src + q);
    }
  }
}

// Eye slots whose source rectangle lies inside their face's copied rectangle run on the device copy, the others on the host frame:
// the two launches get complementary parameter arrays (valid == 2: "not yours").
__global__ void eye_split_kernel(const I2TParams* __restrict__ eye_params, const SrcBox* __restrict__ face_boxes, int n, const int* n_active,
                                 I2TParams* p_dev, I2TParams* p_host) {
  if (n_active) n = min(n, *n_active);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  I2TParams P = eye_params[i];
  const SrcBox e = warp_src_box(P), f = face_boxes[i >> 1];
  const bool inside = P.valid != 1 ||           // invalid slots just get their zero tensor from the device launch
                      roi_stage_covers(f, e, P.src_w, P.src_h);
  I2TParams Q = P;
  Q.valid = 2;
  p_dev[i] = inside ? P : Q;
  p_host[i] = inside ? Q : P;
}

// ------------------------------------------------------------------------------------------------
__global__ void face_detection_to_roi_kernel(const fdl_detection* det, int img_w, int img_h, int size_mode, fdl_rect* out, int* ok) {
  *ok = face_detection_to_roi(det->data, img_w, img_h, size_mode, out) ? 1 : 0;
}
__global__ void eye_rois_kernel(const double* xy, int img_w, int img_h, fdl_rect* out2, int* ok) {
  bool a = eye_roi(xy[0], xy[1], xy[2], xy[3], img_w, img_h, &out2[0]);
  bool b = eye_roi(xy[4], xy[5], xy[6], xy[7], img_w, img_h, &out2[1]);
  *ok = (a && b) ? 1 : 0;
}
__global__ void project_kernel(const float* raw, int n, int tensor_w, int tensor_h, int img_w, int img_h, const double* pad4,
                               const fdl_rect* roi, int flip, float* out) {
  __shared__ ProjectParams pp;
  if (threadIdx.x == 0) project_setup(tensor_w, tensor_h, img_w, img_h, pad4, roi, flip != 0, &pp);
  __syncthreads();
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) project_point(pp, raw + 3 * k, out + 3 * k);
}

}  // namespace

#define FDL_LAUNCHED() (count_launch(), cudaGetLastError())

cudaError_t launch_anchors(const SsdOptions& opt, float* out, int n, cudaStream_t s) {
  anchors_kernel<<<(n + 255) / 256, 256, 0, s>>>(opt, out, n);
  return FDL_LAUNCHED();
}

cudaError_t launch_i2t_setup(const fdl_rect* rois, const int* slot_frame, const int* slot_valid, int n, int img_w, int img_h,
                             int out_w, int out_h, int keep_aspect, double range_min, double range_max, int flip_mode,
                             I2TParams* params, const int* n_active, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  i2t_setup_kernel<<<(n + 63) / 64, 64, 0, s>>>(rois, slot_frame, slot_valid, n, img_w, img_h, out_w, out_h, keep_aspect, range_min,
                                                range_max, flip_mode, params, n_active);
  return FDL_LAUNCHED();
}

cudaError_t launch_i2t(const uint8_t* frames, long long frame_stride, long long row_stride, const I2TParams* params, int n,
                       int out_w, int out_h, float* out, long long out_bstride, uint8_t* out_u8, const int* n_active,
                       cudaStream_t s, int rows_mode, int src_w, int max_ctas, const uint8_t* compact, const int* row_pos,
                       long long compact_fstride) {
  if (n <= 0) return cudaSuccess;
  // max_ctas > 0 (frames are mapped pinned host memory): a few persistent CTAs keep PCIe busy without occupying the SMs
  // another lane's network kernels need
  if (rows_mode && !out_u8) {
    const size_t row_pad = (size_t)((src_w * 3 + 15) & ~15);
    static const int rpi_env = getenv("FDL_I2T_RPI") ? atoi(getenv("FDL_I2T_RPI")) : 0;   // A/B: output rows per work item
    int rpi = (rpi_env >= 1 && rpi_env <= kRowsPerItem) ? rpi_env : kRowsPerItem;
    while (rpi > 1 && 2 * rpi * row_pad > 48 * 1024) rpi >>= 1;
    const size_t smem = 2 * rpi * row_pad;
    if (smem <= 48 * 1024) {
      long long items = (long long)n * ((out_h + rpi - 1) / rpi);
      const long long per_sm = (long long)((200 * 1024) / (smem + 1024)) < 8 ? (long long)((200 * 1024) / (smem + 1024)) : 8;
      const long long cap = max_ctas > 0 ? max_ctas : 148LL * (per_sm < 4 ? 4 : per_sm);      // persistent CTAs (items are strided over the grid)
      if (items > cap) items = cap;
      static const int minb = getenv("FDL_I2T_MINB") ? atoi(getenv("FDL_I2T_MINB")) : 4;      // A/B: 4 CTAs per SM (64 registers) or 3 (85)
      if (minb >= 4)
        i2t_rows_kernel<4><<<(unsigned)items, 256, smem, s>>>(frames, frame_stride, row_stride, params, n, out_w, out_h, out, out_bstride, n_active,
                                                              compact, row_pos, compact_fstride, rpi);
      else
        i2t_rows_kernel<3><<<(unsigned)items, 256, smem, s>>>(frames, frame_stride, row_stride, params, n, out_w, out_h, out, out_bstride, n_active,
                                                              compact, row_pos, compact_fstride, rpi);
      return FDL_LAUNCHED();
    }
  }
  if (max_ctas > 0 && !out_u8) {
    static std::atomic<unsigned long long> tile_attr{0};
    opt_in_smem_once(tile_attr, i2t_tile_kernel, kTileSmem);
    long long items = (long long)n * ((out_w + kTileW - 1) / kTileW) * ((out_h + kTileH - 1) / kTileH);
    if (items > max_ctas) items = max_ctas;
    i2t_tile_kernel<<<(unsigned)items, 256, kTileSmem, s>>>(frames, frame_stride, row_stride, params, n, out_w, out_h, out, out_bstride, n_active);
    return FDL_LAUNCHED();
  }
  // output pixels per thread and work item: 4 for the 192x192 face warps (0.244 -> 0.208 ms per 256 faces), 1 for the 64x64 eye
  // warps (4 measured 0.163 against 0.152 ms per 512 eyes: too few work items left); FDL_I2T_PX overrides for A/B timing
  static const int ppt_env = getenv("FDL_I2T_PX") ? atoi(getenv("FDL_I2T_PX")) : 0;
  const int ppt_auto = out_w * out_h >= 128 * 128 ? 4 : 1;
  const int ppt = ppt_env >= 4 ? 4 : (ppt_env >= 2 ? 2 : (ppt_env == 1 ? 1 : ppt_auto));
  long long grid = (long long)n * ((out_w * out_h + 256 * ppt - 1) / (256 * ppt));
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (ppt == 4) i2t_kernel<4><<<(unsigned)grid, 256, 0, s>>>(frames, frame_stride, row_stride, params, n, out_w, out_h, out, out_bstride, out_u8, n_active);
  else if (ppt == 2) i2t_kernel<2><<<(unsigned)grid, 256, 0, s>>>(frames, frame_stride, row_stride, params, n, out_w, out_h, out, out_bstride, out_u8, n_active);
  else i2t_kernel<1><<<(unsigned)grid, 256, 0, s>>>(frames, frame_stride, row_stride, params, n, out_w, out_h, out, out_bstride, out_u8, n_active);
  return FDL_LAUNCHED();
}

cudaError_t launch_ssd_postprocess(const SsdPostArgs& a, cudaStream_t s) {
  if (a.B <= 0) return cudaSuccess;
  size_t smem = (size_t)a.N * 5 * sizeof(int);
  static std::atomic<unsigned long long> post_attr{0};
  opt_in_smem_once(post_attr, ssd_postprocess_kernel, 160 * 1024);
  ssd_postprocess_kernel<<<a.B, kPostThreads, smem, s>>>(a);
  return FDL_LAUNCHED();
}

cudaError_t launch_face_select(fdl_frame_result* frames, int B, int max_faces, int* slot_frame, int* slot_face, int* n_faces,
                               int* n_eyes, cudaStream_t s) {
  face_select_kernel<<<1, 1024, 0, s>>>(frames, B, max_faces, slot_frame, slot_face, n_faces, n_eyes);
  return FDL_LAUNCHED();
}

cudaError_t launch_face_roi(const fdl_frame_result* frames, const int* slot_frame, const int* slot_face, int max_slots, int max_faces,
                            int img_w, int img_h, fdl_rect* rois, int* slot_valid, fdl_face_result* faces, const int* n_faces,
                            cudaStream_t s) {
  if (max_slots <= 0) return cudaSuccess;
  face_roi_kernel<<<(max_slots + 63) / 64, 64, 0, s>>>(frames, slot_frame, slot_face, max_slots, max_faces, img_w, img_h, rois,
                                                       slot_valid, faces, n_faces);
  return FDL_LAUNCHED();
}

cudaError_t launch_landmark_post(const float* raw, long long raw_bstride, const float* flag, long long flag_bstride,
                                 const I2TParams* params, const fdl_rect* rois, const int* slot_frame, const int* slot_face,
                                 int max_slots, int max_faces, int tensor_w, int tensor_h, fdl_face_result* faces, fdl_rect* eye_rois,
                                 int* eye_frame, int* eye_valid, const int* n_faces, cudaStream_t s, int refine) {
  if (max_slots <= 0) return cudaSuccess;
  landmark_post_kernel<<<max_slots, 128, 0, s>>>(raw, raw_bstride, flag, flag_bstride, params, rois, slot_frame, slot_face, max_slots,
                                                 max_faces, tensor_w, tensor_h, faces, eye_rois, eye_frame, eye_valid, n_faces, refine);
  return FDL_LAUNCHED();
}

cudaError_t launch_iris_post(const float* contour, long long contour_bstride, const float* iris, long long iris_bstride,
                             const I2TParams* params, const fdl_rect* eye_rois, const int* eye_valid, const int* slot_frame,
                             const int* slot_face, int max_eye_slots, int max_faces, int tensor_w, int tensor_h,
                             fdl_face_result* faces, const int* n_eyes, cudaStream_t s, int refine, double focal_length_mm) {
  if (max_eye_slots <= 0) return cudaSuccess;
  iris_post_kernel<<<max_eye_slots, 96, 0, s>>>(contour, contour_bstride, iris, iris_bstride, params, eye_rois, eye_valid, slot_frame,
                                                slot_face, max_eye_slots, max_faces, tensor_w, tensor_h, faces, n_eyes, refine, focal_length_mm);
  return FDL_LAUNCHED();
}
cudaError_t launch_refine_landmarks(const double* face, const double* left, int n_left, const double* right, int n_right, double* out,
                                    cudaStream_t s) {
  refine_landmarks_kernel<<<1, 256, 0, s>>>(face, left, n_left, right, n_right, out);
  return FDL_LAUNCHED();
}
cudaError_t launch_iris_metrics(const double* iris, int img_w, int img_h, double focal_length_mm, double iris_size_px, double* out2,
                                cudaStream_t s) {
  iris_metrics_kernel<<<1, 1, 0, s>>>(iris, img_w, img_h, focal_length_mm, iris_size_px, out2);
  return FDL_LAUNCHED();
}

cudaError_t launch_roi_fill(const uint8_t* host_frames, uint8_t* dev_frames, long long frame_stride, long long row_stride, const I2TParams* params,
                            int n, const int* n_active, SrcBox* boxes, int frame_h, int max_ctas, cudaStream_t s, const uint8_t* compact,
                            const int* row_pos, long long compact_fstride) {
  if (n <= 0) return cudaSuccess;
  long long items = (long long)n * ((frame_h + kFillRows - 1) / kFillRows);
  if (max_ctas > 0 && items > max_ctas) items = max_ctas;
  static const int margin_env = getenv("FDL_ZC_MARGIN") ? atoi(getenv("FDL_ZC_MARGIN")) : 0;
  static const int trim_env = getenv("FDL_ZC_TRIM") ? atoi(getenv("FDL_ZC_TRIM")) : 1;     // copy only the rotated ROI's span of every row
  roi_fill_kernel<<<(unsigned)items, 256, 0, s>>>(host_frames, dev_frames, frame_stride, row_stride, params, n, n_active, boxes, frame_h, margin_env,
                                                  trim_env, compact, row_pos, compact_fstride);
  return FDL_LAUNCHED();
}
cudaError_t launch_eye_split(const I2TParams* eye_params, const SrcBox* face_boxes, int n, const int* n_active, I2TParams* p_dev, I2TParams* p_host,
                             cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  eye_split_kernel<<<(n + 127) / 128, 128, 0, s>>>(eye_params, face_boxes, n, n_active, p_dev, p_host);
  return FDL_LAUNCHED();
}

cudaError_t launch_face_detection_to_roi(const fdl_detection* det, int img_w, int img_h, int size_mode, fdl_rect* out, int* ok,
                                         cudaStream_t s) {
  face_detection_to_roi_kernel<<<1, 1, 0, s>>>(det, img_w, img_h, size_mode, out, ok);
  return FDL_LAUNCHED();
}
cudaError_t launch_eye_rois(const double* lm4xy, int img_w, int img_h, fdl_rect* out2, int* ok, cudaStream_t s) {
  eye_rois_kernel<<<1, 1, 0, s>>>(lm4xy, img_w, img_h, out2, ok);
  return FDL_LAUNCHED();
}
cudaError_t launch_project(const float* raw, int n, int tensor_w, int tensor_h, int img_w, int img_h, const double* pad4,
                           const fdl_rect* roi_or_null, int flip, float* out, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  project_kernel<<<1, 128, 0, s>>>(raw, n, tensor_w, tensor_h, img_w, img_h, pad4, roi_or_null, flip, out);
  return FDL_LAUNCHED();
}

}  // namespace fdl
