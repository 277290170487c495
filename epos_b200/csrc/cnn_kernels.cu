// HBM-bound CNN kernels of the EPOS forward pass (everything except the tensor-core pointwise GEMM):
// entry convs, depthwise 3x3 (strided / atrous), f32->split-bf16, global mean, bilinear resize,
// softmax/argmax, plus an fp32 SIMT GEMM used for validation and for tiny (M = batch) GEMMs.
// Reference call sites are cited in include/epos_b200.h next to each entry point.
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"

namespace epos {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// conv1_1: preprocessing + 3x3 s2 conv (pad 1/1, VALID) + bias + ReLU.  One thread per output pixel and
// 32-channel half: the 27 input values are loaded once (not once per 4-channel group), the filter bank is
// read from shared memory as warp-wide broadcasts, and a thread writes 128 (f32) / 64 + 64 (split) contiguous bytes.
// ------------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(128) conv3x3_rgb_s2_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ y,
                                                             uint16_t* __restrict__ y_split, long long plane_stride,
                                                             int B, int H, int W, int Ho, int Wo) {
  __shared__ __align__(16) float sw[27 * COUT];
  __shared__ __align__(16) float sb[COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  constexpr int HALVES = COUT / 32;
  const long long total = (long long)B * Ho * Wo * HALVES;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    // the half index is the slowest so that a warp covers 32 consecutive pixels of one half
    long long p = idx % ((long long)B * Ho * Wo);
    const int c0 = (int)(idx / ((long long)B * Ho * Wo)) * 32;
    const long long pix = p;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    float v[27];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        const bool in = iy >= 0 && iy < H && ix >= 0 && ix < W;
        const float* px = x + (((long long)b * H + (in ? iy : 0)) * W + (in ? ix : 0)) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c)                                  // feature.py:171-174; zero padding AFTER the scaling
          v[(ky * 3 + kx) * 3 + c] = in ? (2.0f / 255.0f) * __ldg(px + c) - 1.0f : 0.f;
      }
    }
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = sb[c0 + j];
    // same accumulation order as before (taps in ky, kx, c order), so results are unchanged bit for bit
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float4* ww = reinterpret_cast<const float4*>(sw + t * COUT + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 q = ww[j];
        acc[4 * j + 0] = fmaf(v[t], q.x, acc[4 * j + 0]);
        acc[4 * j + 1] = fmaf(v[t], q.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(v[t], q.z, acc[4 * j + 2]);
        acc[4 * j + 3] = fmaf(v[t], q.w, acc[4 * j + 3]);
      }
    }
    const long long off = pix * COUT + c0;
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = fmaxf(acc[j], 0.f);
    if (y) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(y + off + 4 * j) = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    }
    if (y_split) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {                                  // 8 channels = 16 bytes per plane and store
        uint4 hi, lo;
        split_bf16x2(acc[8 * j + 0], acc[8 * j + 1], hi.x, lo.x);
        split_bf16x2(acc[8 * j + 2], acc[8 * j + 3], hi.y, lo.y);
        split_bf16x2(acc[8 * j + 4], acc[8 * j + 5], hi.z, lo.z);
        split_bf16x2(acc[8 * j + 6], acc[8 * j + 7], hi.w, lo.w);
        *reinterpret_cast<uint4*>(y_split + off + 8 * j) = hi;
        *reinterpret_cast<uint4*>(y_split + plane_stride + off + 8 * j) = lo;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// conv1_2: dense 3x3 s1 SAME + bias + ReLU.  One CTA = 16x16 output pixels x COUT channels; the
// 18x18xCIN halo tile and the whole filter bank live in shared memory.
// ------------------------------------------------------------------------------------------------
template <int CIN, int COUT>
__global__ void __launch_bounds__(256) conv3x3_dense_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ y,
                                                            int H, int W) {
  extern __shared__ float smem[];
  constexpr int PS = CIN + 1;                      // padded pixel stride: conflict-free across tx
  float* s_in = smem;                              // [18*18][PS]
  float* s_w = smem + 18 * 18 * PS;                // [9*CIN][COUT]
  const int b = blockIdx.z;
  const int ty0 = blockIdx.y * 16, tx0 = blockIdx.x * 16;
  for (int i = threadIdx.x; i < 9 * CIN * COUT / 4; i += 256)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
  for (int i = threadIdx.x; i < 18 * 18 * (CIN / 4); i += 256) {
    const int c4 = i % (CIN / 4);
    const int p = i / (CIN / 4);
    const int iy = ty0 - 1 + p / 18, ix = tx0 - 1 + p % 18;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = __ldg(reinterpret_cast<const float4*>(x + (((long long)b * H + iy) * W + ix) * CIN) + c4);
    float* d = s_in + p * PS + c4 * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  float acc[COUT];
#pragma unroll
  for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      const float* pin = s_in + ((ty + ky) * 18 + tx + kx) * PS;
      const float* pw = s_w + (ky * 3 + kx) * CIN * COUT;
#pragma unroll 4
      for (int c = 0; c < CIN; ++c) {
        const float a = pin[c];
        const float4* w4 = reinterpret_cast<const float4*>(pw + c * COUT);
#pragma unroll
        for (int j = 0; j < COUT / 4; ++j) {
          const float4 ww = w4[j];
          acc[4 * j + 0] = fmaf(a, ww.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(a, ww.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(a, ww.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(a, ww.w, acc[4 * j + 3]);
        }
      }
    }
  const int oy = ty0 + ty, ox = tx0 + tx;
  if (oy < H && ox < W) {
    float4* out = reinterpret_cast<float4*>(y + (((long long)b * H + oy) * W + ox) * COUT);
#pragma unroll
    for (int j = 0; j < COUT / 4; ++j) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + j);
      out[j] = make_float4(fmaxf(acc[4 * j + 0] + bb.x, 0.f), fmaxf(acc[4 * j + 1] + bb.y, 0.f),
                           fmaxf(acc[4 * j + 2] + bb.z, 0.f), fmaxf(acc[4 * j + 3] + bb.w, 0.f));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Depthwise 3x3.  A warp owns a strip of DW_PX consecutive output pixels of one row and 32 channel groups
// (lane = 4 channels): every tap is one coalesced 512-byte row segment, the 9 filter taps of the lane's channels
// live in registers, bounds tests are warp-uniform, DW_PX independent accumulators give the loads ILP.
// Output as f32 and/or split-bf16.
// ------------------------------------------------------------------------------------------------
constexpr int DW_PX = 4;

template <bool INTERIOR>
__device__ __forceinline__ void dw_strip(const float4* __restrict__ x0, int ld4, int W, int H, int iy0, int ix0, int stride,
                                         int rate, int npx, bool relu_in, const float4 (&wk)[9], float4 (&acc)[DW_PX]) {
  // x0 points at (row iy0, column ix0) of this image for the lane's channel group; taps are addressed with 32-bit
  // element offsets relative to it.
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    if (!INTERIOR) {
      const int iy = iy0 + ky * rate;
      if (iy < 0 || iy >= H) continue;
    }
    const int rowoff = ky * rate * W * ld4;
#pragma unroll
    for (int p = 0; p < DW_PX; ++p) {
      if (!INTERIOR && p >= npx) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int dx = p * stride + kx * rate;
        if (!INTERIOR) {
          const int ix = ix0 + dx;
          if (ix < 0 || ix >= W) continue;
        }
        float4 v = __ldg(x0 + (rowoff + dx * ld4));
        if (relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        const float4 ww = wk[ky * 3 + kx];
        acc[p].x = fmaf(v.x, ww.x, acc[p].x); acc[p].y = fmaf(v.y, ww.y, acc[p].y);
        acc[p].z = fmaf(v.z, ww.z, acc[p].z); acc[p].w = fmaf(v.w, ww.w, acc[p].w);
      }
    }
  }
}

__global__ void __launch_bounds__(256, 3) dwconv3x3_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ y_f32,
                                                           uint16_t* __restrict__ y_split, int ldy_split, long long plane_stride,
                                                           int B, int H, int W, int C, int Ho, int Wo, int stride, int rate,
                                                           int relu_in, int relu_out) {
  const int G = C >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.y * 32 + lane;
  // a block = 8 consecutive output ROWS of one strip column: the rows a warp reads are re-read by its neighbours in the
  // same block (from L1) when rate <= 3
  // (measured: pays for rate <= 2); for larger rates a block = 8 consecutive strips of one row
  const int strips_per_row = (Wo + DW_PX - 1) / DW_PX;
  int sx, oy, b;
  if (rate <= 2) {
    const int row_blocks = (Ho + 7) >> 3;
    sx = blockIdx.x % strips_per_row;
    const int q = blockIdx.x / strips_per_row;
    oy = (q % row_blocks) * 8 + warp; b = q / row_blocks;
  } else {
    const int strip = blockIdx.x * 8 + warp;                    // strips linearised over (b, oy, sx)
    sx = strip % strips_per_row;
    const int q = strip / strips_per_row;
    oy = q % Ho; b = q / Ho;
  }
  if (b >= B || oy >= Ho || g >= G) return;
  const int ox0 = sx * DW_PX;
  const int npx = min(DW_PX, Wo - ox0);
  float4 wk[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wk[t] = __ldg(reinterpret_cast<const float4*>(w + (size_t)t * C) + g);
  const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + g);
  float4 acc[DW_PX];
#pragma unroll
  for (int p = 0; p < DW_PX; ++p) acc[p] = bb;
  const int ld4 = ldx >> 2;
  const int iy0 = oy * stride - rate, ix0 = ox0 * stride - rate;
  const float4* x0 = reinterpret_cast<const float4*>(x) + (((long long)b * H + iy0) * W + ix0) * ld4 + g;
  const bool interior = iy0 >= 0 && iy0 + 2 * rate < H && ix0 >= 0 && ix0 + (DW_PX - 1) * stride + 2 * rate < W &&
                        npx == DW_PX;
  if (interior) dw_strip<true>(x0, ld4, W, H, iy0, ix0, stride, rate, npx, relu_in != 0, wk, acc);
  else dw_strip<false>(x0, ld4, W, H, iy0, ix0, stride, rate, npx, relu_in != 0, wk, acc);
  const long long pix0 = ((long long)b * Ho + oy) * Wo + ox0;
#pragma unroll
  for (int p = 0; p < DW_PX; ++p) {
    if (p >= npx) break;
    float4 a = acc[p];
    if (relu_out) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
    const long long pix = pix0 + p;
    if (y_f32) *(reinterpret_cast<float4*>(y_f32 + pix * C) + g) = a;
    if (y_split) {
      uint2 hi, lo;
      split_bf16x2(a.x, a.y, hi.x, lo.x);
      split_bf16x2(a.z, a.w, hi.y, lo.y);
      *(reinterpret_cast<uint2*>(y_split + pix * ldy_split) + g) = hi;
      *(reinterpret_cast<uint2*>(y_split + plane_stride + pix * ldy_split) + g) = lo;
    }
  }
}

// Input side of the inference path (datagen.py:424-476, _parse_and_preprocess): the decoded uint8 image is resized to
// (new_w, new_h) with misc.resize_image_tf (misc.py:75-91) -- tf.image.resize_area(align_corners=True) when the image is
// not enlarged (in_h >= new_h), tf.image.resize_bilinear(align_corners=True) otherwise -- then cropped at
// (off_y, off_x) to crop_h x crop_w, as float32 in [0, 255].  One thread per output pixel, the three channels together.
// resize_area (TensorFlow 1.12 core/kernels/resize_area_op.cc, restated): scale = (in - 1) / (out - 1) with
// align_corners (in / out when out == 1); output pixel y averages the input span [y scale, (y + 1) scale): source row i
// has weight (i < y scale ? (i + 1 > (y+1) scale ? scale : i + 1 - y scale) : (i + 1 > (y+1) scale ? (y+1) scale - i : 1)),
// rows beyond the image are clamped to the last row; the sum is divided by scale_y scale_x.
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ src, int in_h, int in_w,
                                                            long long src_pitch, float* __restrict__ dst, int new_h,
                                                            int new_w, int off_y, int off_x, int crop_h, int crop_w,
                                                            int area) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
  if (ox >= crop_w || oy >= crop_h) return;
  const int y = oy + off_y, x = ox + off_x;                    // position in the resized image
  float r = 0.f, g = 0.f, b = 0.f;
  if (y >= 0 && y < new_h && x >= 0 && x < new_w) {
    const float sy = (new_h > 1) ? (float)(in_h - 1) / (float)(new_h - 1) : (float)in_h / (float)new_h;
    const float sx = (new_w > 1) ? (float)(in_w - 1) / (float)(new_w - 1) : (float)in_w / (float)new_w;
    if (!area) {
      const float fy = y * sy, fx = x * sx;
      const int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
      const int y1 = min(y0 + 1, in_h - 1), x1 = min(x0 + 1, in_w - 1);
      const float ly = fy - y0, lx = fx - x0;
      const uint8_t* p00 = src + y0 * src_pitch + 3 * x0; const uint8_t* p01 = src + y0 * src_pitch + 3 * x1;
      const uint8_t* p10 = src + y1 * src_pitch + 3 * x0; const uint8_t* p11 = src + y1 * src_pitch + 3 * x1;
      float* o[3] = {&r, &g, &b};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float top = (float)p00[c] + ((float)p01[c] - (float)p00[c]) * lx;
        const float bot = (float)p10[c] + ((float)p11[c] - (float)p10[c]) * lx;
        *o[c] = top + (bot - top) * ly;
      }
    } else {
      const float in_y = y * sy, in_y1 = (y + 1) * sy, in_x = x * sx, in_x1 = (x + 1) * sx;
      const int ys = (int)floorf(in_y), ye = (int)ceilf(in_y1), xs = (int)floorf(in_x), xe = (int)ceilf(in_x1);
      const float norm = 1.0f / (sy * sx);
      for (int i = ys; i < ye; ++i) {
        const float wy = i < in_y ? (i + 1 > in_y1 ? sy : i + 1 - in_y) : (i + 1 > in_y1 ? in_y1 - i : 1.0f);
        const uint8_t* row = src + (long long)min(max(i, 0), in_h - 1) * src_pitch;
        for (int j = xs; j < xe; ++j) {
          const float wx = j < in_x ? (j + 1 > in_x1 ? sx : j + 1 - in_x) : (j + 1 > in_x1 ? in_x1 - j : 1.0f);
          const uint8_t* px = row + 3 * min(max(j, 0), in_w - 1);
          const float wgt = wy * wx * norm;
          r += (float)px[0] * wgt; g += (float)px[1] * wgt; b += (float)px[2] * wgt;
        }
      }
    }
  }
  float* o = dst + ((long long)oy * crop_w + ox) * 3;
  o[0] = r; o[1] = g; o[2] = b;
}

// slim.max_pool2d(3, stride 2, 'SAME') (net_resnet_v1_beta.py:187): TF pads total = max((ceil(n/2)-1)*2+3-n, 0), before =
// floor(total/2) (0/1 for even n, 1/1 for odd n); padded taps never win the max.
__global__ void __launch_bounds__(256) maxpool3x3_s2_kernel(const float* __restrict__ x, float* __restrict__ y_f32,
                                                            uint16_t* __restrict__ y_split, long long plane_stride, int B,
                                                            int H, int W, int C, int Ho, int Wo) {
  const int G = C >> 2;
  const int pt = max((Ho - 1) * 2 + 3 - H, 0) / 2, pl = max((Wo - 1) * 2 + 3 - W, 0) / 2;
  const long long total = (long long)B * Ho * Wo * G;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    long long p = idx / G;
    const int ox = (int)(p % Wo);
    long long q = p / Wo;
    const int oy = (int)(q % Ho);
    const int b = (int)(q / Ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - pt + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - pl + kx;
        if (ix < 0 || ix >= W) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((long long)b * H + iy) * W + ix) * C) + g);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    if (y_f32) *(reinterpret_cast<float4*>(y_f32 + p * C) + g) = m;
    if (y_split) {
      uint2 hi, lo;
      split_bf16x2(m.x, m.y, hi.x, lo.x);
      split_bf16x2(m.z, m.w, hi.y, lo.y);
      *(reinterpret_cast<uint2*>(y_split + p * C) + g) = hi;
      *(reinterpret_cast<uint2*>(y_split + plane_stride + p * C) + g) = lo;
    }
  }
}

// resnet_utils.subsample (external/slim/nets/resnet_utils.py:59-74): every factor-th pixel, f32 -> f32.
__global__ void __launch_bounds__(256) subsample_f32_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y,
                                                            int B, int H, int W, int C, int Ho, int Wo, int factor) {
  const int G = C >> 2;
  const long long total = (long long)B * Ho * Wo * G;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    long long p = idx / G;
    const int ox = (int)(p % Wo);
    long long q = p / Wo;
    const int oy = (int)(q % Ho);
    const int b = (int)(q / Ho);
    *(reinterpret_cast<float4*>(y + p * C) + g) =
        __ldg(reinterpret_cast<const float4*>(x + (((long long)b * H + oy * factor) * W + ox * factor) * ldx) + g);
  }
}

__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, int ldx, uint16_t* __restrict__ y,
                                                         int ldy, long long plane_stride, int B, int H, int W, int C,
                                                         int Ho, int Wo, int sub, int relu) {
  const int G = C >> 2;
  const long long total = (long long)B * Ho * Wo * G;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    long long p = idx / G;
    const int ox = (int)(p % Wo);
    long long q = p / Wo;
    const int oy = (int)(q % Ho);
    const int b = (int)(q / Ho);
    float4 v = __ldg(reinterpret_cast<const float4*>(x + (((long long)b * H + oy * sub) * W + ox * sub) * ldx) + g);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
    split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
    *(reinterpret_cast<uint2*>(y + p * ldy) + g) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
    *(reinterpret_cast<uint2*>(y + plane_stride + p * ldy) + g) = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
  }
}

// x [B][HW][C] -> y [B][C].  Block = 32 channels x 8 row lanes.
__global__ void __launch_bounds__(256) global_mean_kernel(const float* __restrict__ x, float* __restrict__ y, int HW, int C) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int b = blockIdx.y;
  float s = 0.f;
  if (c < C)
    for (int r = threadIdx.y; r < HW; r += 8) s += __ldg(x + ((long long)b * HW + r) * C + c);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    y[(long long)b * C + c] = t / (float)HW;
  }
}

// tf.image.resize_bilinear(align_corners=True): src = dst * (in-1)/(out-1), lerp in f32.
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float* __restrict__ x, float* __restrict__ y, int ldy,
                                                              int B, int Hi, int Wi, int Ho, int Wo, int C) {
  const int G = C >> 2;
  const float sy = (Ho > 1) ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  const float sx = (Wo > 1) ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  const long long total = (long long)B * Ho * Wo * G;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    long long p = idx / G;
    const int ox = (int)(p % Wo);
    long long q = p / Wo;
    const int oy = (int)(q % Ho);
    const int b = (int)(q / Ho);
    const float fy = oy * sy, fx = ox * sx;
    const int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
    const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
    const float ly = fy - y0, lx = fx - x0;
    const float4* base = reinterpret_cast<const float4*>(x + (long long)b * Hi * Wi * C) + g;
    const float4 tl = __ldg(base + ((long long)y0 * Wi + x0) * G), tr = __ldg(base + ((long long)y0 * Wi + x1) * G);
    const float4 bl = __ldg(base + ((long long)y1 * Wi + x0) * G), br = __ldg(base + ((long long)y1 * Wi + x1) * G);
    float4 o;
#define EPOS_LERP(f)                                   \
  {                                                    \
    const float top = tl.f + (tr.f - tl.f) * lx;       \
    const float bot = bl.f + (br.f - bl.f) * lx;       \
    o.f = top + (bot - top) * ly;                      \
  }
    EPOS_LERP(x) EPOS_LERP(y) EPOS_LERP(z) EPOS_LERP(w)
#undef EPOS_LERP
    *(reinterpret_cast<float4*>(y + p * ldy) + g) = o;
  }
}

// Softmax over rows of length n, in place; one warp per row.  Optional argmax (first maximum).
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ x, int64_t* __restrict__ labels,
                                                           long long rows, int n) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    float* row = x + r * n;
    float m = -INFINITY;
    for (int i = lane; i < n; i += 32) m = fmaxf(m, row[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s += expf(row[i] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float best = -1.f;
    int besti = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
      const float v = expf(row[i] - m) / s;
      row[i] = v;
      if (v > best) { best = v; besti = i; }
    }
    if (labels) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
      }
      if (lane == 0) labels[r] = besti;
    }
  }
}

// The same softmax over the F fragment confidences of a (pixel, object) pair, only for the pairs the correspondence
// extraction will read: those whose object confidence exceeds its threshold (corresp.py:46-60).  Other rows keep their
// logits.  Engine path only (model.predict materialises the whole tensor); identical arithmetic per row.
__global__ void __launch_bounds__(256) softmax_rows_masked_kernel(float* __restrict__ x, const float* __restrict__ obj_conf,
                                                                  long long pixels, int O, int F, float min_obj_conf) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long rows = pixels * O;
  for (long long r = warp; r < rows; r += nwarps) {
    const long long p = r / O;
    const int o = (int)(r - p * O);
    if (!(__ldg(obj_conf + p * (O + 1) + o + 1) > min_obj_conf)) continue;
    float* row = x + r * F;
    float m = -INFINITY;
    for (int i = lane; i < F; i += 32) m = fmaxf(m, row[i]);
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o2));
    float s = 0.f;
    for (int i = lane; i < F; i += 32) s += expf(row[i] - m);
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
    for (int i = lane; i < F; i += 32) row[i] = expf(row[i] - m) / s;
  }
}

// Tiny-M fp32 GEMM (M <= 16: the image-pooling branch has M = batch): one warp per output column, lanes stride over K
// (coalesced W rows), the M accumulators live in registers and are reduced with shuffles.
constexpr int SMALLM_MAX = 16;
__global__ void __launch_bounds__(256) pwconv_smallm_kernel(const float* __restrict__ a, int lda, const float* __restrict__ w,
                                                            const float* __restrict__ bias, int bias_group_rows,
                                                            float* __restrict__ d, int ldd, int M, int N, int K, int relu) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  float acc[SMALLM_MAX];
#pragma unroll
  for (int m = 0; m < SMALLM_MAX; ++m) acc[m] = 0.f;
  const float* wr = w + (long long)n * K;
  for (int k = lane; k < K; k += 32) {
    const float wv = __ldg(wr + k);
#pragma unroll
    for (int m = 0; m < SMALLM_MAX; ++m)
      if (m < M) acc[m] = fmaf(__ldg(a + (long long)m * lda + k), wv, acc[m]);
  }
#pragma unroll
  for (int m = 0; m < SMALLM_MAX; ++m) {
    float v = acc[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && m < M) {
      if (bias) v += bias[(bias_group_rows > 0 ? (long long)(m / bias_group_rows) * N : 0) + n];
      if (relu) v = fmaxf(v, 0.f);
      d[(long long)m * ldd + n] = v;
    }
  }
}

// fp32 SIMT GEMM: D = act(A W^T + bias) (+ residual).  64x64 tile, 256 threads, 4x4 per thread.
__global__ void __launch_bounds__(256) pwconv_simt_kernel(const float* __restrict__ a, int lda, const float* __restrict__ w,
                                                          const float* __restrict__ bias, int bias_group_rows,
                                                          const float* __restrict__ residual, int ldr, float* __restrict__ d,
                                                          int ldd, int M, int N, int K, int relu) {
  __shared__ float sa[16][64 + 4];
  __shared__ float sb[16][64 + 4];
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i / 16, k = i % 16;
      sa[k][r] = (m0 + r < M && k0 + k < K) ? a[(long long)(m0 + r) * lda + k0 + k] : 0.f;
      sb[k][r] = (n0 + r < N && k0 + k < K) ? w[(long long)(n0 + r) * K + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = sa[k][ty * 4 + i]; bv[i] = sb[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const float* brow = bias ? bias + (bias_group_rows > 0 ? (long long)(m / bias_group_rows) * N : 0) : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (brow ? brow[n] : 0.f);
      if (residual) v += residual[(long long)m * ldr + n];
      if (relu) v = fmaxf(v, 0.f);                       // after the residual add, as in the tcgen05 kernel
      d[(long long)m * ldd + n] = v;
    }
  }
}

static inline int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;     // 16 resident CTAs of 256 threads x 148 SMs; grid-stride beyond
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace epos

using namespace epos;

extern "C" {

const char* epos_last_error(void) { return g_err; }
int epos_version(void) { return 1; }
int epos_compiled_arch(void) { return 100; }
uint64_t epos_launch_count(void) { return (uint64_t)g_launches.load(); }

int epos_conv3x3_rgb_s2(const float* x, const float* w, const float* bias, float* y, uint16_t* y_split, int B, int H,
                        int W, int Cout, void* stream) {
  EPOS_CHECK_ARG(x && w && bias && (y || y_split) && B > 0 && H > 0 && W > 0);
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = (long long)B * Ho * Wo * (Cout / 32);
  const long long plane = (long long)B * Ho * Wo * Cout;
  if (Cout == 32) {
    conv3x3_rgb_s2_kernel<32><<<grid_for(total, 128), 128, 0, (cudaStream_t)stream>>>(x, w, bias, y, y_split, plane, B, H, W, Ho, Wo);
  } else if (Cout == 64) {
    conv3x3_rgb_s2_kernel<64><<<grid_for(total, 128), 128, 0, (cudaStream_t)stream>>>(x, w, bias, y, y_split, plane, B, H, W, Ho, Wo);
  } else {
    set_error("epos_conv3x3_rgb_s2: Cout=%d unsupported (32, 64)", Cout);
    return EPOS_ERR_UNSUPPORTED;
  }
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

int epos_maxpool3x3_s2(const float* x, float* y_f32, uint16_t* y_split, int B, int H, int W, int C, void* stream) {
  EPOS_CHECK_ARG(x && (y_f32 || y_split) && B > 0 && H > 0 && W > 0 && C > 0 && (C % 4) == 0);
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long long total = (long long)B * Ho * Wo * (C / 4);
  maxpool3x3_s2_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, y_f32, y_split, (long long)B * Ho * Wo * C, B,
                                                                         H, W, C, Ho, Wo);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

int epos_subsample_f32(const float* x, int ldx, float* y, int B, int H, int W, int C, int factor, void* stream) {
  EPOS_CHECK_ARG(x && y && B > 0 && H > 0 && W > 0 && C > 0 && (C % 4) == 0 && (ldx % 4) == 0 && factor >= 1);
  const int Ho = (H - 1) / factor + 1, Wo = (W - 1) / factor + 1;
  const long long total = (long long)B * Ho * Wo * (C / 4);
  subsample_f32_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, B, H, W, C, Ho, Wo, factor);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

int epos_conv3x3_dense(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin,
                       int Cout, void* stream) {
  EPOS_CHECK_ARG(x && w && bias && y && B > 0 && H > 0 && W > 0);
  if (Cin != 32 || Cout != 64) {
    set_error("epos_conv3x3_dense: (Cin,Cout)=(%d,%d) unsupported (32,64)", Cin, Cout);
    return EPOS_ERR_UNSUPPORTED;
  }
  constexpr int smem = (18 * 18 * 33 + 9 * 32 * 64) * 4;
  static std::atomic<int> attr_set[EPOS_MAX_DEVICES];
  const int dslot = device_slot();
  if (!attr_set[dslot].load(std::memory_order_acquire)) {
    EPOS_CUDA(cudaFuncSetAttribute(conv3x3_dense_kernel<32, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[dslot].store(1, std::memory_order_release);
  }
  dim3 grid(ceil_div(W, 16), ceil_div(H, 16), B);
  conv3x3_dense_kernel<32, 64><<<grid, 256, smem, (cudaStream_t)stream>>>(x, w, bias, y, H, W);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

int epos_dwconv3x3(const float* x, int ldx, const float* w, const float* bias, float* y_f32, uint16_t* y_split,
                   int ldy_split, int B, int H, int W, int C, int stride, int rate, int relu_in, int relu_out, void* stream) {
  EPOS_CHECK_ARG(x && w && bias && (y_f32 || y_split));
  EPOS_CHECK_ARG(!y_split || (ldy_split >= C && (ldy_split % 4) == 0));
  EPOS_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && (C % 4) == 0 && (ldx % 4) == 0 && ldx >= C);
  EPOS_CHECK_ARG((stride == 1 || stride == 2) && rate >= 1);
  const int Ho = stride == 1 ? H : (H - 1) / 2 + 1, Wo = stride == 1 ? W : (W - 1) / 2 + 1;
  const int spr = (Wo + DW_PX - 1) / DW_PX;
  const long long blocks = rate <= 2 ? (long long)B * ((Ho + 7) / 8) * spr : ((long long)B * Ho * spr + 7) / 8;
  EPOS_CHECK_ARG(blocks < (1LL << 31) && ceil_div(C / 4, 32) <= 65535 && (long long)H * W * (ldx / 4) < (1LL << 30));
  dim3 grid((unsigned)blocks, (unsigned)ceil_div(C / 4, 32));
  static int variant = -1;
  if (variant < 0) { const char* e = getenv("EPOS_DW_VARIANT"); variant = e ? atoi(e) : 3; }
  if (variant == 3 && stride == 1) {
    // TMA-staged smem-tiled kernel (dw_tile.cu) for rate 1/2/4; other shapes use the register-strip kernel below
    const int rc = dwconv3x3_tiled(x, ldx, w, bias, y_f32, y_split, ldy_split, B, H, W, C, rate, relu_in, relu_out, (cudaStream_t)stream);
    if (rc != EPOS_ERR_UNSUPPORTED) return rc;
  }
  dwconv3x3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, w, bias, y_f32, y_split, ldy_split,
                                                           (long long)B * Ho * Wo * ldy_split, B, H, W, C, Ho, Wo,
                                                           stride, rate, relu_in, relu_out);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

int epos_split_bf16(const float* x, int ldx, uint16_t* y_split, int ldy, size_t y_plane_stride, int B, int H, int W,
                    int C, int subsample, int relu, void* stream) {
  EPOS_CHECK_ARG(x && y_split && B > 0 && H > 0 && W > 0 && C > 0 && (C % 4) == 0 && (ldx % 4) == 0 && (ldy % 4) == 0);
  EPOS_CHECK_ARG(subsample == 1 || subsample == 2);
  const int Ho = (H - 1) / subsample + 1, Wo = (W - 1) / subsample + 1;
  const long long total = (long long)B * Ho * Wo * (C / 4);
  split_bf16_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, ldx, y_split, ldy, (long long)y_plane_stride,
                                                                       B, H, W, C, Ho, Wo, subsample, relu);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

int epos_global_mean(const float* x, float* y, int B, int HW, int C, void* stream) {
  EPOS_CHECK_ARG(x && y && B > 0 && HW > 0 && C > 0);
  dim3 grid(ceil_div(C, 32), B), block(32, 8);
  global_mean_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, y, HW, C);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

int epos_resize_bilinear(const float* x, float* y, int ldy, int B, int Hi, int Wi, int Ho, int Wo, int C, void* stream) {
  EPOS_CHECK_ARG(x && y && B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C > 0 && (C % 4) == 0 && (ldy % 4) == 0);
  const long long total = (long long)B * Ho * Wo * (C / 4);
  resize_bilinear_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, y, ldy, B, Hi, Wi, Ho, Wo, C);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

int epos_softmax_rows(float* x, int64_t* labels, size_t rows, int n, void* stream) {
  EPOS_CHECK_ARG(x && rows > 0 && n > 0);
  const long long blocks = ((long long)rows + 7) / 8;
  softmax_rows_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      x, labels, (long long)rows, n);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

int epos_softmax_rows_masked(float* x, const float* obj_conf, size_t pixels, int num_objs, int num_frags, float min_obj_conf,
                             void* stream) {
  EPOS_CHECK_ARG(x && obj_conf && pixels > 0 && num_objs > 0 && num_frags > 0);
  const long long blocks = ((long long)pixels * num_objs + 7) / 8;
  softmax_rows_masked_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      x, obj_conf, (long long)pixels, num_objs, num_frags, min_obj_conf);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

int epos_pwconv_simt(const float* a, int lda, const float* w, const float* bias, int bias_group_rows,
                     const float* residual, int ldr, float* d, int ldd, int M, int N, int K, int relu, void* stream) {
  EPOS_CHECK_ARG(a && w && d && M > 0 && N > 0 && K > 0 && lda >= K && ldd >= N);
  if (M <= SMALLM_MAX && !residual) {
    pwconv_smallm_kernel<<<ceil_div(N, 8), 256, 0, (cudaStream_t)stream>>>(a, lda, w, bias, bias_group_rows, d, ldd, M, N,
                                                                          K, relu);
    EPOS_LAUNCH_CHECK();
    return EPOS_OK;
  }
  dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
  pwconv_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, lda, w, bias, bias_group_rows, residual, ldr, d, ldd, M,
                                                            N, K, relu);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

int epos_preprocess_u8(const uint8_t* src, int in_h, int in_w, size_t src_pitch, int max_height_before_crop, int crop_h,
                       int crop_w, int off_y, int off_x, const double* K_in, float* dst, double* K_out, void* stream) {
  EPOS_CHECK_ARG(src && dst && in_h > 0 && in_w > 0 && src_pitch >= (size_t)3 * in_w && max_height_before_crop > 0);
  EPOS_CHECK_ARG(crop_h > 0 && crop_w > 0);
  // datagen.py:441-444: new height = min(max_height_before_crop, height), the width follows the same scale (truncated)
  const int new_h = in_h < max_height_before_crop ? in_h : max_height_before_crop;
  const float scale = (float)new_h / (float)in_h;
  const int new_w = (int)((float)in_w * scale);
  EPOS_CHECK_ARG(new_w > 0);
  if (off_y < 0 || off_x < 0 || off_y + crop_h > new_h || off_x + crop_w > new_w) {
    set_error("epos_preprocess_u8: crop %dx%d at (%d,%d) does not fit the resized image %dx%d (datagen.py:451-459 requires it)",
              crop_w, crop_h, off_x, off_y, new_w, new_h);
    return EPOS_ERR_INVALID_ARG;
  }
  if (K_in && K_out) {                                          // datagen.py:461-467 (float32 arithmetic in the reference)
    const float fx = (float)K_in[0] * scale, fy = (float)K_in[4] * scale;
    const float cx = (float)K_in[2] * scale - (float)off_x, cy = (float)K_in[5] * scale - (float)off_y;
    const double Ko[9] = {fx, 0.0, cx, 0.0, fy, cy, 0.0, 0.0, 1.0};
    for (int i = 0; i < 9; ++i) K_out[i] = Ko[i];
  }
  dim3 grid(ceil_div(crop_w, 256), crop_h);
  preprocess_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, in_h, in_w, (long long)src_pitch, dst, new_h, new_w, off_y,
                                                              off_x, crop_h, crop_w, in_h >= new_h ? 1 : 0);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

}  // extern "C"
