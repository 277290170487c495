// Depthwise 3x3 (stride 1, rate 1/2/4) + folded BatchNorm (+ReLU before/after) as a TMA-staged, shared-memory tiled
// kernel for sm_100a.  Replaces slim.separable_conv2d(depth_multiplier=1, num_outputs=None) + BatchNorm of
// /root/reference/epos_lib/net_xception.py:150-170,270-296 and model.py:70-89 on the stride-1 layers.
//
// HBM-bound: algorithmic traffic = one f32 read of the input + one write of the output (f32 and/or split-bf16).
// A CTA owns a TH x 16 spatial tile of one 32-channel slab.  One cp.async.bulk.tensor.4d brings the haloed
// (TH+2r) x (16+2r) x 32-channel box into shared memory (rows of 128 B per pixel); out-of-image taps are zero-filled
// by the TMA unit, so the compute loop has no bounds tests.  A thread = 4 channels x a strip of 4 output pixels: the
// strip's 4+2r input columns are read once per filter row (LDS.128, 8 lanes cover the 128 B of a pixel: conflict-free)
// and reused across the 3 horizontal taps; the 9 taps of the lane's channels live in registers.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tma.cuh"

namespace epos {

constexpr int DT_TW = 16;          // tile width (output pixels)
constexpr int DT_SLAB = 32;        // channels per CTA
constexpr int DT_THREADS = 256;    // 8 channel groups x (4 strips per row x 8 rows per pass); three CTAs per SM

template <int R>
__global__ void __launch_bounds__(DT_THREADS, 3) dwconv3x3_tile_kernel(
    const __grid_constant__ CUtensorMap tmap, const float* __restrict__ w, const float* __restrict__ bias,
    float* __restrict__ y_f32, uint16_t* __restrict__ y_split, int ldy_split, long long plane_stride, int H, int W, int C,
    int TH,
    int tiles_y, int slabs, int relu_in, int relu_out) {
  constexpr int IW = DT_TW + 2 * R;
  extern __shared__ __align__(128) uint8_t dt_smem[];
  float4* tile = reinterpret_cast<float4*>(dt_smem + ((128u - (smem_u32(dt_smem) & 127u)) & 127u));
  __shared__ uint64_t bar;
  // slab is the fastest grid index: CTAs that run together read adjacent 128-byte chunks of the same pixels
  const int slab = blockIdx.x;
  const int x0 = blockIdx.y * DT_TW;
  const int y0 = (blockIdx.z % tiles_y) * TH;
  const int b = blockIdx.z / tiles_y;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&bar, (uint32_t)((TH + 2 * R) * IW * DT_SLAB * 4));
    tma_load_4d(tile, &tmap, &bar, slab * DT_SLAB, x0 - R, y0 - R, b);
  }
  const int cg = threadIdx.x & 7;              // 4-channel group inside the slab
  const int sx = (threadIdx.x >> 3) & 3;       // strip (4 pixels) inside the tile row
  const int sy = threadIdx.x >> 5;             // row inside a pass of 8 rows
  const int c = slab * DT_SLAB + cg * 4;
  const bool c_ok = c < C;
  float4 wk[9];
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c_ok) {
#pragma unroll
    for (int t = 0; t < 9; ++t) wk[t] = __ldg(reinterpret_cast<const float4*>(w + (size_t)t * C + c));
    bb = __ldg(reinterpret_cast<const float4*>(bias + c));
  } else {
#pragma unroll
    for (int t = 0; t < 9; ++t) wk[t] = bb;
  }
  const float in_floor = relu_in ? 0.f : -INFINITY;
  const float out_floor = relu_out ? 0.f : -INFINITY;
  __syncthreads();                              // barrier initialised before anyone polls it
  mbar_wait(&bar, 0);
  for (int row = sy; row < TH; row += 8) {
    float4 acc[4] = {bb, bb, bb, bb};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const float4* rp = tile + ((row + ky * R) * IW + sx * 4) * 8 + cg;
#pragma unroll
      for (int col = 0; col < 4 + 2 * R; ++col) {
        float4 v = rp[col * 8];
        v.x = fmaxf(v.x, in_floor); v.y = fmaxf(v.y, in_floor); v.z = fmaxf(v.z, in_floor); v.w = fmaxf(v.w, in_floor);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int p = col - kx * R;
          if (p >= 0 && p < 4) {
            const float4 ww = wk[ky * 3 + kx];
            acc[p].x = fmaf(v.x, ww.x, acc[p].x); acc[p].y = fmaf(v.y, ww.y, acc[p].y);
            acc[p].z = fmaf(v.z, ww.z, acc[p].z); acc[p].w = fmaf(v.w, ww.w, acc[p].w);
          }
        }
      }
    }
    const int oy = y0 + row;
    if (!c_ok || oy >= H) continue;
    const long long pix0 = ((long long)b * H + oy) * W + x0 + sx * 4;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (x0 + sx * 4 + p >= W) break;
      float4 a = acc[p];
      a.x = fmaxf(a.x, out_floor); a.y = fmaxf(a.y, out_floor); a.z = fmaxf(a.z, out_floor); a.w = fmaxf(a.w, out_floor);
      const long long o = (pix0 + p) * C + c;
      if (y_f32) *reinterpret_cast<float4*>(y_f32 + o) = a;
      const long long os = (pix0 + p) * ldy_split + c;
      if (y_split) {
        uint2 hi, lo;
        split_bf16x2(a.x, a.y, hi.x, lo.x);
        split_bf16x2(a.z, a.w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(y_split + os) = hi;
        *reinterpret_cast<uint2*>(y_split + plane_stride + os) = lo;
      }
    }
  }
}

// Large dilation rates (ASPP: 12 / 24 / 36 on a 60 x 80 map).  With rate r the rows y = ry (mod r) form an independent
// problem: a tap of such a row lies in a row of the same class, r pixels to the side.  A CTA therefore owns one row class
// of one image and one 32-channel slab: every input value is read from HBM exactly once (the register-strip kernel
// re-reads each value up to 9 times from L2), taps come from shared memory, padding is an index test.
__global__ void __launch_bounds__(DT_THREADS) dwconv3x3_rows_kernel(
    const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ bias,
    float* __restrict__ y_f32, uint16_t* __restrict__ y_split, int ldy_split, long long plane_stride, int H, int W, int C,
    int rate, int relu_in, int relu_out) {
  extern __shared__ __align__(128) uint8_t dt_smem[];
  float4* tile = reinterpret_cast<float4*>(dt_smem);
  // slab is the fastest grid index: CTAs that run together read adjacent 128-byte chunks of the same pixels
  const int slab = blockIdx.x, ry = blockIdx.y, b = blockIdx.z;
  const int Hs = (H - ry + rate - 1) / rate;          // rows ry, ry + r, ... of this class
  const int cg = threadIdx.x & 7;
  const int c = slab * DT_SLAB + cg * 4;
  const bool c_ok = c < C;
  const float in_floor = relu_in ? 0.f : -INFINITY;
  const float out_floor = relu_out ? 0.f : -INFINITY;
  const int n = Hs * W * 8;
  // eight independent loads per thread in flight (the loop body has no other work to hide the latency behind)
  for (int base = threadIdx.x; base < n; base += 8 * DT_THREADS) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * DT_THREADS;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < n && c_ok) {
        const int px = idx >> 3, i = px / W, xx = px - i * W;
        v[u] = __ldg(reinterpret_cast<const float4*>(x + (((long long)b * H + ry + i * rate) * W + xx) * ldx + c));
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * DT_THREADS;
      if (idx < n) {
        v[u].x = fmaxf(v[u].x, in_floor); v[u].y = fmaxf(v[u].y, in_floor);
        v[u].z = fmaxf(v[u].z, in_floor); v[u].w = fmaxf(v[u].w, in_floor);
        tile[idx] = c_ok ? v[u] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  float4 wk[9];
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 9; ++t) wk[t] = c_ok ? __ldg(reinterpret_cast<const float4*>(w + (size_t)t * C + c)) : bb;
  if (c_ok) bb = __ldg(reinterpret_cast<const float4*>(bias + c));
  __syncthreads();
  if (!c_ok) return;
  for (int idx = threadIdx.x; idx < n; idx += DT_THREADS) {
    const int px = idx >> 3, i = px / W, xx = px - i * W;
    float4 acc = bb;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ii = i + ky - 1;
      if (ii < 0 || ii >= Hs) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xs = xx + (kx - 1) * rate;
        if (xs < 0 || xs >= W) continue;
        const float4 v = tile[(ii * W + xs) * 8 + cg];
        const float4 ww = wk[ky * 3 + kx];
        acc.x = fmaf(v.x, ww.x, acc.x); acc.y = fmaf(v.y, ww.y, acc.y);
        acc.z = fmaf(v.z, ww.z, acc.z); acc.w = fmaf(v.w, ww.w, acc.w);
      }
    }
    acc.x = fmaxf(acc.x, out_floor); acc.y = fmaxf(acc.y, out_floor); acc.z = fmaxf(acc.z, out_floor); acc.w = fmaxf(acc.w, out_floor);
    const long long pix = ((long long)b * H + ry + i * rate) * W + xx;
    if (y_f32) *reinterpret_cast<float4*>(y_f32 + pix * C + c) = acc;
    if (y_split) {
      uint2 hi, lo;
      split_bf16x2(acc.x, acc.y, hi.x, lo.x);
      split_bf16x2(acc.z, acc.w, hi.y, lo.y);
      *reinterpret_cast<uint2*>(y_split + pix * ldy_split + c) = hi;
      *reinterpret_cast<uint2*>(y_split + plane_stride + pix * ldy_split + c) = lo;
    }
  }
}

// The same row-class kernel with the load phase handed to the TMA unit: one cp.async.bulk.tensor.4d per row of the class
// (box = 32 channels x W pixels; a traversal stride over H would do it in one copy, but elementStrides are limited to 8)
// on one mbarrier.  No registers or issue slots are spent on the copy, so the resident CTAs of an SM overlap one CTA's
// arithmetic with the others' loads; ReLU on the input moves to the tap read.
constexpr int DR_MAX_ROWS = 5;     // rows of a class a thread keeps in registers

__global__ void __launch_bounds__(DT_THREADS) dwconv3x3_rows_tma_kernel(
    const __grid_constant__ CUtensorMap tmap, const float* __restrict__ w, const float* __restrict__ bias,
    float* __restrict__ y_f32, uint16_t* __restrict__ y_split, int ldy_split, long long plane_stride, int H, int W, int C,
    int rate, int relu_in, int relu_out) {
  extern __shared__ __align__(128) uint8_t dt_smem[];
  float4* tile = reinterpret_cast<float4*>(dt_smem + ((128u - (smem_u32(dt_smem) & 127u)) & 127u));
  __shared__ uint64_t bar;
  const int slab = blockIdx.x, ry = blockIdx.y, b = blockIdx.z;
  const int Hs = (H - ry + rate - 1) / rate;          // rows ry, ry + r, ... of this class
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&bar, (uint32_t)(Hs * W * DT_SLAB * 4));
    for (int i = 0; i < Hs; ++i) tma_load_4d(tile + (size_t)i * W * 8, &tmap, &bar, slab * DT_SLAB, 0, ry + i * rate, b);
  }
  const int cg = threadIdx.x & 7;
  const int c = slab * DT_SLAB + cg * 4;
  const bool c_ok = c < C;
  const float in_floor = relu_in ? 0.f : -INFINITY;
  const float out_floor = relu_out ? 0.f : -INFINITY;
  float4 wk[9];
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 9; ++t) wk[t] = c_ok ? __ldg(reinterpret_cast<const float4*>(w + (size_t)t * C + c)) : bb;
  if (c_ok) bb = __ldg(reinterpret_cast<const float4*>(bias + c));
  __syncthreads();                                   // the barrier is initialised before anyone polls it
  mbar_wait(&bar, 0);
  if (!c_ok) return;
  // A thread owns one column (and 4 channels) of the class and walks its rows: an input value feeds the outputs of the
  // rows above, at and below it, so the 9 taps of an output cost 3 shared-memory reads instead of 9.
  for (int item = threadIdx.x; item < W * 8; item += DT_THREADS) {
    const int xx = item >> 3;
    float4 acc[DR_MAX_ROWS];
#pragma unroll
    for (int i = 0; i < DR_MAX_ROWS; ++i) acc[i] = bb;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xs = xx + (kx - 1) * rate;
      if (xs < 0 || xs >= W) continue;
#pragma unroll
      for (int ii = 0; ii < DR_MAX_ROWS; ++ii) {
        if (ii < Hs) {
          float4 v = tile[(ii * W + xs) * 8 + cg];
          v.x = fmaxf(v.x, in_floor); v.y = fmaxf(v.y, in_floor); v.z = fmaxf(v.z, in_floor); v.w = fmaxf(v.w, in_floor);
          if (ii + 1 < DR_MAX_ROWS) {                  // row below: this value is its ky = 0 tap
            const float4 ww = wk[kx];
            acc[ii + 1].x = fmaf(v.x, ww.x, acc[ii + 1].x); acc[ii + 1].y = fmaf(v.y, ww.y, acc[ii + 1].y);
            acc[ii + 1].z = fmaf(v.z, ww.z, acc[ii + 1].z); acc[ii + 1].w = fmaf(v.w, ww.w, acc[ii + 1].w);
          }
          {
            const float4 ww = wk[3 + kx];
            acc[ii].x = fmaf(v.x, ww.x, acc[ii].x); acc[ii].y = fmaf(v.y, ww.y, acc[ii].y);
            acc[ii].z = fmaf(v.z, ww.z, acc[ii].z); acc[ii].w = fmaf(v.w, ww.w, acc[ii].w);
          }
          if (ii >= 1) {                               // row above: its ky = 2 tap
            const float4 ww = wk[6 + kx];
            acc[ii - 1].x = fmaf(v.x, ww.x, acc[ii - 1].x); acc[ii - 1].y = fmaf(v.y, ww.y, acc[ii - 1].y);
            acc[ii - 1].z = fmaf(v.z, ww.z, acc[ii - 1].z); acc[ii - 1].w = fmaf(v.w, ww.w, acc[ii - 1].w);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < DR_MAX_ROWS; ++i) {
      if (i < Hs) {
        float4 a = acc[i];
        a.x = fmaxf(a.x, out_floor); a.y = fmaxf(a.y, out_floor); a.z = fmaxf(a.z, out_floor); a.w = fmaxf(a.w, out_floor);
        const long long pix = ((long long)b * H + ry + i * rate) * W + xx;
        if (y_f32) *reinterpret_cast<float4*>(y_f32 + pix * C + c) = a;
        if (y_split) {
          uint2 hi, lo;
          split_bf16x2(a.x, a.y, hi.x, lo.x);
          split_bf16x2(a.z, a.w, hi.y, lo.y);
          *reinterpret_cast<uint2*>(y_split + pix * ldy_split + c) = hi;
          *reinterpret_cast<uint2*>(y_split + plane_stride + pix * ldy_split + c) = lo;
        }
      }
    }
  }
}

static int launch_dw_rows(const float* x, int ldx, const float* w, const float* bias, float* y_f32, uint16_t* y_split,
                          int ldy_split, int B, int H, int W, int C, int rate, int relu_in, int relu_out, cudaStream_t stream) {
  const int hs_max = ceil_div(H, rate);
  const int smem = hs_max * W * DT_SLAB * 4;
  static std::atomic<int> attr[EPOS_MAX_DEVICES];
  const int dslot = device_slot();
  // TMA path: one box {32 channels, W, 1, 1} per row of the class; box extents are limited to 256
  static int use_tma = -1;
  if (use_tma < 0) { const char* e = getenv("EPOS_DW_ROWS_TMA"); use_tma = e ? atoi(e) : 1; }   // developer A/B
  PFN_encodeTiled enc = get_encode();
  if (use_tma && enc && W <= 256 && hs_max <= DR_MAX_ROWS && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ldx % 4) == 0) {
    CUtensorMap map;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)W * ldx * 4, (cuuint64_t)H * W * ldx * 4};
    cuuint32_t box[4] = {(cuuint32_t)DT_SLAB, (cuuint32_t)W, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) {
      static std::atomic<int> attr_t[EPOS_MAX_DEVICES];
      const int smem_t = smem + 128;
      if (smem_t > attr_t[dslot].load(std::memory_order_acquire)) {
        EPOS_CUDA(cudaFuncSetAttribute(dwconv3x3_rows_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_t));
        attr_t[dslot].store(smem_t, std::memory_order_release);
      }
      dim3 grid(ceil_div(C, DT_SLAB), rate < H ? rate : H, B);
      dwconv3x3_rows_tma_kernel<<<grid, DT_THREADS, smem_t, stream>>>(map, w, bias, y_f32, y_split, ldy_split,
                                                                      (long long)B * H * W * ldy_split, H, W, C, rate,
                                                                      relu_in, relu_out);
      EPOS_LAUNCH_CHECK();
      return EPOS_OK;
    }
  }
  if (smem > attr[dslot].load(std::memory_order_acquire)) {
    EPOS_CUDA(cudaFuncSetAttribute(dwconv3x3_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr[dslot].store(smem, std::memory_order_release);
  }
  dim3 grid(ceil_div(C, DT_SLAB), rate < H ? rate : H, B);
  dwconv3x3_rows_kernel<<<grid, DT_THREADS, smem, stream>>>(x, ldx, w, bias, y_f32, y_split, ldy_split,
                                                            (long long)B * H * W * ldy_split, H, W, C, rate, relu_in, relu_out);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

// f32 NHWC [B][H][W][ldx] viewed as {C, W, H, B}; box {32, 16+2r, TH+2r, 1}, no swizzle, zero fill outside.
static int make_dw_map(CUtensorMap* map, const float* x, int ldx, int B, int H, int W, int C, int TH, int R) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return EPOS_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)W * ldx * 4, (cuuint64_t)H * W * ldx * 4};
  cuuint32_t box[4] = {(cuuint32_t)DT_SLAB, (cuuint32_t)(DT_TW + 2 * R), (cuuint32_t)(TH + 2 * R), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (dw) failed (%d) B=%d H=%d W=%d C=%d ldx=%d TH=%d R=%d", (int)r, B, H, W, C, ldx, TH, R);
    return EPOS_ERR_CUDA;
  }
  return EPOS_OK;
}

template <int R>
static int launch_dw_tile(const CUtensorMap& map, const float* w, const float* bias, float* y_f32, uint16_t* y_split,
                          int ldy_split, int B, int H, int W, int C, int TH, int relu_in, int relu_out, cudaStream_t stream) {
  const int smem = (TH + 2 * R) * (DT_TW + 2 * R) * DT_SLAB * 4 + 128;
  static std::atomic<int> attr[EPOS_MAX_DEVICES];
  const int dslot = device_slot();
  if (smem > attr[dslot].load(std::memory_order_acquire)) {
    EPOS_CUDA(cudaFuncSetAttribute(dwconv3x3_tile_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr[dslot].store(smem, std::memory_order_release);
  }
  const int tiles_y = ceil_div(H, TH), slabs = ceil_div(C, DT_SLAB);
  dim3 grid(slabs, ceil_div(W, DT_TW), tiles_y * B);
  dwconv3x3_tile_kernel<R><<<grid, DT_THREADS, smem, stream>>>(map, w, bias, y_f32, y_split, ldy_split, (long long)B * H * W * ldy_split, H,
                                                              W, C, TH, tiles_y, slabs, relu_in, relu_out);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

// Returns EPOS_ERR_UNSUPPORTED (without setting an error) when the shape is not one this kernel handles; the caller
// then uses the register-strip kernel in cnn_kernels.cu.
int dwconv3x3_tiled(const float* x, int ldx, const float* w, const float* bias, float* y_f32, uint16_t* y_split,
                    int ldy_split, int B, int H, int W, int C, int rate, int relu_in, int relu_out, cudaStream_t stream) {
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (ldx % 4) != 0 || (C % 4) != 0) return EPOS_ERR_UNSUPPORTED;
  if (!(rate == 1 || rate == 2 || rate == 4)) {
    // large rates: one row class per CTA if it fits in shared memory (<= 96 KB keeps two CTAs per SM)
    if (rate >= 6 && (long long)ceil_div(H, rate) * W * DT_SLAB * 4 <= 96 * 1024 && B <= 65535)
      return launch_dw_rows(x, ldx, w, bias, y_f32, y_split, ldy_split, B, H, W, C, rate, relu_in, relu_out, stream);
    return EPOS_ERR_UNSUPPORTED;
  }
  // tile height: a multiple of 4 in [8, 24] that loads the fewest rows in total (tiles x (TH + 2 rate): halo and padded
  // rows both count), among the heights whose haloed tile leaves room for THREE CTAs per SM (the kernel is built for
  // 3 x 256 threads, 80 registers: a CTA waits for its TMA box before it computes, so the latency is hidden by the
  // other resident CTAs -- measured on B200: 2 -> 3 CTAs per SM takes the 60x80x728 layer from 53 to 46 us)
  int TH = 8, best = 1 << 30;
  for (int t = 8; t <= 24; t += 4) {
    if (t > 8 && (t + 2 * rate) * (DT_TW + 2 * rate) * DT_SLAB * 4 + 128 > 74 * 1024) break;
    const int rows = ceil_div(H, t) * (t + 2 * rate);
    if (rows <= best) { best = rows; TH = t; }
  }
  if (const char* e = getenv("EPOS_DW_TH")) { const int t = atoi(e); if (t >= 4 && t <= 32 && t % 4 == 0) TH = t; }   // developer A/B
  if ((long long)ceil_div(H, TH) * B > 65535 || ceil_div(W, DT_TW) > 65535) return EPOS_ERR_UNSUPPORTED;
  CUtensorMap map;
  int rc = make_dw_map(&map, x, ldx, B, H, W, C, TH, rate);
  if (rc) return rc;
  switch (rate) {
    case 1: return launch_dw_tile<1>(map, w, bias, y_f32, y_split, ldy_split, B, H, W, C, TH, relu_in, relu_out, stream);
    case 2: return launch_dw_tile<2>(map, w, bias, y_f32, y_split, ldy_split, B, H, W, C, TH, relu_in, relu_out, stream);
    default: return launch_dw_tile<4>(map, w, bias, y_f32, y_split, ldy_split, B, H, W, C, TH, relu_in, relu_out, stream);
  }
}

}  // namespace epos
