// Shared helpers for the epos_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/epos_b200.h"

namespace epos {

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

#define EPOS_CHECK_ARG(cond)                                                     \
  do {                                                                           \
    if (!(cond)) {                                                               \
      epos::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond); \
      return EPOS_ERR_INVALID_ARG;                                               \
    }                                                                            \
  } while (0)

#define EPOS_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      epos::set_error("%s:%d: CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(e__),    \
                      cudaGetErrorString(e__));                                                 \
      return EPOS_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

#define EPOS_LAUNCH_CHECK()            \
  do {                                 \
    epos::count_launch();              \
    EPOS_CUDA(cudaPeekAtLastError());  \
  } while (0)

// dw_tile.cu: TMA-staged tiled depthwise 3x3 (stride 1, rate 1/2/4); EPOS_ERR_UNSUPPORTED = use the strip kernel
int dwconv3x3_tiled(const float* x, int ldx, const float* w, const float* bias, float* y_f32, uint16_t* y_split,
                    int ldy_split, int B, int H, int W, int C, int rate, int relu_in, int relu_out, cudaStream_t stream);

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute applies to the device that is current at the time of the call, so one-time set-up is tracked
// PER DEVICE (a process may drive several GPUs).  Index of the current device into such a table.
constexpr int EPOS_MAX_DEVICES = 64;
inline int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= EPOS_MAX_DEVICES) return 0;
  return dev;
}

// fp32 -> (hi, lo) bf16 pair with hi + lo ~= x to 16-17 significant bits.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// Two values at once: hi = (bf16(a), bf16(b)) packed (a in the low half), lo = the bf16 residuals.  One packed
// conversion per pair (F2FP) instead of four scalar F2F; identical rounding (nearest-even) to split_bf16.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

}  // namespace epos
