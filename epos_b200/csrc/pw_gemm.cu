// Pointwise (1x1) convolution as a persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[m][n] = act( sum_k A[m][k] W[n][k] + bias[g(m)][n] ) (+ residual[m][n])
//
// Replaces slim.conv2d(1x1)+BatchNorm(+ReLU)(+residual add) of the reference
// (/root/reference/epos_lib/net_xception.py:178-182,297-313; model.py:90-97,223-258,350-352,448-456).
//
// Precision: the reference computes in fp32.  To hold the 1e-3 parity bound through ~75 stacked GEMMs the
// operands are error-compensated bf16 pairs (x = hi + lo): D = Ahi*Whi + Ahi*Wlo + Alo*Whi, three
// kind::f16 MMAs per product with fp32 accumulation in TMEM (~2^-17 relative operand error).
//
// Structure (one CTA per SM, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor.3d loads of the {64 x 128 x 2} A box and the
//               {64 x BLOCK_N x 2} W box (both planes in one copy) into a STAGES-deep smem ring (128B swizzle)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit frees smem slots and
//               publishes the accumulator
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns per warp), bias / ReLU / residual, fp32 and/or
//               split-bf16 stores; double-buffered accumulators (2 x BLOCK_N TMEM columns) overlap the
//               epilogue of tile i with the main loop of tile i+1
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace epos {

constexpr int BLOCK_M = 128;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;

struct GemmEpilogue {
  const float* bias;
  const float* residual;
  float* d_f32;
  uint16_t* d_split;
  long long d_plane_stride;
  int bias_group_rows;
  int ldr, ldd, ldd_split;
  int relu;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  unsigned long long spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > (1ull << 26)) __trap();   // turn a protocol bug into an error, not a hang
  } while (!done);
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> f32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// K-major operand tile with BLOCK_K bf16 per row: BLOCK_K = 64 -> rows of 128 B, 128-byte swizzle, 8-row groups 1024 B
// apart; BLOCK_K = 32 -> rows of 64 B, 64-byte swizzle, 8-row groups 512 B apart.
template <int BLOCK_K>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address
  d |= (uint64_t)0 << 16;                               // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((8 * BLOCK_K * 2) >> 4) << 32;        // stride byte offset: one 8-row group
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
  d |= (uint64_t)(BLOCK_K == 64 ? 2 : 4) << 61;         // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BLOCK_N, int BLOCK_K>
struct GemmCfg {
  static constexpr int A_BYTES = 2 * BLOCK_M * BLOCK_K * 2;       // both planes
  static constexpr int B_BYTES = 2 * BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES < 2 ? 2 : ((200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N, int BLOCK_K>
__global__ void __launch_bounds__(NUM_THREADS, 1)
pw_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const GemmEpilogue ep, int M, int N, int K) {
  using Cfg = GemmCfg<BLOCK_N, BLOCK_K>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BLOCK_M;
        const int n0 = (tile % n_tiles) * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_3d(smem_a + stage * Cfg::A_BYTES, &tmap_a, &full_bar[stage], kb * BLOCK_K, m0, 0);
          tma_load_3d(smem_b + stage * Cfg::B_BYTES, &tmap_w, &full_bar[stage], kb * BLOCK_K, n0, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = (tile % n_tiles) * BLOCK_N;
        int umma_n = N - n0;
        umma_n = umma_n >= BLOCK_N ? BLOCK_N : ((umma_n + 15) & ~15);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(umma_n >> 3) << 17) |
                               ((uint32_t)(BLOCK_M >> 4) << 24);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t a_lo = a_hi + BLOCK_M * BLOCK_K * 2;
          const uint32_t b_hi = smem_u32(smem_b + stage * Cfg::B_BYTES);
          const uint32_t b_lo = b_hi + BLOCK_N * BLOCK_K * 2;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint32_t koff = k * UMMA_K * 2;
            const uint64_t da_hi = make_smem_desc<BLOCK_K>(a_hi + koff), da_lo = make_smem_desc<BLOCK_K>(a_lo + koff);
            const uint64_t db_hi = make_smem_desc<BLOCK_K>(b_hi + koff), db_lo = make_smem_desc<BLOCK_K>(b_lo + koff);
            umma_bf16(tmem_d, da_lo, db_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_bf16(tmem_d, da_hi, db_lo, idesc, 1u);
            umma_bf16(tmem_d, da_hi, db_hi, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);            // frees the smem slot when the MMAs above retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);                // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_f32 = ep.d_f32 && (ep.ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.d_f32) & 15) == 0);
    const bool vec_res = ep.residual && (ep.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0);
    const bool vec_split = ep.d_split && (ep.ldd_split % 8 == 0) && ((reinterpret_cast<uintptr_t>(ep.d_split) & 15) == 0) &&
                           (ep.d_plane_stride % 8 == 0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * BLOCK_M;
      const int n0 = (tile % n_tiles) * BLOCK_N;
      int n_valid = N - n0;
      if (n_valid > BLOCK_N) n_valid = BLOCK_N;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int m = m0 + quarter * 32 + lane;
      const bool row_ok = m < M;
      const float* brow = ep.bias ? ep.bias + (ep.bias_group_rows > 0 ? (long long)(m / ep.bias_group_rows) * N : 0) : nullptr;
      for (int c0 = 0; c0 < n_valid; c0 += 32) {
        uint32_t r[32];
        __syncwarp();                                  // .sync.aligned: the whole warp issues the load together
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0), r);
        if (row_ok) {
        const int nb = n0 + c0;
        const int cnt = (n_valid - c0) < 32 ? (n_valid - c0) : 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float t = __uint_as_float(r[j]);
          if (brow && j < cnt) t += __ldg(brow + nb + j);
          if (ep.relu) t = fmaxf(t, 0.f);
          v[j] = t;
        }
        if (ep.residual) {
          const float* rr = ep.residual + (long long)m * ep.ldr + nb;
          if (vec_res && cnt == 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(rr) + j);
              v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < cnt) v[j] += __ldg(rr + j);
          }
        }
        if (ep.d_f32) {
          float* dd = ep.d_f32 + (long long)m * ep.ldd + nb;
          if (vec_f32 && cnt == 32 && (nb % 4 == 0)) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              reinterpret_cast<float4*>(dd)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < cnt) dd[j] = v[j];
          }
        }
        if (ep.d_split) {
          uint16_t* dh = ep.d_split + (long long)m * ep.ldd_split + nb;
          uint16_t* dl = dh + ep.d_plane_stride;
          if (vec_split && cnt == 32 && (nb % 8 == 0)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t h[4], l[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                __nv_bfloat16 h0, l0, h1, l1;
                split_bf16(v[8 * j + 2 * q], h0, l0);
                split_bf16(v[8 * j + 2 * q + 1], h1, l1);
                h[q] = pack_bf16x2(h0, h1);
                l[q] = pack_bf16x2(l0, l1);
              }
              reinterpret_cast<uint4*>(dh)[j] = make_uint4(h[0], h[1], h[2], h[3]);
              reinterpret_cast<uint4*>(dl)[j] = make_uint4(l[0], l[1], l[2], l[3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < cnt) {
                __nv_bfloat16 h0, l0;
                split_bf16(v[j], h0, l0);
                dh[j] = __bfloat16_as_ushort(h0);
                dl[j] = __bfloat16_as_ushort(l0);
              }
          }
        }
        }  // row_ok
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// [2][rows][ld] bf16, box {block_k, box_rows, 2}
static int make_map(CUtensorMap* map, const void* base, int rows, int cols, int ld, size_t plane_stride, int box_rows,
                    int block_k) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return EPOS_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)block_k, (cuuint32_t)box_rows, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d plane=%zu box_rows=%d", (int)r, rows, cols, ld,
              plane_stride, box_rows);
    return EPOS_ERR_CUDA;
  }
  return EPOS_OK;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BLOCK_N, int BLOCK_K>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mw, const GemmEpilogue& ep, int M, int N, int K,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, BLOCK_K>;
  static bool attr = false;
  if (!attr) {
    EPOS_CUDA(cudaFuncSetAttribute(pw_gemm_kernel<BLOCK_N, BLOCK_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr = true;
  }
  const int tiles = ceil_div(M, BLOCK_M) * ceil_div(N, BLOCK_N);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  pw_gemm_kernel<BLOCK_N, BLOCK_K><<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(ma, mw, ep, M, N, K);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

}  // namespace epos

using namespace epos;

extern "C" int epos_pwconv_gemm(const uint16_t* a_split, int lda, size_t a_plane_stride, const uint16_t* w_split,
                                const float* bias, int bias_group_rows, const float* residual, int ldr, float* d_f32,
                                int ldd, uint16_t* d_split, int ldd_split, size_t d_plane_stride, int M, int N, int K,
                                int relu, void* stream) {
  EPOS_CHECK_ARG(a_split && w_split && (d_f32 || d_split));
  EPOS_CHECK_ARG(M > 0 && N > 0 && K > 0);
  EPOS_CHECK_ARG(lda >= K && (lda % 8) == 0 && (K % 8) == 0 && (a_plane_stride % 8) == 0);
  EPOS_CHECK_ARG((reinterpret_cast<uintptr_t>(a_split) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_split) & 15) == 0);
  EPOS_CHECK_ARG(!d_f32 || ldd >= N);
  EPOS_CHECK_ARG(!d_split || ldd_split >= N);
  EPOS_CHECK_ARG(!residual || ldr >= N);
  int bn = N > 128 ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32));
  // K block: 64 (128-byte swizzle, 2 pipeline stages at BLOCK_N = 256) or 32 (64-byte swizzle, 4 stages).  Measured on
  // B200 (B = 8, full network): BK = 32 is 4 % slower -- the kernel is bound by L2->SM throughput, not by pipeline depth
  // (DESIGN.md section 3), so the default stays 64.  EPOS_GEMM_BK overrides for A/B measurements.
  static int bk_env = -1;
  if (bk_env < 0) {
    const char* e = getenv("EPOS_GEMM_BK");
    bk_env = e ? atoi(e) : 0;
    if (bk_env != 32 && bk_env != 64) bk_env = 0;
  }
  const int bk = bk_env ? bk_env : 64;
  CUtensorMap ma, mw;
  int rc = make_map(&ma, a_split, M, K, lda, a_plane_stride, BLOCK_M, bk);
  if (rc) return rc;
  rc = make_map(&mw, w_split, N, K, K, (size_t)N * K, bn, bk);
  if (rc) return rc;
  GemmEpilogue ep;
  ep.bias = bias; ep.residual = residual; ep.d_f32 = d_f32; ep.d_split = d_split;
  ep.d_plane_stride = (long long)d_plane_stride; ep.bias_group_rows = bias_group_rows;
  ep.ldr = ldr; ep.ldd = ldd; ep.ldd_split = ldd_split; ep.relu = relu;
  cudaStream_t s = (cudaStream_t)stream;
  if (bk == 32) {
    switch (bn) {
      case 256: return launch_gemm<256, 32>(ma, mw, ep, M, N, K, s);
      case 128: return launch_gemm<128, 32>(ma, mw, ep, M, N, K, s);
      case 64: return launch_gemm<64, 32>(ma, mw, ep, M, N, K, s);
      default: return launch_gemm<32, 32>(ma, mw, ep, M, N, K, s);
    }
  }
  switch (bn) {
    case 256: return launch_gemm<256, 64>(ma, mw, ep, M, N, K, s);
    case 128: return launch_gemm<128, 64>(ma, mw, ep, M, N, K, s);
    case 64: return launch_gemm<64, 64>(ma, mw, ep, M, N, K, s);
    default: return launch_gemm<32, 64>(ma, mw, ep, M, N, K, s);
  }
}
