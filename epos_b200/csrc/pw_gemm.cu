// Pointwise (1x1) and dense 3x3 (atrous) convolutions as a persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[m][n] = act( sum_k A[m][k] W[n][k] + bias[g(m)][n] + residual[m][n] )      act = identity | ReLU | softmax-64
//
// Replaces slim.conv2d(1x1)+BatchNorm(+ReLU)(+residual add) of the reference
// (/root/reference/epos_lib/net_xception.py:178-182,297-313; model.py:90-97,223-258,350-352,448-456;
// net_resnet_v1_beta.py:72-88) and, in conv mode (K = 9 taps x C, A boxes fetched by 5-D TMA with zero fill),
// resnet_utils.conv2d_same at stride 1 (external/slim/nets/resnet_utils.py:77-122).
//
// Precision: the reference computes in fp32.  To hold the 1e-3 parity bound through ~75 stacked GEMMs the
// operands are error-compensated bf16 pairs (x = hi + lo): D = Ahi*Whi + Ahi*Wlo + Alo*Whi, three
// kind::f16 MMAs per product with fp32 accumulation in TMEM (~2^-17 relative operand error).
//
// Structure (one CTA per SM, 320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor.3d loads of the {64 x 128 x 2} A box and the
//               {64 x BLOCK_N x 2} W box (both planes in one copy) into a STAGES-deep smem ring (128B swizzle)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit frees smem slots and
//               publishes the accumulator
//   warps 2..9  epilogue (two per TMEM lane quarter, alternate 32-column chunks): tcgen05.ld, bias / ReLU / residual, fp32 and/or
//               split-bf16 stores through a per-warp smem transpose (coalesced); double-buffered accumulators
//               (2 x BLOCK_N TMEM columns) overlap the epilogue of tile i with the main loop of tile i+1
// Work is handed out dynamically (PieceMap + a global counter); PAIR = true runs the same roles on a cluster of two
// CTAs with tcgen05.mma.cta_group::2 (M = 256, each CTA holds half of the W tile): the leader CTA issues the MMAs and owns
// the barriers; the peer's producer signals the leader's "stage full" barrier (plain remote arrive + its TMA's complete_tx),
// tcgen05.commit multicasts "stage empty" / "accumulator full" to both CTAs, the piece index reaches the peer by st.async.
// Cross-CTA signals carry no .release.cluster (that is a MEMBAR.ALL.GPU per arrive); the one signal that follows a read of
// shared memory ("queue slot consumed") is predicated on the value read, because a remote arrive does not wait for an
// earlier LDS of the same thread.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tma.cuh"

namespace epos {

constexpr int BLOCK_M = 128;
constexpr int UMMA_K = 16;
constexpr int GEMM_BK = 64;        // K block (bf16 elements): rows of 128 B, 128-byte swizzle
constexpr int EPI_WARPS = 8;          // two epilogue warps per TMEM lane quarter, each taking alternate column chunks
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;

struct GemmEpilogue {
  const float* bias;
  const float* residual;
  float* d_f32;
  uint16_t* d_split;
  long long d_plane_stride;
  int bias_group_rows;
  int ldr, ldd, ldd_split;
  int relu;
  int softmax64;    // act == 2: softmax over aligned groups of 64 output columns (after bias), fused into the epilogue
};

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

// CTA-pair (cta_group::2) forms: one MMA spans the two SMs of a TPC (M = 256: each CTA's tensor core computes its 128
// rows, each CTA's smem holds half of the N rows of W); commits arrive on the same barrier offset in both CTAs.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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

// tcgen05.ld of 32 lanes x 32 columns, split into issue and wait so that independent global loads can be put in flight
// in between.  The wait names the destination registers as in/out operands: nothing may read them before it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// 3x3 (atrous) convolution as an implicit GEMM: K = 9 taps x C channels, an M tile is an 8 x 16 block of output
// pixels of one image, and the A box of tap (ky, kx) is the same block shifted by ((ky-1) rate, (kx-1) rate) --
// out-of-image rows/columns are zero-filled by the TMA unit, which is exactly TF 'SAME' padding at stride 1.
struct ConvGeom {
  int enabled;
  int H, W;                // output (= input) spatial size
  int tiles_x, tiles_y;    // ceil(W / 16), ceil(H / 8)
  int rate;
  int cpb;                 // channel blocks per tap = C / BLOCK_K
};
constexpr int CONV_TW = 16, CONV_TH = 8;

// Work distribution.  The output is cut into PIECES: full 128 x BLOCK_N tiles in n-major order (pieces handed out
// together share A tiles in L2), and -- when the tile count leaves a remainder of at most half a wave (the 38400 x 728
// layers are 900 tiles = 6.08 waves on 148 SMs) -- the remainder tiles cut into blocks of 64 columns, so that the tail
// costs a fraction of a tile time instead of a full one.  Pieces are handed out DYNAMICALLY: the producer warp takes
// the next piece index from a global counter and publishes it to the MMA and epilogue warps through a small
// shared-memory queue.  The kernel therefore makes progress with however many CTAs are resident -- it can share the
// GPU with the long-running pose-fitting CTAs of the previous batch (engine.py) instead of waiting for their SMs.
struct PieceMap {
  int n_tiles, bulk_end, block_n, unit, bpr, g0, num_pieces, tile_m;
  // m_tiles row-tiles of tile_m rows (128, or 256 for a CTA pair); G = CTAs (or CTA pairs) sharing the work
  __host__ __device__ PieceMap(int m_tiles, int N, int block_n_, int tile_m_, int G) {
    block_n = block_n_; tile_m = tile_m_;
    n_tiles = (N + block_n - 1) / block_n;
    const int num_tiles = m_tiles * n_tiles;
    int rem = num_tiles % G;
    unit = block_n < 64 ? block_n : 64;
    const int upt = block_n / unit;
    if (rem * 2 > G || upt == 1) rem = 0;          // a remainder above half a wave (or unsplittable tiles) stays whole
    bulk_end = num_tiles - rem;
    bpr = (N + unit - 1) / unit;                   // 64-column blocks per row of tiles
    g0 = (bulk_end / n_tiles) * bpr + (bulk_end % n_tiles) * upt;
    num_pieces = bulk_end + (m_tiles * bpr - g0);
  }
  // rows [m0, m0+tile_m) x columns [n0, n0+n_cols) (n_cols may run past N: mask with N)
  __host__ __device__ void decode(int id, int& m0, int& n0, int& n_cols) const {
    if (id < bulk_end) {
      m0 = (id / n_tiles) * tile_m; n0 = (id % n_tiles) * block_n; n_cols = block_n;
    } else {
      const int g = g0 + (id - bulk_end);
      const int row = g / bpr;
      m0 = row * tile_m; n0 = (g - row * bpr) * unit; n_cols = unit;
    }
  }
};
constexpr int SCHED_DEPTH = 4;
struct SchedSlot { int next, done; };               // global: next piece index, CTAs finished (the last one resets both)

template <int BLOCK_N, int BLOCK_K, bool PAIR = false>
struct GemmCfg {
  static constexpr int A_BYTES = 2 * BLOCK_M * BLOCK_K * 2;       // both planes
  static constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;     // W rows per plane held by one CTA
  static constexpr int B_BYTES = 2 * B_ROWS * BLOCK_K * 2;
  static constexpr int B_BOX_ROWS = PAIR ? 32 : (BLOCK_N < 64 ? BLOCK_N : 64);   // small W box (rows per plane) for partial pieces
  static constexpr int B_BOX_BYTES = B_BOX_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = EPI_WARPS * 32 * 32 * 4;    // epilogue transpose buffers, one per warp
  static constexpr int RING_BUDGET = 227 * 1024 - 1024 - 512 - STAGING_BYTES;
  static constexpr int STAGES = RING_BUDGET / STAGE_BYTES < 2 ? 2 : (RING_BUDGET / STAGE_BYTES > 8 ? 8 : RING_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers, piece queue*/ + STAGING_BYTES;
};

template <int BLOCK_N, int BLOCK_K, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
pw_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
               const __grid_constant__ CUtensorMap tmap_w64,
               const GemmEpilogue ep, const ConvGeom cg, SchedSlot* __restrict__ sched, int M, int N, int K, int dbg) {
  using Cfg = GemmCfg<BLOCK_N, BLOCK_K, PAIR>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (128B swizzle atoms) by an offset from the __shared__ symbol, so that the compiler still knows
  // the address space (STS/LDS instead of generic accesses in the epilogue)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint64_t* sched_full = bars + 2 * STAGES + 4;
  uint64_t* sched_empty = sched_full + SCHED_DEPTH;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sched_empty + SCHED_DEPTH);
  volatile int* sched_ids = reinterpret_cast<volatile int*>(tmem_ptr_smem + 1);
  static_assert((2 * STAGES + 4 + 2 * SCHED_DEPTH) * 8 + 4 + SCHED_DEPTH * 4 <= 512, "barrier area");
  float* staging = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
  const int m_tiles = cg.enabled ? (M / (cg.H * cg.W)) * cg.tiles_x * cg.tiles_y : (M + BLOCK_M - 1) / BLOCK_M;

  // CTA pair: rank 0 (leader) issues the MMAs and owns the barriers that both CTAs signal
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr int NCTA = PAIR ? 2 : 1;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w64) : "memory");
    // full: one arrive.expect_tx per producer; tmem_empty: the epilogue warps of each CTA; sched_empty: every consumer of
    // the piece queue (single CTA: MMA + epilogue warps; pair: + the peer's producer and its epilogue warps)
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], NCTA); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EPI_WARPS * NCTA); }
    for (int i = 0; i < SCHED_DEPTH; ++i) { mbar_init(&sched_full[i], 1); mbar_init(&sched_empty[i], (1 + EPI_WARPS) * NCTA); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if constexpr (PAIR) cluster_sync_all();            // both CTAs' barriers exist before anything remote touches them
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"((uint32_t)Cfg::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"((uint32_t)Cfg::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
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
      const PieceMap pm(PAIR ? (m_tiles + 1) / 2 : m_tiles, N, BLOCK_N, PAIR ? 2 * BLOCK_M : BLOCK_M,
                        PAIR ? gridDim.x / 2 : gridDim.x);
      int sslot = 0;
      uint32_t sphase = 0;
      int m0, n0, n_cols;
      int id = leader ? atomicAdd(&sched->next, 1) : 0;
      while (true) {
        if (leader) {
          // take the index after this one now: the atomic's round trip hides behind this piece's loads
          const int id_next = id < pm.num_pieces ? atomicAdd(&sched->next, 1) : pm.num_pieces;
          mbar_wait(&sched_empty[sslot], sphase ^ 1);
          const int pub = id < pm.num_pieces ? id : -1;
          sched_ids[sslot] = pub;
          if constexpr (PAIR) {
            // the peer's copy of the index travels with its own completion count (st.async + complete_tx)
            const uint32_t peer_bar = mapa_u32(&sched_full[sslot], 1);
            mbar_expect_tx_cluster(peer_bar, 4);
            st_async_cluster_u32(mapa_u32(const_cast<int*>(&sched_ids[sslot]), 1), (uint32_t)pub, peer_bar);
          }
          mbar_arrive(&sched_full[sslot]);                            // release: publishes the index to the consumers
          if (++sslot == SCHED_DEPTH) { sslot = 0; sphase ^= 1; }
          if (id >= pm.num_pieces) break;
          pm.decode(id, m0, n0, n_cols);
          id = id_next;
        } else {
          // peer CTA of a pair: a consumer of the leader's piece queue
          mbar_wait(&sched_full[sslot], sphase);
          const int got = sched_ids[sslot];
          // the arrive is predicated on the value read: a remote arrive does not wait for an earlier LDS of the same
          // thread, and the leader overwrites the slot as soon as the last consumer has signalled (seen on B200: a
          // consumer read the NEXT round's index)
          if (got >= -1) mbar_arrive_cluster(mapa_u32(&sched_empty[sslot], 0));
          if (++sslot == SCHED_DEPTH) { sslot = 0; sphase ^= 1; }
          if (got < 0) break;
          pm.decode(got, m0, n0, n_cols);
        }
        if constexpr (PAIR) m0 += (int)rank * BLOCK_M;               // this CTA's 128 rows of the 256-row tile
        int n_rows = N - n0;                                       // W rows this piece needs
        if (n_rows > n_cols) n_rows = n_cols;
        const bool full = n_cols == BLOCK_N;                        // whole tile: one box (rows past N are zero-filled)
        // pair: this CTA holds rows [wn0, wn0 + umma_n/2) of W (umma_n = the pair's MMA width)
        int wn0 = n0;
        if constexpr (PAIR) wn0 = n0 + (int)rank * (((n_rows + 15) & ~15) >> 1);
        const uint32_t full_addr = PAIR ? mapa_u32(&full_bar[0], 0) : 0u;   // leader's full barriers (cluster address)
        int cb = 0, cy0 = 0, cx0 = 0;
        if (cg.enabled) {
          const int t = m0 / BLOCK_M, per_img = cg.tiles_x * cg.tiles_y;
          cb = t / per_img;
          const int q = t - cb * per_img;
          cy0 = (q / cg.tiles_x) * CONV_TH; cx0 = (q % cg.tiles_x) * CONV_TW;
        }
        const int n_boxes = (n_rows + Cfg::B_BOX_ROWS - 1) / Cfg::B_BOX_ROWS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if constexpr (PAIR) {
            const uint32_t fb = full_addr + (uint32_t)stage * 8u;
            // developer switches (scripts/dev_gemm.py): 256 = the peer loads nothing, 512 = the leader loads nothing,
            // 1024 = nobody loads (pure barrier protocol rate)
            if ((dbg & 1024) || ((dbg & 256) && rank == 1) || ((dbg & 512) && rank == 0)) {
              mbar_expect_tx_cluster(fb, 0);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
              continue;
            }
            mbar_expect_tx_cluster(fb, Cfg::A_BYTES + (full ? Cfg::B_BYTES : 2 * Cfg::B_BOX_BYTES));
            if (cg.enabled) {
              const int tap = kb / cg.cpb, c0 = (kb - tap * cg.cpb) * BLOCK_K;
              const int ky = tap / 3, kx = tap - ky * 3;
              tma_load_5d_pair(smem_a + stage * Cfg::A_BYTES, &tmap_a, fb, c0, cx0 + (kx - 1) * cg.rate,
                               cy0 + (ky - 1) * cg.rate, cb, 0);
            } else {
              tma_load_3d_pair(smem_a + stage * Cfg::A_BYTES, &tmap_a, fb, kb * BLOCK_K, m0, 0);
            }
            // whole tiles: one box of BLOCK_N/2 rows per plane; 64-column tail pieces: one box of 32 rows per plane
            tma_load_3d_pair(smem_b + stage * Cfg::B_BYTES, full ? &tmap_w : &tmap_w64, fb, kb * BLOCK_K, wn0, 0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], ((dbg & 128) ? 0 : Cfg::A_BYTES) + ((dbg & 64) ? 0 : (full ? Cfg::B_BYTES : 2 * n_boxes * Cfg::B_BOX_BYTES)));
          if (cg.enabled) {
            const int tap = kb / cg.cpb, c0 = (kb - tap * cg.cpb) * BLOCK_K;
            const int ky = tap / 3, kx = tap - ky * 3;
            tma_load_5d(smem_a + stage * Cfg::A_BYTES, &tmap_a, &full_bar[stage], c0, cx0 + (kx - 1) * cg.rate,
                        cy0 + (ky - 1) * cg.rate, cb, 0);
          } else if (!(dbg & 128)) tma_load_3d(smem_a + stage * Cfg::A_BYTES, &tmap_a, &full_bar[stage], kb * BLOCK_K, m0, 0);
          if (!(dbg & 64)) {
            uint8_t* b_hi = smem_b + stage * Cfg::B_BYTES;
            uint8_t* b_lo = b_hi + BLOCK_N * BLOCK_K * 2;
            if (full) {
              tma_load_3d(b_hi, &tmap_w, &full_bar[stage], kb * BLOCK_K, n0, 0);      // both planes, all BLOCK_N rows
            } else
            for (int j = 0; j < n_boxes; ++j) {
              tma_load_3d(b_hi + j * Cfg::B_BOX_BYTES, &tmap_w64, &full_bar[stage], kb * BLOCK_K, n0 + j * Cfg::B_BOX_ROWS, 0);
              tma_load_3d(b_lo + j * Cfg::B_BOX_BYTES, &tmap_w64, &full_bar[stage], kb * BLOCK_K, n0 + j * Cfg::B_BOX_ROWS, 1);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (the leader CTA of a pair issues for both) =====================
    if (lane == 0 && leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const PieceMap pm(PAIR ? (m_tiles + 1) / 2 : m_tiles, N, BLOCK_N, PAIR ? 2 * BLOCK_M : BLOCK_M,
                        PAIR ? gridDim.x / 2 : gridDim.x);
      int sslot = 0;
      uint32_t sphase = 0;
      int m0, n0, n_cols;
      while (true) {
        mbar_wait(&sched_full[sslot], sphase);
        const int id = sched_ids[sslot];
        if (id >= -1) mbar_arrive(&sched_empty[sslot]);
        if (++sslot == SCHED_DEPTH) { sslot = 0; sphase ^= 1; }
        if (id < 0) break;
        pm.decode(id, m0, n0, n_cols);
        int umma_n = N - n0;
        if (umma_n > n_cols) umma_n = n_cols;
        umma_n = (umma_n + 15) & ~15;                              // columns past the piece are computed but never stored
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(umma_n >> 3) << 17) |
                               ((uint32_t)((NCTA * BLOCK_M) >> 4) << 24);
        // rows per plane of the W box this piece was loaded with (pair: half of the MMA width per CTA)
        const int b_plane_rows = PAIR ? (n_cols == BLOCK_N ? BLOCK_N / 2 : Cfg::B_BOX_ROWS) : BLOCK_N;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t a_lo = a_hi + BLOCK_M * BLOCK_K * 2;
          const uint32_t b_hi = smem_u32(smem_b + stage * Cfg::B_BYTES);
          const uint32_t b_lo = b_hi + (uint32_t)b_plane_rows * BLOCK_K * 2;
          // the last K block of a ragged K (728 = 11 x 64 + 24) only needs the UMMA_K steps that hold data
          const int k_left = K - kb * BLOCK_K;
          const int k_steps = k_left >= BLOCK_K ? BLOCK_K / UMMA_K : (k_left + UMMA_K - 1) / UMMA_K;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            if (k >= k_steps) break;
            const uint32_t koff = k * UMMA_K * 2;
            const uint64_t da_hi = make_smem_desc<BLOCK_K>(a_hi + koff), da_lo = make_smem_desc<BLOCK_K>(a_lo + koff);
            const uint64_t db_hi = make_smem_desc<BLOCK_K>(b_hi + koff), db_lo = make_smem_desc<BLOCK_K>(b_lo + koff);
            if (dbg & 8) continue;
            if constexpr (PAIR) {
              umma_bf16_pair(tmem_d, da_lo, db_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              umma_bf16_pair(tmem_d, da_hi, db_lo, idesc, 1u);
              umma_bf16_pair(tmem_d, da_hi, db_hi, idesc, 1u);
              continue;
            }
            umma_bf16(tmem_d, da_lo, db_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (dbg & 16) continue;
            umma_bf16(tmem_d, da_hi, db_lo, idesc, 1u);
            umma_bf16(tmem_d, da_hi, db_hi, idesc, 1u);
          }
          // frees the smem slot (in both CTAs of a pair) when the MMAs above retire
          if constexpr (PAIR) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (PAIR) umma_commit_pair(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    // Fast path precondition (uniform over the launch): 16-byte aligned rows everywhere, N a multiple of 4, and a
    // per-group bias whose groups are at least one warp-quarter tall (at most one group boundary per 32 rows).
    const bool fast =
        (N % 4 == 0) &&
        (!ep.d_f32 || ((ep.ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.d_f32) & 15) == 0))) &&
        (!ep.residual || ((ep.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0))) &&
        (!ep.d_split || ((ep.ldd_split % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.d_split) & 7) == 0) &&
                         (ep.d_plane_stride % 4 == 0))) &&
        (!ep.bias || (((reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0) &&
                      (ep.bias_group_rows == 0 || ep.bias_group_rows >= 32)));
    // Each warp transposes its 32x32 accumulator chunk through a private XOR-swizzled smem buffer (conflict-free both
    // ways): tcgen05.ld hands a lane one ROW, but coalesced global access wants 8 consecutive lanes on 128 contiguous
    // bytes of a row.  After the transpose lane (rsub, jj) owns columns 4*jj..4*jj+3 of rows rsub, rsub+4, ...
    float4* stg = reinterpret_cast<float4*>(staging + (warp - 2) * 1024);
    const int half = (warp - 2) >> 2;                // which of the two warps of this quarter: alternate column chunks
    const int rsub = lane >> 3, jj = lane & 7;
    const float relu_floor = ep.relu ? 0.f : -INFINITY;
    const PieceMap pm(PAIR ? (m_tiles + 1) / 2 : m_tiles, N, BLOCK_N, PAIR ? 2 * BLOCK_M : BLOCK_M,
                        PAIR ? gridDim.x / 2 : gridDim.x);
    int sslot = 0;
    uint32_t sphase = 0;
    int m0, n0, n_cols;
    while (true) {
      // lane 0 takes the index and hands it to the warp; its "slot consumed" signal is predicated on the value (see the
      // producer: the signal must not overtake the read)
      int id = 0;
      if (lane == 0) {
        mbar_wait(&sched_full[sslot], sphase);
        id = sched_ids[sslot];
      }
      id = __shfl_sync(0xffffffffu, id, 0);
      if (lane == 0 && id >= -1) {
        if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(&sched_empty[sslot], 0)); else mbar_arrive(&sched_empty[sslot]);
      }
      if (++sslot == SCHED_DEPTH) { sslot = 0; sphase ^= 1; }
      if (id < 0) break;
      pm.decode(id, m0, n0, n_cols);
      if constexpr (PAIR) m0 += (int)rank * BLOCK_M;
      int n_valid = N - n0;
      if (n_valid > n_cols) n_valid = n_cols;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row_base = m0 + quarter * 32;
      // output row of accumulator row (quarter, rr = 4 i + rsub), or -1 when it lies outside the problem
      long long mrow[8];
      if (cg.enabled) {
        const int t = m0 / BLOCK_M, per_img = cg.tiles_x * cg.tiles_y;
        const int cb = t / per_img, q = t - cb * per_img;
        const int cy0 = (q / cg.tiles_x) * CONV_TH + quarter * 2, cx0 = (q % cg.tiles_x) * CONV_TW;
        const bool tile_ok = t < m_tiles;             // the second half of the last CTA-pair tile may not exist
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int y = cy0 + (i >> 2), x = cx0 + 4 * (i & 3) + rsub;
          mrow[i] = (tile_ok && y < cg.H && x < cg.W) ? ((long long)cb * cg.H + y) * cg.W + x : -1;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) mrow[i] = (row_base + rsub + 4 * i < M) ? (long long)(row_base + rsub + 4 * i) : -1;
      }
      const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      if (dbg & 32) {
      } else if (fast && ep.softmax64) {
        // Fused softmax over groups of 64 columns (tf.nn.softmax over the fragment axis, model.py:676-678): a lane owns
        // one row, so the 64 logits of a group are 2 x 32 registers and max / sum need no cross-lane traffic.
#pragma unroll 1
        for (int c0 = half * 64; c0 < n_valid; c0 += 128) {
          uint32_t r0[32], r1[32];
          __syncwarp();
          tmem_ld32_issue(tacc + (uint32_t)c0, r0);
          tmem_ld32_wait(r0);
          tmem_ld32_issue(tacc + (uint32_t)(c0 + 32), r1);
          tmem_ld32_wait(r1);
          float mx = -INFINITY;
          if (ep.bias) {
            const float4* bp = reinterpret_cast<const float4*>(ep.bias + n0 + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 qa = __ldg(bp + j), qb = __ldg(bp + 8 + j);
              r0[4 * j] = __float_as_uint(__uint_as_float(r0[4 * j]) + qa.x);
              r0[4 * j + 1] = __float_as_uint(__uint_as_float(r0[4 * j + 1]) + qa.y);
              r0[4 * j + 2] = __float_as_uint(__uint_as_float(r0[4 * j + 2]) + qa.z);
              r0[4 * j + 3] = __float_as_uint(__uint_as_float(r0[4 * j + 3]) + qa.w);
              r1[4 * j] = __float_as_uint(__uint_as_float(r1[4 * j]) + qb.x);
              r1[4 * j + 1] = __float_as_uint(__uint_as_float(r1[4 * j + 1]) + qb.y);
              r1[4 * j + 2] = __float_as_uint(__uint_as_float(r1[4 * j + 2]) + qb.z);
              r1[4 * j + 3] = __float_as_uint(__uint_as_float(r1[4 * j + 3]) + qb.w);
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[j]), __uint_as_float(r1[j])));
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float e0 = __expf(__uint_as_float(r0[j]) - mx), e1 = __expf(__uint_as_float(r1[j]) - mx);
            r0[j] = __float_as_uint(e0); r1[j] = __float_as_uint(e1);
            sum += e0 + e1;
          }
          const float inv = 1.0f / sum;               // one division per row; 2 ulp from e / sum, far inside the 1e-3 bound
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t* r = h ? r1 : r0;
              stg[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * j]) * inv, __uint_as_float(r[4 * j + 1]) * inv,
                                                             __uint_as_float(r[4 * j + 2]) * inv, __uint_as_float(r[4 * j + 3]) * inv);
            }
            __syncwarp();
            const int col = n0 + c0 + 32 * h + jj * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = i * 4 + rsub;
              const float4 v = stg[rr * 8 + (jj ^ (rr & 7))];
              if (col < N && mrow[i] >= 0) *reinterpret_cast<float4*>(ep.d_f32 + mrow[i] * ep.ldd + col) = v;
            }
          }
        }
      } else if (fast) {
        int g0 = 0, boundary = 0x7fffffff;
        if (ep.bias && ep.bias_group_rows > 0) {
          g0 = row_base / ep.bias_group_rows;
          boundary = (g0 + 1) * ep.bias_group_rows;
        }
        const bool two_groups = boundary < row_base + 32 && boundary < M;
#pragma unroll 1
        for (int c0 = half * 32; c0 < n_valid; c0 += 64) {
          const int col = n0 + c0 + jj * 4;
          const bool col_ok = col < N;
          uint32_t r[32];
          __syncwarp();
          tmem_ld32_issue(tacc + (uint32_t)c0, r);
          // bias and residual do not depend on the accumulator: fetch them while the TMEM load is in flight
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
          if (ep.bias && col_ok) {
            b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + (long long)g0 * N + col));
            b1 = two_groups ? __ldg(reinterpret_cast<const float4*>(ep.bias + (long long)(g0 + 1) * N + col)) : b0;
          }
          float4 res[8];
          if (ep.residual && !(dbg & 4)) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              res[i] = (col_ok && mrow[i] >= 0) ? __ldg(reinterpret_cast<const float4*>(ep.residual + mrow[i] * ep.ldr + col))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          tmem_ld32_wait(r);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            stg[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                           __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + rsub;
            float4 v = stg[rr * 8 + (jj ^ (rr & 7))];
            const float4 b = (row_base + rr >= boundary) ? b1 : b0;
            // ReLU follows the residual add (ResNet bottleneck: relu(shortcut + residual), net_resnet_v1_beta.py:88);
            // the Xception units never combine the two
            v.x = fmaxf(v.x + b.x + res[i].x, relu_floor);
            v.y = fmaxf(v.y + b.y + res[i].y, relu_floor);
            v.z = fmaxf(v.z + b.z + res[i].z, relu_floor);
            v.w = fmaxf(v.w + b.w + res[i].w, relu_floor);
            if (col_ok && mrow[i] >= 0 && !(dbg & 2)) {
              if (ep.d_f32) *reinterpret_cast<float4*>(ep.d_f32 + mrow[i] * ep.ldd + col) = v;
              if (ep.d_split) {
                uint2 hi, lo;
                split_bf16x2(v.x, v.y, hi.x, lo.x);
                split_bf16x2(v.z, v.w, hi.y, lo.y);
                uint16_t* ph = ep.d_split + mrow[i] * ep.ldd_split + col;
                *reinterpret_cast<uint2*>(ph) = hi;
                *reinterpret_cast<uint2*>(ph + ep.d_plane_stride) = lo;
              }
            }
          }
        }
      } else {
        // Generic path (odd N or unaligned views; only the 22-channel object head takes it): lane = row, scalar stores.
        long long m = row_base + lane < M ? row_base + lane : -1;
        if (cg.enabled) {
          const int t = m0 / BLOCK_M, per_img = cg.tiles_x * cg.tiles_y;
          const int cb = t / per_img, q = t - cb * per_img, r_ = quarter * 32 + lane;
          const int y = (q / cg.tiles_x) * CONV_TH + (r_ >> 4), x = (q % cg.tiles_x) * CONV_TW + (r_ & 15);
          m = (t < m_tiles && y < cg.H && x < cg.W) ? ((long long)cb * cg.H + y) * cg.W + x : -1;
        }
        const float* brow = ep.bias ? ep.bias + (ep.bias_group_rows > 0 && m >= 0 ? (m / ep.bias_group_rows) * N : 0) : nullptr;
#pragma unroll 1
        for (int c0 = half * 32; c0 < n_valid; c0 += 64) {
          uint32_t r[32];
          __syncwarp();
          tmem_ld32_issue(tacc + (uint32_t)c0, r);
          tmem_ld32_wait(r);
          // stage through smem so that the scalar loop below can index dynamically without local memory
#pragma unroll
          for (int j = 0; j < 8; ++j)
            stg[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                           __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          __syncwarp();
          const int cnt = (n_valid - c0) < 32 ? (n_valid - c0) : 32;
          if (m >= 0) {
            const float* srow = reinterpret_cast<const float*>(stg + lane * 8);
#pragma unroll 1
            for (int j = 0; j < cnt; ++j) {
              const int nn = n0 + c0 + j;
              float t = srow[(((j >> 2) ^ (lane & 7)) << 2) + (j & 3)];
              if (brow) t += __ldg(brow + nn);
              if (ep.residual) t += __ldg(ep.residual + m * ep.ldr + nn);
              t = fmaxf(t, relu_floor);
              if (ep.d_f32) ep.d_f32[m * ep.ldd + nn] = t;
              if (ep.d_split) {
                __nv_bfloat16 h0, l0;
                split_bf16(t, h0, l0);
                ep.d_split[m * ep.ldd_split + nn] = __bfloat16_as_ushort(h0);
                ep.d_split[ep.d_plane_stride + m * ep.ldd_split + nn] = __bfloat16_as_ushort(l0);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(&tmem_empty[acc], 0)); else mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    // the last CTA to finish re-arms the counters for the next launch that uses this slot
    __threadfence();
    if (atomicAdd(&sched->done, 1) == (int)gridDim.x - 1) {
      sched->next = 0; sched->done = 0;
      __threadfence();
    }
  }
  if constexpr (PAIR) cluster_sync_all();            // the peer may still read its TMEM / signal barriers in this CTA
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) {
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                   : "memory");
    } else {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                   : "memory");
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------------------
// [2][rows][ld] bf16, box {block_k, box_rows, box_planes}
static int make_map(CUtensorMap* map, const void* base, int rows, int cols, int ld, size_t plane_stride, int box_rows,
                    int block_k, int box_planes) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return EPOS_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)block_k, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
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

// Ring of scheduler slots in device memory (zeroed once; every kernel leaves its slot zeroed).  A slot is reused
// SCHED_RING launches later, long after the launch that used it has drained.
constexpr int SCHED_RING = 256;
static SchedSlot* sched_slot() {
  static SchedSlot* ring[64] = {};
  static unsigned seq[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!ring[dev]) {
    if (cudaMalloc(&ring[dev], SCHED_RING * sizeof(SchedSlot)) != cudaSuccess) return nullptr;
    if (cudaMemset(ring[dev], 0, SCHED_RING * sizeof(SchedSlot)) != cudaSuccess) return nullptr;
  }
  return ring[dev] + (seq[dev]++ % SCHED_RING);
}

template <int BLOCK_N, int BLOCK_K, bool PAIR>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mw64, const GemmEpilogue& ep,
                       const ConvGeom& cg, int M, int N, int K, cudaStream_t stream, int dbg) {
  using Cfg = GemmCfg<BLOCK_N, BLOCK_K, PAIR>;
  static std::atomic<int> attr[EPOS_MAX_DEVICES];
  static int max_pairs_dev[EPOS_MAX_DEVICES];
  const int dslot = device_slot();
  int& max_pairs = max_pairs_dev[dslot];
  auto kern = pw_gemm_kernel<BLOCK_N, BLOCK_K, PAIR>;
  if (!attr[dslot].load(std::memory_order_acquire)) {
    EPOS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    if (PAIR) {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(2 * (num_sms() / 2)); q.blockDim = dim3(NUM_THREADS); q.dynamicSmemBytes = Cfg::SMEM_BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      q.attrs = at; q.numAttrs = 1;
      EPOS_CUDA(cudaOccupancyMaxActiveClusters(&max_pairs, kern, &q));
      if (getenv("EPOS_GEMM_VERBOSE")) fprintf(stderr, "[epos] pw_gemm pair: max active clusters %d (SMs %d)\n", max_pairs, num_sms());
      if (max_pairs <= 0) { set_error("pw_gemm: no CTA pair fits (cudaOccupancyMaxActiveClusters = %d)", max_pairs); return EPOS_ERR_CUDA; }
    }
    attr[dslot].store(1, std::memory_order_release);
  }
  // one CTA (or CTA pair) per SM (TPC); with less than a wave of tiles PieceMap spreads 64-column blocks
  const long long m_tiles = cg.enabled ? (long long)(M / (cg.H * cg.W)) * cg.tiles_x * cg.tiles_y : ceil_div(M, BLOCK_M);
  const long long units = (PAIR ? (m_tiles + 1) / 2 : m_tiles) * ceil_div(N, BLOCK_N < 64 ? BLOCK_N : 64);
  const int cap = PAIR ? max_pairs : num_sms();
  const int n_units = units < cap ? (int)units : cap;
  SchedSlot* slot = sched_slot();
  if (!slot) { set_error("pw_gemm: cannot allocate the scheduler slots"); return EPOS_ERR_CUDA; }
  if (PAIR) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * n_units); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    EPOS_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mw, mw64, ep, cg, slot, M, N, K, dbg));
    count_launch();
    return EPOS_OK;
  }
  kern<<<n_units, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(ma, mw, mw64, ep, cg, slot, M, N, K, dbg);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

// K block of a launch: 64 bf16 everywhere; EPOS_GEMM_BK=32 (developer switch) halves it for the pointwise 256-column tiles
static int gemm_bk_for(int bn, int conv) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("EPOS_GEMM_BK"); env = e ? atoi(e) : 64; }
  return (env == 32 && bn == 256 && !conv) ? 32 : GEMM_BK;
}

// W maps, epilogue descriptor and dispatch on the N tile shared by the pointwise and the 3x3 entry points.
// K block = 64 bf16 (128-byte swizzle, 2 smem stages at BLOCK_N = 256).  A 32-wide K block (64-byte swizzle, 4 stages)
// was measured 4-15 % slower on B200: the main loop is bound by L2->SM throughput, not by pipeline depth.
static int run_gemm(const CUtensorMap& ma, const uint16_t* w_split, int ldw, const float* bias, int bias_group_rows,
                    const float* residual, int ldr, float* d_f32, int ldd, uint16_t* d_split, int ldd_split,
                    size_t d_plane_stride, const ConvGeom& cg, int M, int N, int K, int relu, cudaStream_t s) {
  EPOS_CHECK_ARG(!d_f32 || ldd >= N);
  EPOS_CHECK_ARG(!d_split || ldd_split >= N);
  EPOS_CHECK_ARG(!residual || ldr >= N);
  EPOS_CHECK_ARG(relu >= 0 && relu <= 2);
  EPOS_CHECK_ARG(ldw >= K && (ldw % 8) == 0);
  const int bn = N > 128 ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32));
  const char* de = getenv("EPOS_GEMM_DEBUG");   // developer A/B switches (scripts/dev_gemm.py); 0 in production
  const int dbg = de ? atoi(de) : 0;
  // CTA pairs (cta_group::2, 256 x 256 tiles, each CTA holds half of the W tile) wherever BLOCK_N = 256.  Measured on B200
  // (scripts/dev_gemm.py, profiles/gemm_ab_r02c.log): 38400x728x728 112.7 -> 92.8 us, decoder 81 -> 72, heads -6 %, the
  // MMA-bound deep-K layers unchanged.  (Until round 2 the remote arrives were .release.cluster = a MEMBAR.ALL.GPU per
  // pipeline stage, which made pairs slower than single CTAs on every K < 1024 layer.)  EPOS_GEMM_PAIR = 0 forces single CTAs.
  static int pair_env = -1;
  if (pair_env < 0) { const char* e = getenv("EPOS_GEMM_PAIR"); pair_env = e ? atoi(e) : 1; }
  const long long m_tiles = cg.enabled ? (long long)(M / (cg.H * cg.W)) * cg.tiles_x * cg.tiles_y : ceil_div(M, BLOCK_M);
  const bool pair = pair_env && bn == 256 && m_tiles >= 2;
  // developer A/B (scripts/dev_gemm.py): EPOS_GEMM_BK=32 runs the 256-column tiles with a 32-wide K block (64-byte swizzle,
  // twice the pipeline depth); the caller's A map must then be built with the same K block (epos_pwconv_gemm does)
  const int bk = gemm_bk_for(bn, cg.enabled);
  CUtensorMap mw, mw64;
  int rc = make_map(&mw, w_split, N, K, ldw, (size_t)N * ldw, pair ? bn / 2 : bn, bk, 2);
  if (rc) return rc;
  rc = pair ? make_map(&mw64, w_split, N, K, ldw, (size_t)N * ldw, 32, bk, 2)
            : make_map(&mw64, w_split, N, K, ldw, (size_t)N * ldw, bn < 64 ? bn : 64, bk, 1);
  if (rc) return rc;
  GemmEpilogue ep;
  ep.bias = bias; ep.residual = residual; ep.d_f32 = d_f32; ep.d_split = d_split;
  ep.d_plane_stride = (long long)d_plane_stride; ep.bias_group_rows = bias_group_rows;
  ep.ldr = ldr; ep.ldd = ldd; ep.ldd_split = ldd_split; ep.relu = relu == 1; ep.softmax64 = relu == 2;
  if (relu == 2) {
    // fused softmax needs whole 64-column groups per piece and the aligned fast path; f32 output only
    EPOS_CHECK_ARG(d_f32 && !d_split && !residual && bias_group_rows == 0 && (N % 64) == 0 && (ldd % 4) == 0 &&
                   (reinterpret_cast<uintptr_t>(d_f32) & 15) == 0 && (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0));
  }
  if (bk == 32) {
    if (pair) return launch_gemm<256, 32, true>(ma, mw, mw64, ep, cg, M, N, K, s, dbg);
    return launch_gemm<256, 32, false>(ma, mw, mw64, ep, cg, M, N, K, s, dbg);
  }
  if (pair) return launch_gemm<256, GEMM_BK, true>(ma, mw, mw64, ep, cg, M, N, K, s, dbg);
  switch (bn) {
    case 256: return launch_gemm<256, GEMM_BK, false>(ma, mw, mw64, ep, cg, M, N, K, s, dbg);
    case 128: return launch_gemm<128, GEMM_BK, false>(ma, mw, mw64, ep, cg, M, N, K, s, dbg);
    case 64: return launch_gemm<64, GEMM_BK, false>(ma, mw, mw64, ep, cg, M, N, K, s, dbg);
    default: return launch_gemm<32, GEMM_BK, false>(ma, mw, mw64, ep, cg, M, N, K, s, dbg);
  }
}

// split-bf16 NHWC activations [2][B][H][W][C] viewed as {C, W, H, B, plane}; box {64, 16, 8, 1, 2} = the 128-row A tile
static int make_conv_map(CUtensorMap* map, const uint16_t* x, int B, int H, int W, int C, int ldx, size_t plane_stride) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return EPOS_ERR_CUDA; }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 2};
  cuuint64_t strides[4] = {(cuuint64_t)ldx * 2, (cuuint64_t)W * ldx * 2, (cuuint64_t)H * W * ldx * 2,
                           (cuuint64_t)plane_stride * 2};
  cuuint32_t box[5] = {(cuuint32_t)GEMM_BK, CONV_TW, CONV_TH, 1, 2};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<uint16_t*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (conv) failed (%d) B=%d H=%d W=%d C=%d ldx=%d", (int)r, B, H, W, C, ldx);
    return EPOS_ERR_CUDA;
  }
  return EPOS_OK;
}

}  // namespace epos

using namespace epos;

extern "C" int epos_pwconv_gemm(const uint16_t* a_split, int lda, size_t a_plane_stride, const uint16_t* w_split, int ldw,
                                const float* bias, int bias_group_rows, const float* residual, int ldr, float* d_f32,
                                int ldd, uint16_t* d_split, int ldd_split, size_t d_plane_stride, int M, int N, int K,
                                int relu, void* stream) {
  EPOS_CHECK_ARG(a_split && w_split && (d_f32 || d_split));
  EPOS_CHECK_ARG(M > 0 && N > 0 && K > 0);
  EPOS_CHECK_ARG(lda >= K && (lda % 8) == 0 && (K % 8) == 0 && (a_plane_stride % 8) == 0);
  EPOS_CHECK_ARG((reinterpret_cast<uintptr_t>(a_split) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_split) & 15) == 0);
  CUtensorMap ma;
  const int bn_ = N > 128 ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32));
  int rc = make_map(&ma, a_split, M, K, lda, a_plane_stride, BLOCK_M, gemm_bk_for(bn_, 0), 2);
  if (rc) return rc;
  ConvGeom cg = {};
  return run_gemm(ma, w_split, ldw, bias, bias_group_rows, residual, ldr, d_f32, ldd, d_split, ldd_split, d_plane_stride, cg,
                  M, N, K, relu, (cudaStream_t)stream);
}

extern "C" int epos_conv3x3_gemm(const uint16_t* x_split, int ldx, size_t x_plane_stride, const uint16_t* w_split, int ldw,
                                 const float* bias, const float* residual, int ldr, float* d_f32, int ldd,
                                 uint16_t* d_split, int ldd_split, size_t d_plane_stride, int B, int H, int W, int C,
                                 int N, int rate, int relu, void* stream) {
  EPOS_CHECK_ARG(x_split && w_split && (d_f32 || d_split));
  EPOS_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && N > 0 && rate >= 1 && relu != 2);
  EPOS_CHECK_ARG((C % GEMM_BK) == 0 && ldx >= C && (ldx % 8) == 0 && (x_plane_stride % 8) == 0);
  EPOS_CHECK_ARG((reinterpret_cast<uintptr_t>(x_split) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_split) & 15) == 0);
  EPOS_CHECK_ARG((long long)B * H * W < (1LL << 31));
  CUtensorMap ma;
  int rc = make_conv_map(&ma, x_split, B, H, W, C, ldx, x_plane_stride);
  if (rc) return rc;
  ConvGeom cg;
  cg.enabled = 1; cg.H = H; cg.W = W; cg.tiles_x = ceil_div(W, CONV_TW); cg.tiles_y = ceil_div(H, CONV_TH);
  cg.rate = rate; cg.cpb = C / GEMM_BK;
  return run_gemm(ma, w_split, ldw, bias, 0, residual, ldr, d_f32, ldd, d_split, ldd_split, d_plane_stride, cg, B * H * W, N,
                  9 * C, relu, (cudaStream_t)stream);
}

// Host view of the work distribution of the GEMM kernel (the same PieceMap code the device runs): piece i covers rows
// [out[3i], out[3i] + tile_m) x columns [out[3i+1], out[3i+1] + out[3i+2]).  Returns the number of pieces (writes at
// most `cap`).  Used by the CPU tests to check that every output element is produced exactly once.
extern "C" int epos_gemm_pieces(int m_tiles, int N, int block_n, int tile_m, int num_ctas, int32_t* out, int cap) {
  EPOS_CHECK_ARG(m_tiles > 0 && N > 0 && num_ctas > 0 && (tile_m == 128 || tile_m == 256));
  EPOS_CHECK_ARG(block_n == 32 || block_n == 64 || block_n == 128 || block_n == 256);
  const PieceMap pm(m_tiles, N, block_n, tile_m, num_ctas);
  for (int i = 0; i < pm.num_pieces && i < cap && out; ++i) pm.decode(i, out[3 * i], out[3 * i + 1], out[3 * i + 2]);
  return pm.num_pieces;
}
