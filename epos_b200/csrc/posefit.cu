// Batched single-instance pose fitting on sm_100a: the behaviour of pyprogressivex.find6DPoses
// (/root/reference/external/progressive-x/src/pyprogressivex/src/progressivex_python.cpp:36-134,222-336) for P
// independent (image, object) problems, one CTA per problem.
//
// GC-RANSAC (.../graph-cut-ransac/src/pygcransac/include/GCRANSAC.h:206-530) is sequential by construction: best-so-far
// updates, the LO trigger (iteration > 20 on a new best), the cumulative budget of 9 graph cuts and the coverage exit
// all depend on the order of hypotheses.  The kernels keep that order exactly while evaluating hypotheses in parallel:
//
//   prep     points (u_n, v_n, x, y, z), dense pixel ids, k-nearest neighbour graph (f32, 5-D), reverse adjacency
//   main     persistent per problem: 16 RANSAC passes at a time, ONE WARP PER PASS (sample -> Kneip P3P -> score every
//            solution over all N points with ballot counting and a per-warp pixel bitset), then an in-order replay of
//            the chunk applies the reference's update / early-out / LO-trigger / termination rules
//   cut      graph-cut labeling (GCRANSAC.h:812-920): f64 preflow-push in waves + reverse BFS = nodes that can still
//            reach the sink, which is what the reference's BK max-flow labels SINK (graph.h:112-115,478-488)
//   trials   the <= 20 inner fits of graphCutLocalOptimization (GCRANSAC.h:737-792), one warp per trial
//            (sample 21 inliers -> DLT + LM -> score), replayed in order
//   final    iterated least squares, final non-minimal fit, final LM refinement (GCRANSAC.h:480-521,
//            progressivex_python.cpp:257-312), pose record + labeling
//
// The host launches a FIXED schedule (prep, 11 x [main, cut, trials], final): a problem uses a slot only when its
// state says so, nothing is read back, so the whole sequence is CUDA-graph capturable.
#include "common.cuh"
#include "pose_fit.cuh"

namespace epos {
namespace pose {

constexpr int NMAX = 4096;          // max correspondences per problem (shared-memory resident point set)
constexpr int MAXNB = 8;            // storage stride of neighbour lists
constexpr int MAX_TRIALS = 20;
constexpr int THREADS = MAX_TRIALS * 32;      // fit kernel: 20 warps = one warp per LO trial
constexpr int WARPS = THREADS / 32;
constexpr int PT = 1024;                      // prep kernel threads
constexpr int TRIAL_THREADS = THREADS;
constexpr int PER_THREAD = (NMAX + THREADS - 1) / THREADS;   // contiguous points per thread in ordered compactions
constexpr int DIST_INF = 0x7fffffff;
constexpr size_t SMEM_CUT_DYN = 222 * 1024;
// Residual capacities below CUT_EPS count as saturated when the SINK segment is determined (capacities are O(lambda) =
// O(0.1)): exact-arithmetic ties (an outlier whose terminal capacity equals the total capacity of its arcs) are then
// resolved the same way by every max-flow algorithm, the oracle's Dinic included (oracle/posefit.cpp).
constexpr double CUT_EPS = 1e-9;
constexpr int TRACE_ROUNDS = 16, TRACE_COLS = 72;   // debugging aid (epos_fit_debug_trace)

enum Phase { PH_MAIN = 0, PH_LO = 1, PH_FINAL = 2, PH_DONE = 3 };

struct ProbState {
  double Kinv[9];
  double thr_n, sq_trunc;
  double best_model[12];
  double lo_model[12];
  double coverage;
  unsigned long long iter, max_iteration, seed;
  int N, used_pixels, valid, phase, pass;
  int best_value, best_inl, lo_value, lo_inl;
  int lo_runs, gc_count, lo_final, lo_stage;
  int ni, found, err, pad;
  int chunk_base, chunk_n;
  long long t_sample, t_score, t_replay, t_total;   // main_kernel clock64() accumulators (profiling aid)
  long long t_cut, t_trials, t_final, t_fit;        // other kernels; t_fit = non-minimal fits inside trials (warp 0)
  long long n_scored_main, n_scored_lo, n_scored_final;   // models scored over all N points (roofline accounting)
  long long pad2;
};

struct PassRecord;

struct Workspace {
  ProbState* st;
  double* pts;              // [P][7][NMAX]  rows: un, vn, x, y, z, u, v
  unsigned short* pix;      // [P][NMAX]
  short* nbr;               // [P][NMAX][MAXNB]
  unsigned char* owned;     // [P][NMAX][MAXNB]  1 = this (node, slot) owns an undirected edge
  int* rev_off;             // [P][NMAX+1]
  int* rev_idx;             // [P][NMAX*MAXNB]   entries x*MAXNB+k of edges owned by x that end in this node
  unsigned short* inl;      // [P][NMAX]
  double* flow;             // [P][NMAX*MAXNB]
  double* capf;             // [P][NMAX*MAXNB]
  double* exc;              // [P][NMAX]
  double* dd;               // [P][NMAX]
  int* dist;                // [P][NMAX]
  int* lstart;              // [P][NMAX+2]  BFS level boundaries
  unsigned short* order;    // [P][NMAX]    BFS queue
  PassRecord* recs;         // [P][CHUNK]   hypotheses of the current chunk of RANSAC passes
  int* trace;               // [P][TRACE_ROUNDS][TRACE_COLS] LO rounds: gc, ni, updated, lo_value, lo_inl, (ok, inl, pix) x 20
};

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static size_t workspace_layout(int P, void* base, Workspace* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return base ? (char*)base + o : (char*)nullptr; };
  char* p;
  p = take((size_t)P * sizeof(ProbState)); if (w) w->st = (ProbState*)p;
  p = take((size_t)P * 7 * NMAX * 8); if (w) w->pts = (double*)p;
  p = take((size_t)P * NMAX * 2); if (w) w->pix = (unsigned short*)p;
  p = take((size_t)P * NMAX * MAXNB * 2); if (w) w->nbr = (short*)p;
  p = take((size_t)P * NMAX * MAXNB); if (w) w->owned = (unsigned char*)p;
  p = take((size_t)P * (NMAX + 1) * 4); if (w) w->rev_off = (int*)p;
  p = take((size_t)P * NMAX * MAXNB * 4); if (w) w->rev_idx = (int*)p;
  p = take((size_t)P * NMAX * 2); if (w) w->inl = (unsigned short*)p;
  p = take((size_t)P * NMAX * MAXNB * 8); if (w) w->flow = (double*)p;
  p = take((size_t)P * NMAX * MAXNB * 8); if (w) w->capf = (double*)p;
  p = take((size_t)P * NMAX * 8); if (w) w->exc = (double*)p;
  p = take((size_t)P * NMAX * 8); if (w) w->dd = (double*)p;
  p = take((size_t)P * NMAX * 4); if (w) w->dist = (int*)p;
  p = take((size_t)P * (NMAX + 2) * 4); if (w) w->lstart = (int*)p;
  p = take((size_t)P * NMAX * 2); if (w) w->order = (unsigned short*)p;
  p = take((size_t)P * 80 * 432); if (w) w->recs = (PassRecord*)p;
  p = take((size_t)P * TRACE_ROUNDS * TRACE_COLS * 4); if (w) w->trace = (int*)p;
  return off;
}

__device__ inline unsigned long long iteration_bound(double confidence, int inl, int N) {
  if (confidence == 1.0) return ~0ULL;
  const double q = pow((double)inl / N, 3.0);
  const double l2 = log(1 - q);
  if (fabs(l2) < DBL_EPSILON) return ~0ULL;
  const double it = log(1.0 - confidence) / l2;
  return (unsigned long long)it + 1ULL;
}

// block-wide exclusive scan of one int per thread (THREADS threads); returns exclusive prefix, total via *total
template <int NW>
__device__ inline int block_excl_scan(int v, int* sh /* NW+1 ints */, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[w] = x;
  __syncthreads();
  if (w == 0) {
    int s = lane < NW ? sh[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    if (lane < NW) sh[lane] = s;      // inclusive over warps
  }
  __syncthreads();
  const int base = w > 0 ? sh[w - 1] : 0;
  *total = sh[NW - 1];
  __syncthreads();
  return base + x - v;
}

// =====================================================================================================
// prep
// =====================================================================================================
__global__ void __launch_bounds__(PT, 1)
prep_kernel(Workspace ws, const double* __restrict__ c2d, const double* __restrict__ c3d, const int* __restrict__ offsets,
            const int* __restrict__ counts, const double* __restrict__ Kmat, const unsigned long long* __restrict__ seeds,
            epos_fit_params prm, int* __restrict__ labeling, double* __restrict__ poses) {
  extern __shared__ unsigned char smem_raw[];
  const int p = blockIdx.x, tid = threadIdx.x;
  ProbState* st = ws.st + p;
  const int N = counts[p];
  const int off = offsets[p];
  double* rec = poses + (size_t)p * EPOS_POSE_RECORD_DOUBLES;
  if (tid < EPOS_POSE_RECORD_DOUBLES) rec[tid] = 0.0;
  for (int i = tid; i < N; i += PT) labeling[off + i] = 0;
  const bool ok = N >= 6 && N <= NMAX;          // scripts/infer.py:417-422 skips objects with < 6 correspondences
  if (tid == 0) {
    st->N = N; st->valid = ok ? 1 : 0; st->err = N > NMAX ? 1 : 0;
    st->phase = ok ? PH_MAIN : PH_DONE;
    st->iter = 0; st->pass = 0; st->best_value = 0; st->best_inl = 0; st->coverage = 0.0;
    st->lo_runs = 0; st->gc_count = 0; st->lo_final = 0; st->lo_stage = 0; st->ni = 0; st->found = 0;
    st->lo_value = 0; st->lo_inl = 0; st->used_pixels = 0; st->chunk_base = 0; st->chunk_n = 0;
    st->t_sample = st->t_score = st->t_replay = st->t_total = 0;
    st->t_cut = st->t_trials = st->t_final = st->t_fit = 0;
    st->n_scored_main = st->n_scored_lo = st->n_scored_final = 0;
    st->seed = seeds[p];
    st->max_iteration = iteration_bound(1.0 /* set below */, 1, N > 0 ? N : 1);
    for (int i = 0; i < 12; ++i) { st->best_model[i] = 0.0; st->lo_model[i] = 0.0; }
    const double* K = Kmat + 9 * p;
    double Ki[9];
    if (!inv3(K, Ki)) { for (int i = 0; i < 9; ++i) Ki[i] = 0.0; st->valid = 0; st->phase = PH_DONE; }
    for (int i = 0; i < 9; ++i) st->Kinv[i] = Ki[i];
    st->thr_n = prm.threshold / (0.5 * (K[0] + K[4]));
    const double tt = 1.5 * st->thr_n;
    st->sq_trunc = tt * tt;
  }
  __syncthreads();
  if (!ok || st->phase == PH_DONE) return;

  // ---- points ----
  double* P5 = ws.pts + (size_t)p * 7 * NMAX;
  float* q = reinterpret_cast<float*>(smem_raw);                         // [5][NMAX] f32 neighbourhood coordinates
  unsigned long long* table = reinterpret_cast<unsigned long long*>(smem_raw + 5 * NMAX * 4);   // 8192 hash slots
  int* scan_sh = reinterpret_cast<int*>(smem_raw + 5 * NMAX * 4 + 8192 * 8);
  for (int i = tid; i < 8192; i += PT) table[i] = ~0ULL;
  __syncthreads();
  for (int i = tid; i < N; i += PT) {
    const double u = c2d[2 * (size_t)(off + i)], v = c2d[2 * (size_t)(off + i) + 1];
    const double x = c3d[3 * (size_t)(off + i)], y = c3d[3 * (size_t)(off + i) + 1], z = c3d[3 * (size_t)(off + i) + 2];
    P5[i] = st->Kinv[0] * u + st->Kinv[1] * v + st->Kinv[2];
    P5[NMAX + i] = st->Kinv[3] * u + st->Kinv[4] * v + st->Kinv[5];
    P5[2 * NMAX + i] = x; P5[3 * NMAX + i] = y; P5[4 * NMAX + i] = z;
    P5[5 * NMAX + i] = u; P5[6 * NMAX + i] = v;
    q[i] = (float)u; q[NMAX + i] = (float)v;
    q[2 * NMAX + i] = (float)(x * prm.scaling_from_millimeters);
    q[3 * NMAX + i] = (float)(y * prm.scaling_from_millimeters);
    q[4 * NMAX + i] = (float)(z * prm.scaling_from_millimeters);
    // distinct (int)u,(int)v pixels (progressivex_python.cpp:97-99, scoring_function.h:247-249): hash insert
    const unsigned long long key = ((unsigned long long)(unsigned int)(int)u << 32) | (unsigned long long)(unsigned int)(int)v;
    unsigned int h = (unsigned int)(mix64(key) & 8191ULL);
    for (;;) {
      const unsigned long long prev = atomicCAS(&table[h], ~0ULL, key);
      if (prev == ~0ULL || prev == key) break;
      h = (h + 1) & 8191u;
    }
  }
  __syncthreads();
  // dense ids of the occupied slots
  {
    int cnt = 0;
    const int per = 8192 / PT;
    for (int k = 0; k < per; ++k) cnt += table[tid * per + k] != ~0ULL;
    int total;
    int base = block_excl_scan<PT / 32>(cnt, scan_sh, &total);
    // the key's high bits are overwritten by the dense id: slot -> (id << 40 | low 40 bits kept for matching is not
    // possible), so ids go to a parallel array placed over the neighbourhood scratch that follows
    unsigned short* slot_id = reinterpret_cast<unsigned short*>(scan_sh + 64);
    for (int k = 0; k < per; ++k)
      if (table[tid * per + k] != ~0ULL) slot_id[tid * per + k] = (unsigned short)base++;
    if (tid == 0) st->used_pixels = total;
    __syncthreads();
    unsigned short* pixg = ws.pix + (size_t)p * NMAX;
    for (int i = tid; i < N; i += PT) {
      const double u = c2d[2 * (size_t)(off + i)], v = c2d[2 * (size_t)(off + i) + 1];
      const unsigned long long key = ((unsigned long long)(unsigned int)(int)u << 32) | (unsigned long long)(unsigned int)(int)v;
      unsigned int h = (unsigned int)(mix64(key) & 8191ULL);
      while (table[h] != key) h = (h + 1) & 8191u;
      pixg[i] = slot_id[h];
    }
  }
  __syncthreads();

  // ---- neighbourhood graph: the max_neighbors nearest points within the radius (f32, 5-D, ties by index) ----
  // A uniform grid over (u, v) with cells at least one radius wide bounds the search to 3 x 3 cells.
  const int KN = prm.max_neighbors < MAXNB ? prm.max_neighbors : MAXNB;
  const float rad = (float)prm.neighborhood_ball_radius;
  const float r2 = rad * rad;
  short* nbr = ws.nbr + (size_t)p * NMAX * MAXNB;
  constexpr int MAXCELL = 4096;
  int* cell_start = reinterpret_cast<int*>(table);                        // [ncell + 1] (hash table is dead now)
  unsigned short* sorted = reinterpret_cast<unsigned short*>(cell_start + MAXCELL + 8);   // [N] point ids by cell
  __shared__ float s_red[4][PT / 32];
  __shared__ float s_box[4];
  {
    float mnu = INFINITY, mnv = INFINITY, mxu = -INFINITY, mxv = -INFINITY;
    for (int i = tid; i < N; i += PT) {
      mnu = fminf(mnu, q[i]); mxu = fmaxf(mxu, q[i]); mnv = fminf(mnv, q[NMAX + i]); mxv = fmaxf(mxv, q[NMAX + i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mnu = fminf(mnu, __shfl_xor_sync(0xffffffffu, mnu, o)); mxu = fmaxf(mxu, __shfl_xor_sync(0xffffffffu, mxu, o));
      mnv = fminf(mnv, __shfl_xor_sync(0xffffffffu, mnv, o)); mxv = fmaxf(mxv, __shfl_xor_sync(0xffffffffu, mxv, o));
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = mnu; s_red[1][tid >> 5] = mxu; s_red[2][tid >> 5] = mnv; s_red[3][tid >> 5] = mxv; }
    __syncthreads();
    if (tid == 0) {
      for (int k = 1; k < PT / 32; ++k) {
        s_red[0][0] = fminf(s_red[0][0], s_red[0][k]); s_red[1][0] = fmaxf(s_red[1][0], s_red[1][k]);
        s_red[2][0] = fminf(s_red[2][0], s_red[2][k]); s_red[3][0] = fmaxf(s_red[3][0], s_red[3][k]);
      }
      float cs = rad * 1.0001f + 1e-6f;                                   // >= radius, so neighbours are within +-1 cell
      int gw, gh;
      for (;;) {
        gw = (int)floorf((s_red[1][0] - s_red[0][0]) / cs) + 1;
        gh = (int)floorf((s_red[3][0] - s_red[2][0]) / cs) + 1;
        if (gw > 0 && gh > 0 && (long long)gw * gh <= MAXCELL) break;
        cs *= 2.0f;
        if (!(cs < 1e30f)) { gw = gh = 1; break; }                        // non-finite coordinates: one cell
      }
      s_box[0] = s_red[0][0]; s_box[1] = s_red[2][0]; s_box[2] = cs; s_box[3] = __int_as_float(gw | (gh << 16));
    }
    __syncthreads();
  }
  const float bu = s_box[0], bv = s_box[1], ics = 1.0f / s_box[2];
  const int gw = __float_as_int(s_box[3]) & 0xffff, gh = __float_as_int(s_box[3]) >> 16;
  const int ncell = gw * gh;
  auto cell_of = [&](int i, int* cx, int* cy) {
    int x = (int)floorf((q[i] - bu) * ics), y = (int)floorf((q[NMAX + i] - bv) * ics);
    *cx = x < 0 ? 0 : (x >= gw ? gw - 1 : x);
    *cy = y < 0 ? 0 : (y >= gh ? gh - 1 : y);
  };
  for (int c = tid; c <= ncell; c += PT) cell_start[c] = 0;
  __syncthreads();
  for (int i = tid; i < N; i += PT) { int cx, cy; cell_of(i, &cx, &cy); atomicAdd(&cell_start[cy * gw + cx + 1], 1); }
  __syncthreads();
  {
    // inclusive scan of cell_start[1..ncell] (<= 8192 cells = 16 per thread)
    const int per = (MAXCELL + PT - 1) / PT;
    int loc = 0;
    for (int k = 0; k < per; ++k) { const int c = tid * per + k + 1; if (c <= ncell) loc += cell_start[c]; }
    int total;
    int base = block_excl_scan<PT / 32>(loc, scan_sh, &total);
    for (int k = 0; k < per; ++k) { const int c = tid * per + k + 1; if (c <= ncell) { base += cell_start[c]; cell_start[c] = base; } }
  }
  __syncthreads();
  // fill: cursors run backwards from the cell ends; the order inside a cell is arbitrary (the selection below compares
  // (distance, index) pairs, so the result does not depend on it)
  int* cursor = cell_start + MAXCELL + 8 + (NMAX / 2) + 8;                // after `sorted`
  for (int c = tid; c < ncell; c += PT) cursor[c] = cell_start[c + 1];
  __syncthreads();
  for (int i = tid; i < N; i += PT) { int cx, cy; cell_of(i, &cx, &cy); sorted[atomicSub(&cursor[cy * gw + cx], 1) - 1] = (unsigned short)i; }
  __syncthreads();
  for (int i = tid; i < N; i += PT) {
    float bd[MAXNB];
    int bj[MAXNB];
#pragma unroll
    for (int k = 0; k < MAXNB; ++k) { bd[k] = INFINITY; bj[k] = -1; }
    const float a0 = q[i], a1 = q[NMAX + i], a2 = q[2 * NMAX + i], a3 = q[3 * NMAX + i], a4 = q[4 * NMAX + i];
    int cx, cy;
    cell_of(i, &cx, &cy);
    for (int yy = max(cy - 1, 0); yy <= min(cy + 1, gh - 1); ++yy) {
      const int c0 = yy * gw + max(cx - 1, 0), c1 = yy * gw + min(cx + 1, gw - 1);
      for (int s_ = cell_start[c0]; s_ < cell_start[c1 + 1]; ++s_) {      // the <= 3 cells of a row are contiguous
        const int j = sorted[s_];
        float e, d = 0.f;
        e = a0 - q[j]; d = __fmaf_rn(e, e, d);
        e = a1 - q[NMAX + j]; d = __fmaf_rn(e, e, d);
        e = a2 - q[2 * NMAX + j]; d = __fmaf_rn(e, e, d);
        e = a3 - q[3 * NMAX + j]; d = __fmaf_rn(e, e, d);
        e = a4 - q[4 * NMAX + j]; d = __fmaf_rn(e, e, d);
        if (j == i || !(d <= r2)) continue;
        // insert keeping (d, j) ascending lexicographically
        float cd = d; int cj = j;
        bool ins = false;
#pragma unroll
        for (int k = 0; k < MAXNB; ++k) {
          if (ins || bj[k] < 0 || cd < bd[k] || (cd == bd[k] && cj < bj[k])) {
            const float td = bd[k]; const int tj = bj[k];
            bd[k] = cd; bj[k] = cj; cd = td; cj = tj;
            ins = true;
            if (cj < 0) break;
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < MAXNB; ++k) nbr[(size_t)i * MAXNB + k] = (short)((k < KN) ? bj[k] : -1);
  }
  __syncthreads();
  // ---- edge ownership + reverse adjacency (GCRANSAC.h:864-907: each undirected pair is added once, by the first
  // endpoint that lists it in ascending point order) ----
  unsigned char* owned = ws.owned + (size_t)p * NMAX * MAXNB;
  int* rev_off = ws.rev_off + (size_t)p * (NMAX + 1);
  int* rev_idx = ws.rev_idx + (size_t)p * NMAX * MAXNB;
  int* indeg = reinterpret_cast<int*>(smem_raw);                         // q is dead now
  for (int i = tid; i <= N; i += PT) indeg[i] = 0;
  __syncthreads();
  for (int i = tid; i < N; i += PT)
    for (int k = 0; k < MAXNB; ++k) {
      const int j = nbr[(size_t)i * MAXNB + k];
      unsigned char own = 0;
      if (j >= 0 && j != i) {
        own = 1;
        if (j < i)
          for (int m = 0; m < MAXNB; ++m) own &= (nbr[(size_t)j * MAXNB + m] != i);
      }
      owned[(size_t)i * MAXNB + k] = own;
      if (own) atomicAdd(&indeg[j], 1);
    }
  __syncthreads();
  {
    // exclusive scan of indeg over N nodes (N <= 4096 = 8 per thread)
    const int per = NMAX / PT;
    int loc[per];
    int s = 0;
    for (int k = 0; k < per; ++k) { const int i = tid * per + k; loc[k] = i < N ? indeg[i] : 0; s += loc[k]; }
    int* scan2 = indeg + NMAX + 8;
    int total;
    int base = block_excl_scan<PT / 32>(s, scan2, &total);
    for (int k = 0; k < per; ++k) { const int i = tid * per + k; if (i < N) rev_off[i] = base; base += loc[k]; }
    if (tid == 0) rev_off[N] = total;
    __syncthreads();
    for (int i = tid; i < N; i += PT) indeg[i] = rev_off[i];          // fill cursors
    __syncthreads();
    for (int i = tid; i < N; i += PT)
      for (int k = 0; k < MAXNB; ++k)
        if (owned[(size_t)i * MAXNB + k]) {
          const int j = nbr[(size_t)i * MAXNB + k];
          rev_idx[atomicAdd(&indeg[j], 1)] = i * MAXNB + k;
        }
    __syncthreads();
    for (int i = tid; i < N; i += PT) {                                // sort each short list (reproducibility)
      const int b = rev_off[i], e = rev_off[i + 1];
      for (int a = b + 1; a < e; ++a) {
        const int v = rev_idx[a];
        int c = a - 1;
        while (c >= b && rev_idx[c] > v) { rev_idx[c + 1] = rev_idx[c]; --c; }
        rev_idx[c + 1] = v;
      }
    }
  }
}

// =====================================================================================================
// shared-memory point set + warp scoring
// =====================================================================================================
struct SmemPoints {
  double* un; double* vn; double* x; double* y; double* z;
  unsigned short* pix;
};

__device__ inline unsigned char* load_points(unsigned char* smem, const Workspace& ws, int p, int N, SmemPoints* sp,
                                            bool& loaded) {
  double* d = reinterpret_cast<double*>(smem);
  sp->un = d; sp->vn = d + NMAX; sp->x = d + 2 * NMAX; sp->y = d + 3 * NMAX; sp->z = d + 4 * NMAX;
  sp->pix = reinterpret_cast<unsigned short*>(d + 5 * NMAX);
  if (loaded) return smem + 5 * NMAX * 8 + NMAX * 2;
  loaded = true;
  const double* P5 = ws.pts + (size_t)p * 7 * NMAX;
  const unsigned short* pg = ws.pix + (size_t)p * NMAX;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    sp->un[i] = P5[i]; sp->vn[i] = P5[NMAX + i]; sp->x[i] = P5[2 * NMAX + i]; sp->y[i] = P5[3 * NMAX + i];
    sp->z[i] = P5[4 * NMAX + i];
    sp->pix[i] = pg[i];
  }
  return smem + 5 * NMAX * 8 + NMAX * 2;
}

// Inlier test of EPOSScoringFunction::getScore (scoring_function.h:236-242): r^2 < T with
// r^2 = (px/pz - un)^2 + (py/pz - vn)^2.  Evaluated without the divisions as
// (px - un pz)^2 + (py - vn pz)^2 < T pz^2, which is the same predicate for every finite pz != 0 (pz = 0 or a NaN
// gives "outlier" in both forms).
__device__ __forceinline__ bool is_inlier(double un, double vn, double x, double y, double z, const double* m, double T) {
  const double px = fma(m[0], x, fma(m[1], y, fma(m[2], z, m[3])));
  const double py = fma(m[4], x, fma(m[5], y, fma(m[6], z, m[7])));
  const double pz = fma(m[8], x, fma(m[9], y, fma(m[10], z, m[11])));
  const double a = fma(-un, pz, px), b = fma(-vn, pz, py);
  return fma(a, a, b * b) < T * (pz * pz);
}

// EPOSScoringFunction::getScore (scoring_function.h:220-267) by one warp: inlier count by ballot, distinct pixels
// through a per-warp bitset.  bits: NMAX/32 words owned by this warp.  Four points per lane are in flight.
__device__ inline void score_warp(const SmemPoints& sp, int N, const double* model, double sq_trunc, unsigned int* bits,
                                  int lane, int* inl_out, int* pix_out) {
  double m[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) m[k] = model[k];
  for (int k = lane; k < NMAX / 32; k += 32) bits[k] = 0u;
  __syncwarp();
  int inl = 0;
  for (int base = 0; base < N; base += 128) {
    bool in[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * 32 + lane;
      in[u] = false;
      if (i < N) in[u] = is_inlier(sp.un[i], sp.vn[i], sp.x[i], sp.y[i], sp.z[i], m, sq_trunc);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (in[u]) {
        const unsigned int pid = sp.pix[base + u * 32 + lane];
        atomicOr(&bits[pid >> 5], 1u << (pid & 31));
      }
      inl += __popc(__ballot_sync(0xffffffffu, in[u]));
    }
  }
  __syncwarp();
  int px = 0;
  for (int k = lane; k < NMAX / 32; k += 32) px += __popc(bits[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) px += __shfl_xor_sync(0xffffffffu, px, o);
  *inl_out = inl;
  *pix_out = px;
}

// thread 0 only: end of graphCutLocalOptimization (GCRANSAC.h:799-808) + the caller's bookkeeping (:418-427)
__device__ inline void finalize_lo(ProbState* st, const epos_fit_params& prm) {
  if (st->best_value < st->lo_value) {
    st->best_value = st->lo_value;
    st->best_inl = st->lo_inl;
    for (int i = 0; i < 12; ++i) st->best_model[i] = st->lo_model[i];
  }
  st->lo_stage = 0;
  if (st->lo_final) {
    st->phase = PH_FINAL;
  } else {
    st->max_iteration = iteration_bound(1.0, st->best_inl, st->N);
    st->coverage = (double)st->best_value / (double)st->used_pixels;
    st->phase = PH_MAIN;
  }
}

// =====================================================================================================
// main: one warp per RANSAC pass, in-order replay
// =====================================================================================================
struct PassRecord {
  double models[48];
  int nm, fails;
  int inl[4], pix[4];
};

constexpr int CHUNK = 80;            // RANSAC passes evaluated per chunk (a multiple of the 20 warps) (records survive LO rounds in the workspace)

__device__ __noinline__ void phase_main(const Workspace& ws, const epos_fit_params& prm, int p, unsigned char* smem_raw,
                                        bool& pts_loaded) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ProbState* st = ws.st + p;
  const int N = st->N;
  SmemPoints sp;
  unsigned char* rest = load_points(smem_raw, ws, p, N, &sp, pts_loaded);
  unsigned int* bits = reinterpret_cast<unsigned int*>(rest) + warp * (NMAX / 32);
  double* best_model = reinterpret_cast<double*>(rest + WARPS * (NMAX / 32) * 4);
  PassRecord* recs = ws.recs + (size_t)p * CHUNK;                  // global (L2): hypotheses do not depend on the state
  // state (identical in every thread)
  unsigned long long iter = st->iter, max_iteration = st->max_iteration;
  const unsigned long long seed = st->seed;
  int pass = st->pass, best_value = st->best_value, best_inl = st->best_inl, lo_runs = st->lo_runs, gc_count = st->gc_count;
  int chunk_base = st->chunk_base, chunk_n = st->chunk_n;
  double coverage = st->coverage;
  const int used_pixels = st->used_pixels;
  const double sq_trunc = st->sq_trunc;
  const unsigned long long max_iters = (unsigned long long)prm.max_iters, min_iters = (unsigned long long)prm.min_iters;
  if (tid < 12) best_model[tid] = st->best_model[tid];
  __syncthreads();
  bool ended = false, to_lo = false;
  long long t_sample = 0, t_score = 0, t_replay = 0, n_scored = 0;
  const long long t_begin = clock64();
  while (true) {
    long long t0 = clock64();
    if (pass >= chunk_base + chunk_n) {
      // ---- phase A: one THREAD per pass: sample until a valid sample gives >= 1 admissible P3P pose ----
      chunk_base = pass; chunk_n = CHUNK;
      if (tid < CHUNK) {
        PassRecord* rc = recs + tid;
        const int my_pass = chunk_base + tid;
        int fails = -1, nm = 0;
        const double* PU = ws.pts + (size_t)p * 7 * NMAX + 5 * NMAX;   // original pixel coordinates (u row, v row)
        double models[48];
        while (++fails < prm.max_unsuccessful) {
          int s[3];
          if (!unique_set(seed, 0, (u64)my_pass, (u64)fails, N, 3, s)) continue;
          // isValidSample (perspective_n_point_estimator.h:172-198): pixel-space triangle area > min_triangle_area
          const double u0 = PU[s[0]], v0 = PU[NMAX + s[0]];
          const double area = 0.5 * fabs((PU[s[1]] - u0) * (PU[NMAX + s[2]] - v0) - (PU[s[2]] - u0) * (PU[NMAX + s[1]] - v0));
          if (!(area > prm.min_triangle_area)) continue;
          double un[3], vn[3], X[3][3];
          for (int k = 0; k < 3; ++k) {
            un[k] = sp.un[s[k]]; vn[k] = sp.vn[s[k]];
            X[k][0] = sp.x[s[k]]; X[k][1] = sp.y[s[k]]; X[k][2] = sp.z[s[k]];
          }
          nm = p3p_kneip(un, vn, X, models);
          if (nm > 0) break;
        }
        rc->nm = nm; rc->fails = fails;
        for (int k = 0; k < 12 * nm; ++k) rc->models[k] = models[k];
      }
      __syncthreads();
      t_sample += clock64() - t0; t0 = clock64();
      // ---- phase B: one WARP per pass (round-robin): score every solution over all N points ----
      for (int k = warp; k < CHUNK; k += WARPS) {
        PassRecord* rc = recs + k;
        const int nm = rc->nm;
        for (int m = 0; m < nm; ++m) {
          int inl, px;
          score_warp(sp, N, rc->models + 12 * m, sq_trunc, bits, lane, &inl, &px);
          if (lane == 0) { rc->inl[m] = inl; rc->pix[m] = px; }
        }
      }
      __syncthreads();
      if (tid == 0) { int sc = 0; for (int k = 0; k < CHUNK; ++k) sc += recs[k].nm; n_scored += sc; }
      t_score += clock64() - t0; t0 = clock64();
    }
    // ---- in-order replay (every thread runs the same scalar code on the same records) ----
    for (int k = pass - chunk_base; k < chunk_n; ++k) {
      const unsigned long long lim = max_iteration < max_iters ? max_iteration : max_iters;
      if (!(min_iters > iter || iter < lim)) { ended = true; break; }
      if (min_iters < iter) {
        if (iter > max_iteration || iter > max_iters || prm.min_coverage < coverage) { ended = true; break; }
      }
      bool do_lo = false;
      ++iter;
      const PassRecord* rc = recs + k;
      iter += (unsigned long long)rc->fails;
      const int nm = rc->nm;
      for (int m = 0; m < nm; ++m) {
        int s_inl = rc->inl[m], s_val = rc->pix[m];
        if (s_inl + 1 < best_inl) { s_inl = 0; s_val = 0; }       // early-out of getScore, scoring_function.h:257-259
        if (best_value < s_val) {
          best_value = s_val; best_inl = s_inl;
          __syncthreads();
          if (tid < 12) best_model[tid] = rc->models[12 * m + tid];
          __syncthreads();
          do_lo = iter > (unsigned long long)prm.min_iters_before_lo && best_inl > 3;
          max_iteration = iteration_bound(1.0, best_inl, N);
          coverage = (double)best_value / (double)used_pixels;
        }
      }
      ++pass;
      if (do_lo) {
        lo_runs += 2;                                               // GCRANSAC.h:409 and :702
        if (gc_count + 1 < prm.max_graph_cuts) { to_lo = true; break; }
        ++gc_count;                                                 // budget exhausted: the while at :710 exits at once
      }
    }
    __syncthreads();
    t_replay += clock64() - t0;
    if (ended || to_lo) break;
  }
  if (tid == 0) {
    st->t_sample += t_sample; st->t_score += t_score; st->t_replay += t_replay; st->t_total += clock64() - t_begin;
    st->n_scored_main += n_scored;
    st->iter = iter; st->max_iteration = max_iteration; st->pass = pass; st->best_value = best_value;
    st->best_inl = best_inl; st->coverage = coverage; st->chunk_base = chunk_base; st->chunk_n = chunk_n;
    for (int i = 0; i < 12; ++i) st->best_model[i] = best_model[i];
    bool start_lo = to_lo;
    if (ended) {
      if (best_inl <= 3) {
        st->phase = PH_DONE; st->found = 0;
      } else if (lo_runs == 0) {                                    // final LO if none ran (GCRANSAC.h:453-466)
        lo_runs += 2;
        if (gc_count + 1 < prm.max_graph_cuts) { start_lo = true; st->lo_final = 1; }
        else { ++gc_count; st->phase = PH_FINAL; }
      } else {
        st->phase = PH_FINAL;
      }
    }
    if (start_lo) {
      st->phase = PH_LO; st->lo_stage = 0;
      if (!ended) st->lo_final = 0;
      st->lo_value = best_value; st->lo_inl = best_inl;
      for (int i = 0; i < 12; ++i) st->lo_model[i] = best_model[i];
    }
    st->lo_runs = lo_runs; st->gc_count = gc_count;
  }
}

// =====================================================================================================
// cut: graph-cut labeling of lo_model
// =====================================================================================================
__device__ __noinline__ void phase_cut(const Workspace& ws, const epos_fit_params& prm, int p, unsigned char* smem_raw) {
  __shared__ int s_qn, s_any;
  __shared__ int scan_sh[WARPS + 1];
  const int tid = threadIdx.x;
  ProbState* st = ws.st + p;
  const int gc0 = st->gc_count;
  __syncthreads();
  if (gc0 + 1 >= prm.max_graph_cuts) {                             // while (++graph_cut_number < max) fails
    if (tid == 0) { ++st->gc_count; finalize_lo(st, prm); }
    return;
  }
  const long long t_begin = clock64();
  const int N = st->N;
  const double* P5 = ws.pts + (size_t)p * 7 * NMAX;
  const short* nbr = ws.nbr + (size_t)p * NMAX * MAXNB;
  const unsigned char* owned = ws.owned + (size_t)p * NMAX * MAXNB;
  const int* rev_off = ws.rev_off + (size_t)p * (NMAX + 1);
  const int* rev_idx = ws.rev_idx + (size_t)p * NMAX * MAXNB;
  double* capf = ws.capf + (size_t)p * NMAX * MAXNB;
  int* lstart = ws.lstart + (size_t)p * (NMAX + 2);
  const int KN = prm.max_neighbors < MAXNB ? prm.max_neighbors : MAXNB;   // per-node stride of flow / capf
  // shared-memory carve-up by access frequency; what does not fit stays in the (L2-resident) workspace
  size_t used = 0;
  auto carve = [&](size_t bytes, void* fallback) -> void* {
    bytes = (bytes + 15) & ~(size_t)15;
    if (used + bytes <= SMEM_CUT_DYN) { void* r = smem_raw + used; used += bytes; return r; }
    return fallback;
  };
  int* dist = (int*)carve((size_t)N * 4, ws.dist + (size_t)p * NMAX);
  unsigned short* order = (unsigned short*)carve((size_t)N * 2, ws.order + (size_t)p * NMAX);
  double* exc = (double*)carve((size_t)N * 8, ws.exc + (size_t)p * NMAX);
  double* flow = (double*)carve((size_t)N * KN * 8, ws.flow + (size_t)p * NMAX * MAXNB);
  double* dd = (double*)carve((size_t)N * 8, ws.dd + (size_t)p * NMAX);
  const double lambda = prm.spatial_coherence_weight, oml = 1.0 - lambda, T = st->sq_trunc;
  double model[12];
  for (int i = 0; i < 12; ++i) model[i] = st->lo_model[i];
  // ---- unary terms (GCRANSAC.h:843-860) ----
  for (int i = tid; i < N; i += THREADS) {
    const double r2 = sq_residual(P5[i], P5[NMAX + i], P5[2 * NMAX + i], P5[3 * NMAX + i], P5[4 * NMAX + i], model);
    const double qd = r2 / T;
    double d = qd < 0.0 ? 0.0 : (qd > 1.0 ? 1.0 : qd);
    if (!(qd == qd)) d = 0.0;
    dd[i] = d;
    const double e = 1.0 - d;
    double u0, u1;
    if (r2 <= T) { u0 = oml * e; u1 = 0.0; } else { u0 = 0.0; u1 = oml * (1.0 - e); }
    exc[i] = u1 - u0;
  }
  __syncthreads();
  // ---- pairwise terms (energy.h:217-253): tr[x] -= A, arcs x->y cap lambda - A, y->x cap lambda ----
  for (int i = tid; i < N; i += THREADS) {
    double tr = exc[i];
    for (int k = 0; k < KN; ++k) {
      if (!owned[(size_t)i * MAXNB + k]) continue;
      const int j = nbr[(size_t)i * MAXNB + k];
      const double e00 = 0.5 * (dd[i] + dd[j]);
      const double A = e00 * lambda;
      tr += 0.0 - A;
      capf[(size_t)i * KN + k] = lambda - A;
      flow[(size_t)i * KN + k] = 0.0;
    }
    exc[i] = tr;
  }
  __syncthreads();
  // ---- preflow-push in waves.  Each round: reverse BFS from the nodes that still have sink capacity (exc < 0)
  // over residual arcs (frontier queue, work ~ reached set), then one push wave from the farthest level down. ----
  // The BFS depth is capped (iterative deepening): excess far from every sink can only matter after the arcs next to
  // the sinks have carried flow, so shallow rounds come first; a round proves termination only if its BFS ran dry.
  int depth_cap = 2;
  for (int round = 0; round < 4096; ++round) {
    if (tid == 0) { s_qn = 0; s_any = 0; }
    __syncthreads();
    for (int i = tid; i < N; i += THREADS) {
      const bool root = exc[i] < -CUT_EPS;
      dist[i] = root ? 0 : DIST_INF;
      if (root) order[atomicAdd(&s_qn, 1)] = (unsigned short)i;
    }
    __syncthreads();
    int level = 0, begin = 0, end = s_qn;
    if (tid == 0) lstart[0] = 0;
    while (begin < end && level < depth_cap) {
      for (int q = begin + tid; q < end; q += THREADS) {
        const int v = order[q];
        for (int k = 0; k < KN; ++k)                                 // edges owned by v: arc u -> v has residual lambda + flow
          if (owned[(size_t)v * MAXNB + k]) {
            const int u = nbr[(size_t)v * MAXNB + k];
            if (dist[u] == DIST_INF && lambda + flow[(size_t)v * KN + k] > CUT_EPS &&
                atomicCAS(&dist[u], DIST_INF, level + 1) == DIST_INF) {
              order[atomicAdd(&s_qn, 1)] = (unsigned short)u;
              if (exc[u] > 0.0) s_any = 1;
            }
          }
        for (int a = rev_off[v]; a < rev_off[v + 1]; ++a) {           // edges owned by u that end in v: arc u -> v forward
          const int e0 = rev_idx[a];
          const int u = e0 / MAXNB;
          const size_t e = (size_t)u * KN + (e0 % MAXNB);
          if (dist[u] == DIST_INF && capf[e] - flow[e] > CUT_EPS && atomicCAS(&dist[u], DIST_INF, level + 1) == DIST_INF) {
            order[atomicAdd(&s_qn, 1)] = (unsigned short)u;
            if (exc[u] > 0.0) s_any = 1;
          }
        }
      }
      __syncthreads();
      begin = end; end = s_qn; ++level;
      if (tid == 0) lstart[level] = begin;
      __syncthreads();
    }
    const bool exhausted = begin >= end;                              // the BFS ran dry: dist is exact for every node
    const int maxlevel = exhausted ? level - 1 : level;              // deepest labelled level
    if (!exhausted && tid == 0) lstart[level + 1] = end;
    __syncthreads();
    if (!s_any) {
      if (exhausted) break;
      depth_cap *= 4;                                                 // nothing to push within the cap: look deeper
      continue;
    }
    for (int d = maxlevel; d >= 1; --d) {
      const int qb = lstart[d], qe = lstart[d + 1];
      for (int q = qb + tid; q < qe; q += THREADS) {
        const int i = order[q];
        double ex = exc[i];
        if (!(ex > 0.0)) continue;
        for (int k = 0; k < KN && ex > 0.0; ++k)
          if (owned[(size_t)i * MAXNB + k]) {
            const int j = nbr[(size_t)i * MAXNB + k];
            if (dist[j] != d - 1) continue;
            const size_t e = (size_t)i * KN + k;
            const double res = capf[e] - flow[e];
            if (!(res > 0.0)) continue;
            if (ex >= res) { flow[e] = capf[e]; atomicAdd(&exc[j], res); ex -= res; }      // saturating: exact bound
            else { flow[e] += ex; atomicAdd(&exc[j], ex); ex = 0.0; }
          }
        for (int a = rev_off[i]; a < rev_off[i + 1] && ex > 0.0; ++a) {
          const int e0 = rev_idx[a];
          const int x = e0 / MAXNB;
          if (dist[x] != d - 1) continue;
          const size_t e = (size_t)x * KN + (e0 % MAXNB);
          const double res = lambda + flow[e];
          if (!(res > 0.0)) continue;
          if (ex >= res) { flow[e] = -lambda; atomicAdd(&exc[x], res); ex -= res; }
          else { flow[e] -= ex; atomicAdd(&exc[x], ex); ex = 0.0; }
        }
        exc[i] = ex;
      }
      __syncthreads();
    }
  }
  // ---- inliers = SINK segment = nodes with a residual path to a node that still has sink capacity ----
  unsigned short* inl = ws.inl + (size_t)p * NMAX;
  {
    const int per = PER_THREAD;
    int cnt = 0;
    for (int k = 0; k < per; ++k) { const int i = tid * per + k; cnt += (i < N && dist[i] != DIST_INF); }
    int total;
    int base = block_excl_scan<WARPS>(cnt, scan_sh, &total);
    for (int k = 0; k < per; ++k) { const int i = tid * per + k; if (i < N && dist[i] != DIST_INF) inl[base++] = (unsigned short)i; }
    if (tid == 0) { st->ni = total; ++st->gc_count; st->lo_stage = 1; st->t_cut += clock64() - t_begin; }
  }
}

// =====================================================================================================
// trials: inner RANSAC of the local optimisation, one warp per trial
// =====================================================================================================
struct TrialRecord { double model[12]; int ok, inl, pix, pad; };

__device__ __noinline__ void phase_trials(const Workspace& ws, const epos_fit_params& prm, int p, unsigned char* smem_raw,
                                          bool& pts_loaded) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int TW = TRIAL_THREADS / 32;
  ProbState* st = ws.st + p;
  const long long t_begin = clock64();
  long long t_fit = 0;
  const int N = st->N, ni = st->ni;
  SmemPoints sp;
  unsigned char* rest = load_points(smem_raw, ws, p, N, &sp, pts_loaded);
  TrialRecord* recs = reinterpret_cast<TrialRecord*>(rest);
  rest += MAX_TRIALS * sizeof(TrialRecord);
  double* fit_sh = reinterpret_cast<double*>(rest) + warp * FIT_SCRATCH_DOUBLES;
  unsigned int* bits = reinterpret_cast<unsigned int*>(fit_sh);        // the fit scratch is idle while scoring
  rest += TW * FIT_SCRATCH_DOUBLES * 8;
  unsigned short* sample = reinterpret_cast<unsigned short*>(rest) + warp * 32;
  const unsigned short* inl = ws.inl + (size_t)p * NMAX;
  const int trials = prm.max_lo_trials < MAX_TRIALS ? prm.max_lo_trials : MAX_TRIALS;
  const int sample_size = ni < 21 ? ni : 21;
  const double sq_trunc = st->sq_trunc;
  const int gc = st->gc_count;
  for (int t = tid; t < MAX_TRIALS; t += TRIAL_THREADS) recs[t].ok = 0;
  __syncthreads();
  WarpGroup g;
  g.rank = lane; g.size = 32; g.sh = fit_sh;
  int n_eval = 0;                      // trials actually evaluated (the all-inlier case repeats one model)
  if (sample_size < ni) n_eval = trials;
  else if (3 < ni) n_eval = 1;
  for (int t = warp; t < n_eval; t += TW) {
    if (sample_size < ni) {
      int sel[21];
      unique_set(st->seed, 1, (u64)gc, (u64)t, ni, sample_size, sel);
      if (lane < sample_size) sample[lane] = inl[sel[lane]];
    } else {
      if (lane < sample_size) sample[lane] = inl[lane];
    }
    __syncwarp();
    PointView pv;
    pv.un = sp.un; pv.vn = sp.vn; pv.x = sp.x; pv.y = sp.y; pv.z = sp.z; pv.idx = sample; pv.n = sample_size;
    double model[12];
    const long long tf0 = clock64();
    const bool ok = fit_nonminimal_group(g, pv, model);
    __syncwarp();
    t_fit += clock64() - tf0;
    if (ok) {
      TrialRecord* rc = recs + t;
      if (lane < 12) rc->model[lane] = model[lane];
      __syncwarp();
      int in_, px;
      score_warp(sp, N, rc->model, sq_trunc, bits, lane, &in_, &px);
      if (lane == 0) { rc->ok = 1; rc->inl = in_; rc->pix = px; }
    }
    __syncwarp();
  }
  __syncthreads();
  if (tid == 0) {
    st->t_fit += t_fit;
    for (int t = 0; t < n_eval; ++t) st->n_scored_lo += recs[t].ok ? 1 : 0;
    bool updated = false;
    int mv = st->lo_value, mi = st->lo_inl;
    if (sample_size < ni) {
      for (int t = 0; t < trials; ++t) {
        const TrialRecord* rc = recs + t;
        if (!rc->ok) continue;                                      // failed fit: `continue` (GCRANSAC.h:748-752)
        int s_inl = rc->inl, s_val = rc->pix;
        if (s_inl + 1 < mi) { s_inl = 0; s_val = 0; }
        if (mv < s_val) { updated = true; mv = s_val; mi = s_inl; for (int i = 0; i < 12; ++i) st->lo_model[i] = rc->model[i]; }
      }
    } else if (3 < ni) {
      const TrialRecord* rc = recs;                                 // identical model in every trial: first one decides
      if (rc->ok) {
        int s_inl = rc->inl, s_val = rc->pix;
        if (s_inl + 1 < mi) { s_inl = 0; s_val = 0; }
        if (mv < s_val) { updated = true; mv = s_val; mi = s_inl; for (int i = 0; i < 12; ++i) st->lo_model[i] = rc->model[i]; }
      }
    }
    st->lo_value = mv; st->lo_inl = mi;
    if (gc >= 1 && gc <= TRACE_ROUNDS) {
      int* tr = ws.trace + ((size_t)p * TRACE_ROUNDS + (gc - 1)) * TRACE_COLS;
      tr[0] = gc; tr[1] = ni; tr[2] = updated ? 1 : 0; tr[3] = mv; tr[4] = mi;
      for (int t = 0; t < MAX_TRIALS; ++t) {
        const bool have = t < n_eval;
        tr[5 + 3 * t] = have ? recs[t].ok : -1; tr[6 + 3 * t] = have && recs[t].ok ? recs[t].inl : 0;
        tr[7 + 3 * t] = have && recs[t].ok ? recs[t].pix : 0;
      }
    }
    if (updated) st->lo_stage = 0;                                  // another labeling round (GCRANSAC.h:794-796)
    else finalize_lo(st, prm);
    st->t_trials += clock64() - t_begin;
  }
}

// =====================================================================================================
// final: iterated least squares, final fit, LM refinement, outputs
// =====================================================================================================
// CTA-wide score with inlier list (ordered): returns inlier count, distinct pixels; list written to `out`.
__device__ inline void score_cta(const SmemPoints& sp, int N, const double* model, double sq_trunc, unsigned int* bits,
                                 int* scan_sh, unsigned short* out, int* inl_out, int* pix_out, int* red_sh) {
  const int tid = threadIdx.x;
  for (int k = tid; k < NMAX / 32; k += THREADS) bits[k] = 0u;
  __syncthreads();
  const int per = PER_THREAD;
  unsigned int mask = 0;
  int cnt = 0;
  for (int k = 0; k < per; ++k) {
    const int i = tid * per + k;
    if (i < N) {
      if (is_inlier(sp.un[i], sp.vn[i], sp.x[i], sp.y[i], sp.z[i], model, sq_trunc)) {
        mask |= 1u << k; ++cnt;
        const unsigned int pid = sp.pix[i];
        atomicOr(&bits[pid >> 5], 1u << (pid & 31));
      }
    }
  }
  int total;
  int base = block_excl_scan<WARPS>(cnt, scan_sh, &total);
  if (out)
    for (int k = 0; k < per; ++k)
      if (mask & (1u << k)) out[base++] = (unsigned short)(tid * per + k);
  if (tid == 0) *red_sh = 0;
  __syncthreads();
  int px = 0;
  for (int k = tid; k < NMAX / 32; k += THREADS) px += __popc(bits[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) px += __shfl_xor_sync(0xffffffffu, px, o);
  if ((tid & 31) == 0 && px) atomicAdd(red_sh, px);
  __syncthreads();
  *inl_out = total;
  *pix_out = *red_sh;
  __syncthreads();
}

__device__ __noinline__ void phase_final(const Workspace& ws, const epos_fit_params& prm, int p, unsigned char* smem_raw,
                                         bool& pts_loaded, const int* __restrict__ offsets, double* __restrict__ poses,
                                         int* __restrict__ labeling) {
  __shared__ int scan_sh[WARPS + 1];
  __shared__ int red_sh;
  const int tid = threadIdx.x;
  ProbState* st = ws.st + p;
  double* rec = poses + (size_t)p * EPOS_POSE_RECORD_DOUBLES;
  if (st->phase != PH_FINAL) {
    if (tid == 0 && st->valid) { rec[13] = (double)st->iter; rec[15] = (double)st->gc_count; }
    if (tid == 0 && st->err) rec[14] = -1.0;                      // more than NMAX correspondences
    return;
  }
  const long long t_begin = clock64();
  const int N = st->N;
  SmemPoints sp;
  unsigned char* rest = load_points(smem_raw, ws, p, N, &sp, pts_loaded);
  unsigned int* bits = reinterpret_cast<unsigned int*>(rest); rest += (NMAX / 32) * 4;
  unsigned short* listA = reinterpret_cast<unsigned short*>(rest); rest += NMAX * 2;
  unsigned short* listB = reinterpret_cast<unsigned short*>(rest); rest += NMAX * 2;
  unsigned short* listC = reinterpret_cast<unsigned short*>(rest); rest += NMAX * 2;
  double* fit_sh = reinterpret_cast<double*>(rest); rest += FIT_SCRATCH_DOUBLES * 8;
  double* part = reinterpret_cast<double*>(rest); rest += WARPS * FIT_NRED * 8;
  double* bcast = reinterpret_cast<double*>(rest);                   // 16 doubles: result of a warp-level fit
  __syncthreads();
  CtaGroup g;
  g.rank = tid; g.size = THREADS; g.sh = fit_sh; g.part = part;
  WarpGroup wg;
  wg.rank = tid & 31; wg.size = 32; wg.sh = fit_sh;
  // Non-minimal fit on an index list: small lists (the no-consensus case) are fitted by warp 0 alone, which avoids a
  // block barrier per reduction; large lists use the whole CTA.
  auto fit_list = [&](const unsigned short* list, int n, double* m2) -> bool {
    PointView v;
    v.un = sp.un; v.vn = sp.vn; v.x = sp.x; v.y = sp.y; v.z = sp.z; v.idx = list; v.n = n;
    if (n > WARP_FIT_MAX) return fit_nonminimal_group(g, v, m2);
    __syncthreads();
    if (tid < 32) {
      double mm[12];
      const bool ok = fit_nonminimal_group(wg, v, mm);
      if (tid == 0) { bcast[12] = ok ? 1.0 : 0.0; for (int i = 0; i < 12; ++i) bcast[i] = ok ? mm[i] : 0.0; }
    }
    __syncthreads();
    for (int i = 0; i < 12; ++i) m2[i] = bcast[i];
    const bool ok = bcast[12] != 0.0;
    __syncthreads();
    return ok;
  };
  const double sq_trunc = st->sq_trunc;
  double best_model[12];
  for (int i = 0; i < 12; ++i) best_model[i] = st->best_model[i];
  const int best_value = st->best_value;
  int nA, pxA;
  int n_scored = 1;
  score_cta(sp, N, best_model, sq_trunc, bits, scan_sh, listA, &nA, &pxA, &red_sh);      // GCRANSAC.h:470-478
  PointView pv;
  pv.un = sp.un; pv.vn = sp.vn; pv.x = sp.x; pv.y = sp.y; pv.z = sp.z;
  bool refit_applied = false;
  if (nA > 3) {                                                                          // GCRANSAC.h:533-657
    double cur[12];
    for (int i = 0; i < 12; ++i) cur[i] = best_model[i];
    for (int i = tid; i < nA; i += THREADS) listB[i] = listA[i];
    __syncthreads();
    int nB = nA, iterations = 0;
    while (++iterations < prm.max_lsq_iters) {
      double m2[12];
      if (!fit_list(listB, nB, m2)) break;
      int nC, pxC;
      score_cta(sp, N, m2, sq_trunc, bits, scan_sh, listC, &nC, &pxC, &red_sh);
      ++n_scored;
      if (nC < 3) break;
      if (nC <= nB) break;
      for (int i = 0; i < 12; ++i) cur[i] = m2[i];
      for (int i = tid; i < nC; i += THREADS) listB[i] = listC[i];
      __syncthreads();
      nB = nC;
    }
    if (iterations > 1) {
      int nC, pxC;
      score_cta(sp, N, cur, sq_trunc, bits, scan_sh, listC, &nC, &pxC, &red_sh);
      ++n_scored;
      if (best_value < pxC) {
        refit_applied = true;
        for (int i = 0; i < 12; ++i) best_model[i] = cur[i];
        for (int i = tid; i < nC; i += THREADS) listA[i] = listC[i];
        __syncthreads();
        nA = nC;
      }
    }
  }
  if (!refit_applied) {                                                                  // GCRANSAC.h:510-521
    double m2[12];
    if (fit_list(listA, nA, m2))
      for (int i = 0; i < 12; ++i) best_model[i] = m2[i];
  }
  if (prm.apply_numerical_optimization && nA >= 6) {                                     // progressivex_python.cpp:257-312
    const double R[9] = {best_model[0], best_model[1], best_model[2], best_model[4], best_model[5], best_model[6],
                         best_model[8], best_model[9], best_model[10]};
    double param[6];
    matrix_to_rodrigues(R, param);
    param[3] = best_model[3]; param[4] = best_model[7]; param[5] = best_model[11];
    pv.idx = listA; pv.n = nA;
    if (nA > WARP_FIT_MAX) {
      lm_refine_group(g, pv, param);
    } else {
      __syncthreads();
      if (tid < 32) {
        lm_refine_group(wg, pv, param);
        if (tid == 0) for (int i = 0; i < 6; ++i) bcast[i] = param[i];
      }
      __syncthreads();
      for (int i = 0; i < 6; ++i) param[i] = bcast[i];
      __syncthreads();
    }
    bool fin = true;
    for (int i = 0; i < 6; ++i) fin &= isfinite(param[i]);
    if (fin) {
      double R2[9];
      rodrigues_to_matrix(param, R2, nullptr);
      for (int r = 0; r < 3; ++r) {
        best_model[r * 4] = R2[r * 3]; best_model[r * 4 + 1] = R2[r * 3 + 1]; best_model[r * 4 + 2] = R2[r * 3 + 2];
        best_model[r * 4 + 3] = param[3 + r];
      }
    }
  }
  __syncthreads();
  if (tid < 12) rec[tid] = best_model[tid];
  if (tid == 0) {
    rec[12] = (double)nA; rec[13] = (double)st->iter; rec[14] = 1.0; rec[15] = (double)st->gc_count;
    st->found = 1; st->phase = PH_DONE; st->t_final += clock64() - t_begin; st->n_scored_final += n_scored;
  }
  const int off = offsets[p];
  for (int i = tid; i < nA; i += THREADS) labeling[off + listA[i]] = 1;
}

// One persistent CTA per problem runs the whole state machine; phases re-carve the dynamic shared memory
// (the cut overwrites the point set, which is reloaded from the L2-resident workspace afterwards).
__global__ void __launch_bounds__(THREADS, 1)
fit_kernel(Workspace ws, epos_fit_params prm, const int* __restrict__ offsets, double* __restrict__ poses,
           int* __restrict__ labeling) {
  extern __shared__ unsigned char smem_raw[];
  // Longest-processing-time-first: CTAs are dispatched in blockIdx order and P usually exceeds the SM count (one CTA
  // per SM), so block b takes the problem with the b-th largest point count -- the expensive problems start in the
  // first wave and the small ones fill the tail.  Problems are independent (own seed, own workspace slice), so the
  // mapping does not change any result.
  __shared__ int s_problem;
  {
    const int P = gridDim.x;
    for (int j = threadIdx.x; j < P; j += THREADS) {
      const int nj = ws.st[j].N;
      int rank = 0;
      for (int k = 0; k < P; ++k) {
        const int nk = ws.st[k].N;
        rank += (nk > nj || (nk == nj && k < j)) ? 1 : 0;
      }
      if (rank == (int)blockIdx.x) s_problem = j;
    }
    __syncthreads();
  }
  const int p = s_problem;
  ProbState* st = ws.st + p;
  bool pts_loaded = false;
  for (int guard = 0; guard < 1 << 14; ++guard) {
    __syncthreads();
    const int phase = st->phase, stage = st->lo_stage;
    __syncthreads();
    if (phase == PH_MAIN) {
      phase_main(ws, prm, p, smem_raw, pts_loaded);
    } else if (phase == PH_LO) {
      if (stage == 0) { phase_cut(ws, prm, p, smem_raw); pts_loaded = false; }
      else phase_trials(ws, prm, p, smem_raw, pts_loaded);
    } else {
      break;
    }
  }
  __syncthreads();
  phase_final(ws, prm, p, smem_raw, pts_loaded, offsets, poses, labeling);
}

constexpr size_t SMEM_POINTS = 5 * NMAX * 8 + NMAX * 2;
constexpr size_t SMEM_PREP = 5 * NMAX * 4 + 8192 * 8 + 64 * 4 + 8192 * 2 + 64;
static_assert(sizeof(PassRecord) <= 432 && CHUNK == 80, "workspace_layout reserves 80 x 432 bytes of pass records");
constexpr size_t SMEM_MAIN = SMEM_POINTS + WARPS * (NMAX / 32) * 4 + 12 * 8 + 64;
constexpr size_t SMEM_CUT = SMEM_CUT_DYN;
static_assert(FIT_SCRATCH_DOUBLES * 8 >= (NMAX / 32) * 4, "bitset must fit in the fit scratch");
constexpr size_t SMEM_TRIALS = SMEM_POINTS + MAX_TRIALS * sizeof(TrialRecord) +
                               (TRIAL_THREADS / 32) * FIT_SCRATCH_DOUBLES * 8 + (TRIAL_THREADS / 32) * 32 * 2 + 64;
constexpr size_t SMEM_FINAL = SMEM_POINTS + (NMAX / 32) * 4 + 3 * NMAX * 2 + FIT_SCRATCH_DOUBLES * 8 + WARPS * FIT_NRED * 8 + 16 * 8 + 64;
constexpr size_t cmax(size_t a, size_t b) { return a > b ? a : b; }
constexpr size_t SMEM_FIT = cmax(cmax(SMEM_MAIN, SMEM_CUT), cmax(SMEM_TRIALS, SMEM_FINAL));
static_assert(SMEM_FIT + 1024 <= 227 * 1024, "fit kernel shared memory");

}  // namespace pose
}  // namespace epos

using namespace epos;
using namespace epos::pose;

extern "C" {

void epos_fit_params_default(epos_fit_params* p) {
  if (!p) return;
  p->threshold = 4.0; p->spatial_coherence_weight = 0.1; p->neighborhood_ball_radius = 20.0;
  p->scaling_from_millimeters = 0.1; p->min_triangle_area = 0.0; p->min_coverage = 0.5;
  p->max_iters = 400; p->min_iters = 10; p->min_iters_before_lo = 20; p->max_lo_trials = 20;
  p->max_graph_cuts = 10; p->max_lsq_iters = 10; p->max_unsuccessful = 100; p->max_neighbors = 5;
  p->apply_numerical_optimization = 1; p->reserved = 0;
}

int epos_fit_max_points(void) { return NMAX; }

// Profiling / debugging aid: copies per-problem counters out of a workspace after epos_fit_poses (synchronises).
// out [P][EPOS_FIT_DEBUG_COLS] i64: N, used_pixels, iterations, passes, graph_cuts, lo_runs, phase, best_inliers,
//                  clocks (sample+P3P, scoring, replay, main total, cut, trials, final, trial fits),
//                  models scored over all N points in main / LO trials / final, reserved.
int epos_fit_debug_state(const void* workspace, int P, long long* out) {
  EPOS_CHECK_ARG(workspace && out && P > 0);
  Workspace ws;
  workspace_layout(P, const_cast<void*>(workspace), &ws);
  ProbState* h = (ProbState*)malloc((size_t)P * sizeof(ProbState));
  if (!h) return EPOS_ERR_CUDA;
  cudaError_t e = cudaMemcpy(h, ws.st, (size_t)P * sizeof(ProbState), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { free(h); set_error("epos_fit_debug_state: %s", cudaGetErrorString(e)); return EPOS_ERR_CUDA; }
  for (int i = 0; i < P; ++i) {
    long long* o = out + (size_t)i * EPOS_FIT_DEBUG_COLS;
    o[0] = h[i].N; o[1] = h[i].used_pixels; o[2] = (long long)h[i].iter; o[3] = h[i].pass; o[4] = h[i].gc_count;
    o[5] = h[i].lo_runs; o[6] = h[i].phase; o[7] = h[i].best_inl; o[8] = h[i].t_sample; o[9] = h[i].t_score;
    o[10] = h[i].t_replay; o[11] = h[i].t_total; o[12] = h[i].t_cut; o[13] = h[i].t_trials; o[14] = h[i].t_final;
    o[15] = h[i].t_fit;
    o[16] = h[i].n_scored_main; o[17] = h[i].n_scored_lo; o[18] = h[i].n_scored_final; o[19] = 0;
  }
  free(h);
  return EPOS_OK;
}

// Debugging aid (synchronous): the local-optimisation rounds of problem p of the last epos_fit_poses:
// out [16][72] i32 rows (graph-cut number, labelled inliers, updated, lo value, lo inliers, then (ok, inliers, pixels)
// of the 20 inner fits; ok = -1: trial not evaluated).  Rows of rounds that did not run keep stale data.
int epos_fit_debug_trace(const void* workspace, int P, int p, int32_t* out) {
  EPOS_CHECK_ARG(workspace && out && P > 0 && p >= 0 && p < P);
  Workspace ws;
  workspace_layout(P, const_cast<void*>(workspace), &ws);
  EPOS_CUDA(cudaMemcpy(out, ws.trace + (size_t)p * TRACE_ROUNDS * TRACE_COLS, (size_t)TRACE_ROUNDS * TRACE_COLS * 4,
                       cudaMemcpyDeviceToHost));
  return EPOS_OK;
}

size_t epos_fit_workspace_bytes(int P, int max_points, const epos_fit_params* params) {
  (void)max_points; (void)params;
  if (P <= 0) return 0;
  return workspace_layout(P, nullptr, nullptr);
}

static int set_smem_attrs() {
  static std::atomic<int> done[EPOS_MAX_DEVICES];
  const int dslot = device_slot();
  if (done[dslot].load(std::memory_order_acquire)) return EPOS_OK;
  EPOS_CUDA(cudaFuncSetAttribute(prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PREP));
  EPOS_CUDA(cudaFuncSetAttribute(fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FIT));
  done[dslot].store(1, std::memory_order_release);
  return EPOS_OK;
}

// Optional per-launch timing of the two kernels with CUDA events on the launch stream (bench.py's RANSAC roofline).
static bool g_fit_timing = false;
static cudaEvent_t g_fit_ev[3] = {nullptr, nullptr, nullptr};

int epos_fit_enable_timing(int on) {
  g_fit_timing = on != 0;
  if (g_fit_timing && !g_fit_ev[0])
    for (int i = 0; i < 3; ++i) EPOS_CUDA(cudaEventCreate(&g_fit_ev[i]));
  return EPOS_OK;
}

int epos_fit_last_kernel_ms(float* prep_ms, float* fit_ms) {
  EPOS_CHECK_ARG(prep_ms && fit_ms);
  if (!g_fit_ev[0]) { set_error("epos_fit_last_kernel_ms: timing was never enabled"); return EPOS_ERR_INVALID_ARG; }
  EPOS_CUDA(cudaEventSynchronize(g_fit_ev[2]));
  EPOS_CUDA(cudaEventElapsedTime(prep_ms, g_fit_ev[0], g_fit_ev[1]));
  EPOS_CUDA(cudaEventElapsedTime(fit_ms, g_fit_ev[1], g_fit_ev[2]));
  return EPOS_OK;
}

int epos_fit_poses(const double* coord_2d, const double* coord_3d, const int32_t* offsets, const int32_t* counts, int P,
                   const double* K, const uint64_t* seeds, const epos_fit_params* params, double* poses,
                   int32_t* labeling, void* workspace, size_t workspace_bytes, void* stream) {
  EPOS_CHECK_ARG(coord_2d && coord_3d && offsets && counts && K && seeds && params && poses && labeling && workspace);
  EPOS_CHECK_ARG(P > 0);
  EPOS_CHECK_ARG(params->max_neighbors >= 0 && params->max_neighbors <= MAXNB);
  EPOS_CHECK_ARG(params->max_lo_trials >= 0 && params->max_lo_trials <= MAX_TRIALS);
  EPOS_CHECK_ARG(params->max_graph_cuts >= 1 && params->max_graph_cuts <= 64);
  EPOS_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  if (workspace_bytes < workspace_layout(P, nullptr, nullptr)) {
    set_error("epos_fit_poses: workspace too small (%zu < %zu)", workspace_bytes, workspace_layout(P, nullptr, nullptr));
    return EPOS_ERR_INVALID_ARG;
  }
  int rc = set_smem_attrs();
  if (rc) return rc;
  Workspace ws;
  workspace_layout(P, workspace, &ws);
  cudaStream_t s = (cudaStream_t)stream;
  if (g_fit_timing) EPOS_CUDA(cudaEventRecord(g_fit_ev[0], s));
  prep_kernel<<<P, PT, SMEM_PREP, s>>>(ws, coord_2d, coord_3d, offsets, counts, K,
                                            reinterpret_cast<const unsigned long long*>(seeds), *params, labeling, poses);
  EPOS_LAUNCH_CHECK();
  if (g_fit_timing) EPOS_CUDA(cudaEventRecord(g_fit_ev[1], s));
  fit_kernel<<<P, THREADS, SMEM_FIT, s>>>(ws, *params, offsets, poses, labeling);
  EPOS_LAUNCH_CHECK();
  if (g_fit_timing) EPOS_CUDA(cudaEventRecord(g_fit_ev[2], s));
  return EPOS_OK;
}

}  // extern "C"
