// Batched single-instance pose fitting on sm_100a: the behaviour of pyprogressivex.find6DPoses
// (/root/reference/external/progressive-x/src/pyprogressivex/src/progressivex_python.cpp:36-134,222-336) for P
// independent (image, object) problems, one CTA per problem.
//
// GC-RANSAC (.../graph-cut-ransac/src/pygcransac/include/GCRANSAC.h:206-530) is sequential by construction: best-so-far
// updates, the LO trigger (iteration > 20 on a new best), the cumulative budget of 9 graph cuts and the coverage exit
// all depend on the order of hypotheses.  The kernels keep that order exactly while evaluating hypotheses in parallel:
//
//   prep_kernel   points (u_n, v_n, x, y, z), dense pixel ids, k-nearest neighbour graph (f32, 5-D), reverse adjacency
//   fit_kernel    ONE PERSISTENT CTA PER PROBLEM runs the whole state machine, re-carving its shared memory per phase:
//     main     up to 160 RANSAC passes at a time (never more than the iteration budget still allows): a thread per pass
//              samples until a valid sample gives an admissible Kneip P3P pose; a warp per pass scores its solutions
//              over all N points, two solutions per sweep (per-lane inlier counts, per-warp pixel bitsets, early exit
//              when neither can reach the early-out bound); then warp 0 replays the chunk in order -- lanes test 32
//              passes at once for "stops the loop / accepts a model", only such a pass runs the scalar update /
//              early-out / LO-trigger / termination code of the reference
//     cut      graph-cut labeling (GCRANSAC.h:812-920): f64 preflow-push in waves + reverse BFS = nodes that can still
//              reach the sink, which is what the reference's BK max-flow labels SINK (graph.h:112-115,478-488)
//     trials   the <= 20 inner fits of graphCutLocalOptimization (GCRANSAC.h:737-792), one warp per trial
//              (sample 21 inliers -> DLT + LM -> score), replayed in order
//     final    iterated least squares, final non-minimal fit, final LM refinement (GCRANSAC.h:480-521,
//              progressivex_python.cpp:257-312), pose record + labeling
//   progx_kernel  the same phases inside the Progressive-X outer loop (multi-instance problems; see further down)
//
// Two launches per batch (prep, fit), nothing is read back: the sequence is CUDA-graph capturable (engine.py).
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
  double best_value, lo_value;      // Score::value is a double (pixel count; minus (shared support)^2 with a compound model)
  int best_inl, lo_inl;
  // Progressive-X proposal engine only (single-instance problems keep cpref = NULL, stream_base = 0): the compound
  // model's preference vector and the random-stream pair (stream_base, stream_base + 1) of the current proposal
  const double* cpref;
  unsigned long long stream_base;
  int multi, n_final;               // multi = 1: phase_final keeps (model, inlier list) in the workspace, no final LM
  int lo_runs, gc_count, lo_final, lo_stage;
  int ni, found, err, pad;
  int chunk_base, chunk_n;
  long long t_sample, t_score, t_replay, t_total;   // main_kernel clock64() accumulators (profiling aid)
  long long t_cut, t_trials, t_final, t_fit;        // other kernels; t_fit = non-minimal fits inside trials (warp 0)
  long long n_scored_main, n_scored_lo, n_scored_final;   // models scored over all N points (roofline accounting)
  long long pts_skipped;                                  // points the early-out of the main loop did not visit
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
  // ---- Progressive-X (multi-instance problems only; NULL otherwise) ----
  struct MultiState* ms;    // [P]
  double* fin_model;        // [P][12]            model of the last proposal (phase_final, multi = 1)
  unsigned short* fin_inl;  // [P][NMAX]          its inlier list (st->n_final entries)
  double* models;           // [P][MAX_INSTANCES][12]
  double* pref;             // [P][PEARL_MAX][NMAX]  preference vectors of the instances (as of their acceptance)
  double* compound;         // [P][NMAX]
  double* r2tab;            // [P][PEARL_MAX][NMAX]  squared residuals of the current PEARL iteration
  int* labels;              // [P][NMAX]
  double* cr;               // [P][NMAX*MAXNB]    reverse capacities of the expansion graph (capf = forward)
  double* af;               // [P][NMAX]          flow on the arc aux(label) -> site
  unsigned char* mult;      // [P][NMAX][MAXNB]   multiplicity (1 / 2) of an owned neighbour pair
  unsigned short* lists;    // [P][NMAX]          scratch index list (instance members)
};
constexpr int MAX_INSTANCES = 32;   // instances returned at most ("all instances" mode; PEARL mode keeps <= 5)
constexpr int PEARL_MAX = 6;        // instances PEARL may hold at once (max_model_number_for_optimization <= 5, + the proposal)

struct MultiState {
  long long total_iterations;
  int n_models, unaccepted, first_events, n_outliers, proposals, accepted, pearl_iterations, moves, done;
  int max_models, sped_up, pad;
  double scores[MAX_INSTANCES];
};

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static size_t workspace_layout(int P, void* base, Workspace* w, bool multi = false) {
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
  p = take((size_t)P * 160 * 464); if (w) w->recs = (PassRecord*)p;
  p = take((size_t)P * TRACE_ROUNDS * TRACE_COLS * 4); if (w) w->trace = (int*)p;
  if (w) { w->ms = nullptr; w->fin_model = nullptr; w->fin_inl = nullptr; w->models = nullptr; w->pref = nullptr;
           w->compound = nullptr; w->r2tab = nullptr; w->labels = nullptr; w->cr = nullptr; w->af = nullptr;
           w->mult = nullptr; w->lists = nullptr; }
  if (multi) {
    p = take((size_t)P * sizeof(MultiState)); if (w) w->ms = (MultiState*)p;
    p = take((size_t)P * 12 * 8); if (w) w->fin_model = (double*)p;
    p = take((size_t)P * NMAX * 2); if (w) w->fin_inl = (unsigned short*)p;
    p = take((size_t)P * MAX_INSTANCES * 12 * 8); if (w) w->models = (double*)p;
    p = take((size_t)P * PEARL_MAX * NMAX * 8); if (w) w->pref = (double*)p;
    p = take((size_t)P * NMAX * 8); if (w) w->compound = (double*)p;
    p = take((size_t)P * PEARL_MAX * NMAX * 8); if (w) w->r2tab = (double*)p;
    p = take((size_t)P * NMAX * 4); if (w) w->labels = (int*)p;
    p = take((size_t)P * NMAX * MAXNB * 8); if (w) w->cr = (double*)p;
    p = take((size_t)P * NMAX * 8); if (w) w->af = (double*)p;
    p = take((size_t)P * NMAX * MAXNB); if (w) w->mult = (unsigned char*)p;
    p = take((size_t)P * NMAX * 2); if (w) w->lists = (unsigned short*)p;
  }
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
    st->iter = 0; st->pass = 0; st->best_value = 0.0; st->best_inl = 0; st->coverage = 0.0;
    st->cpref = nullptr; st->stream_base = 0ULL; st->multi = 0; st->n_final = 0;
    st->lo_runs = 0; st->gc_count = 0; st->lo_final = 0; st->lo_stage = 0; st->ni = 0; st->found = 0;
    st->lo_value = 0.0; st->lo_inl = 0; st->used_pixels = 0; st->chunk_base = 0; st->chunk_n = 0;
    st->t_sample = st->t_score = st->t_replay = st->t_total = 0;
    st->t_cut = st->t_trials = st->t_final = st->t_fit = 0;
    st->n_scored_main = st->n_scored_lo = st->n_scored_final = 0; st->pts_skipped = 0;
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
  auto knn = [&](auto kn_tag) {
    // KEEP = entries kept per point: the compile-time neighbour count (default 5) or MAXNB for any other setting.  Only
    // the KN nearest are written, and once KEEP are known a candidate that does not beat the worst of them -- most of
    // them, in the dense many-to-many neighbourhoods -- costs two comparisons instead of a walk over the list.
    constexpr int KEEP = decltype(kn_tag)::value;
    for (int i = tid; i < N; i += PT) {
      float bd[KEEP];
      int bj[KEEP];
#pragma unroll
      for (int k = 0; k < KEEP; ++k) { bd[k] = INFINITY; bj[k] = -1; }
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
          if (bj[KEEP - 1] >= 0 && !(d < bd[KEEP - 1] || (d == bd[KEEP - 1] && j < bj[KEEP - 1]))) continue;
          // insert keeping (d, j) ascending lexicographically
          float cd = d; int cj = j;
          bool ins = false;
#pragma unroll
          for (int k = 0; k < KEEP; ++k) {
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
      for (int k = 0; k < MAXNB; ++k) nbr[(size_t)i * MAXNB + k] = (short)((k < KN && k < KEEP) ? bj[k < KEEP ? k : 0] : -1);
    }
  };
  if (KN == 5) knn(std::integral_constant<int, 5>{}); else knn(std::integral_constant<int, MAXNB>{});
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
// cpref != NULL (Progressive-X): *val_out = pixels - (shared support)^2, shared = sum_i min(cpref_i, max(0, 1 - r_i^2/T))
// over the inliers (scoring_function_with_compound_model.h:216-263).  Summation order as defined by the oracle: point i
// goes to the partial sum of lane i mod 32 in ascending order, then a butterfly over the lanes.
// floor_inl > 0: the caller will discard any model with inliers + 1 < floor_inl (the early-out of getScore,
// scoring_function.h:257-259, as the replay applies it); scoring stops as soon as the count cannot reach that bound any
// more and reports zero inliers -- which the replay turns into the same decision.  Returns the points NOT visited.
__device__ inline int score_warp(const SmemPoints& sp, int N, const double* model, double sq_trunc, unsigned int* bits,
                                 int lane, int* inl_out, int* pix_out, const double* __restrict__ cpref = nullptr,
                                 double* val_out = nullptr, int floor_inl = 0) {
  double m[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) m[k] = model[k];
  for (int k = lane; k < NMAX / 32; k += 32) bits[k] = 0u;
  __syncwarp();
  int inl = 0;
  double shared = 0.0;
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
        const int i = base + u * 32 + lane;
        const unsigned int pid = sp.pix[i];
        atomicOr(&bits[pid >> 5], 1u << (pid & 31));
        if (cpref) {
          double pref = 1.0 - sq_residual(sp.un[i], sp.vn[i], sp.x[i], sp.y[i], sp.z[i], m) / sq_trunc;
          if (!(pref > 0.0)) pref = 0.0;
          const double c = cpref[i];
          shared += c < pref ? c : pref;
        }
      }
      inl += __popc(__ballot_sync(0xffffffffu, in[u]));
    }
    const int left = N - base - 128;                 // points not visited yet (warp-uniform, like inl)
    if (left > 0 && inl + left + 1 < floor_inl) {
      *inl_out = 0; *pix_out = 0;
      if (val_out) *val_out = 0.0;
      return left;
    }
  }
  __syncwarp();
  int px = 0;
  for (int k = lane; k < NMAX / 32; k += 32) px += __popc(bits[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) px += __shfl_xor_sync(0xffffffffu, px, o);
  *inl_out = inl;
  *pix_out = px;
  if (val_out) {
    double v = (double)px;
    if (cpref) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) shared += __shfl_xor_sync(0xffffffffu, shared, o);
      v -= shared * shared;
    }
    *val_out = v;
  }
  return 0;
}

// Two models in one sweep over the points (the solutions of a P3P sample come in pairs more often than not): a point
// is read from shared memory once and tested against both -- the single-model sweep is bound as much by its 40 B per
// point of shared-memory reads as by the FP64 pipe.  Same per-point arithmetic, same counting and summation order per
// model as score_warp; bitsA / bitsB: two per-warp bitsets.  Returns the points not visited (both models, when neither can
// reach floor_inl any more).
template <bool CP>                 // CP: compound (Progressive-X) score; cpref is ignored otherwise
__device__ inline int score_warp2(const SmemPoints& sp, int N, const double* modelA, const double* modelB, double sq_trunc,
                                  unsigned int* bitsA, unsigned int* bitsB, int lane, int* inl_out, int* pix_out,
                                  const double* __restrict__ cpref, double* val_out, int floor_inl) {
  double ma[12], mb[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) { ma[k] = modelA[k]; mb[k] = modelB[k]; }
  for (int k = lane; k < NMAX / 32; k += 32) { bitsA[k] = 0u; bitsB[k] = 0u; }
  __syncwarp();
  int cntA = 0, cntB = 0;            // per-lane inlier counts; summed over the warp every fourth block and at the end
  double sharedA = 0.0, sharedB = 0.0;
  for (int base = 0; base < N; base += 96) {
    bool ia[3], ib[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int i = base + u * 32 + lane;
      ia[u] = ib[u] = false;
      if (i < N) {
        const double un = sp.un[i], vn = sp.vn[i], x = sp.x[i], y = sp.y[i], z = sp.z[i];
        ia[u] = is_inlier(un, vn, x, y, z, ma, sq_trunc);
        ib[u] = is_inlier(un, vn, x, y, z, mb, sq_trunc);
      }
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int i = base + u * 32 + lane;
      if (ia[u] || ib[u]) {
        const unsigned int pid = sp.pix[i];
        if (ia[u]) atomicOr(&bitsA[pid >> 5], 1u << (pid & 31));
        if (ib[u]) atomicOr(&bitsB[pid >> 5], 1u << (pid & 31));
        if (CP) {
          const double c = cpref[i];
          if (ia[u]) {
            double pref = 1.0 - sq_residual(sp.un[i], sp.vn[i], sp.x[i], sp.y[i], sp.z[i], ma) / sq_trunc;
            if (!(pref > 0.0)) pref = 0.0;
            sharedA += c < pref ? c : pref;
          }
          if (ib[u]) {
            double pref = 1.0 - sq_residual(sp.un[i], sp.vn[i], sp.x[i], sp.y[i], sp.z[i], mb) / sq_trunc;
            if (!(pref > 0.0)) pref = 0.0;
            sharedB += c < pref ? c : pref;
          }
        }
      }
      cntA += ia[u] ? 1 : 0;
      cntB += ib[u] ? 1 : 0;
    }
    const int left = N - base - 96;
    if (floor_inl > 0 && left > 0 && (base % 288) == 192) {          // every 288 points: can either model still matter?
      const int inlA = __reduce_add_sync(0xffffffffu, cntA), inlB = __reduce_add_sync(0xffffffffu, cntB);
      if (inlA + left + 1 < floor_inl && inlB + left + 1 < floor_inl) {
        inl_out[0] = inl_out[1] = 0; pix_out[0] = pix_out[1] = 0; val_out[0] = val_out[1] = 0.0;
        return 2 * left;
      }
    }
  }
  __syncwarp();
  const int inlA = __reduce_add_sync(0xffffffffu, cntA), inlB = __reduce_add_sync(0xffffffffu, cntB);
  int pxA = 0, pxB = 0;
  for (int k = lane; k < NMAX / 32; k += 32) { pxA += __popc(bitsA[k]); pxB += __popc(bitsB[k]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pxA += __shfl_xor_sync(0xffffffffu, pxA, o);
    pxB += __shfl_xor_sync(0xffffffffu, pxB, o);
  }
  inl_out[0] = inlA; inl_out[1] = inlB; pix_out[0] = pxA; pix_out[1] = pxB;
  double vA = (double)pxA, vB = (double)pxB;
  if (CP) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sharedA += __shfl_xor_sync(0xffffffffu, sharedA, o);
      sharedB += __shfl_xor_sync(0xffffffffu, sharedB, o);
    }
    vA -= sharedA * sharedA; vB -= sharedB * sharedB;
  }
  val_out[0] = vA; val_out[1] = vB;
  return 0;
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
  double val[4];
};

constexpr int CHUNK = 160;           // RANSAC passes evaluated per chunk at most (a multiple of the 20 warps; records survive LO rounds in the workspace)

__device__ __noinline__ void phase_main(const Workspace& ws, const epos_fit_params& prm, int p, unsigned char* smem_raw,
                                        bool& pts_loaded) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ProbState* st = ws.st + p;
  const int N = st->N;
  SmemPoints sp;
  unsigned char* rest = load_points(smem_raw, ws, p, N, &sp, pts_loaded);
  unsigned int* bits = reinterpret_cast<unsigned int*>(rest) + warp * (NMAX / 32);
  unsigned int* bits2 = bits + WARPS * (NMAX / 32);                 // second per-warp bitset (two models per sweep)
  double* best_model = reinterpret_cast<double*>(rest + 2 * WARPS * (NMAX / 32) * 4);
  // the replay's view of the chunk: the small per-pass fields staged in shared memory (the replay is a dependent chain of
  // reads; from the L2-resident records it cost ~2000 cycles per pass, i.e. a quarter of the main phase)
  double* s_val = best_model + 12;                                 // [CHUNK][4]
  int* s_inl = reinterpret_cast<int*>(s_val + CHUNK * 4);          // [CHUNK][4]
  int* s_nm = s_inl + CHUNK * 4;                                   // [CHUNK]
  int* s_fails = s_nm + CHUNK;                                     // [CHUNK]
  PassRecord* recs = ws.recs + (size_t)p * CHUNK;                  // global (L2): hypotheses do not depend on the state
  auto stage_chunk = [&]() {
    if (tid < CHUNK) {
      const PassRecord* rc = recs + tid;
      s_nm[tid] = rc->nm; s_fails[tid] = rc->fails;
      for (int m = 0; m < 4; ++m) { s_inl[tid * 4 + m] = rc->inl[m]; s_val[tid * 4 + m] = rc->val[m]; }
    }
  };
  // state (identical in every thread; the replay's results travel through rs)
  __shared__ struct { unsigned long long iter, max_iteration; double best_value, coverage; int pass, best_inl, lo_runs, gc_count, flags; } rs;
  unsigned long long iter = st->iter, max_iteration = st->max_iteration;
  const unsigned long long seed = st->seed;
  int pass = st->pass, best_inl = st->best_inl, lo_runs = st->lo_runs, gc_count = st->gc_count;
  double best_value = st->best_value;
  const double* cpref = st->cpref;
  const unsigned long long stream0 = st->stream_base;
  int chunk_base = st->chunk_base, chunk_n = st->chunk_n;
  double coverage = st->coverage;
  const int used_pixels = st->used_pixels;
  const double sq_trunc = st->sq_trunc;
  const unsigned long long max_iters = (unsigned long long)prm.max_iters, min_iters = (unsigned long long)prm.min_iters;
  if (tid < 12) best_model[tid] = st->best_model[tid];
  if (pass < chunk_base + chunk_n) stage_chunk();                  // records of a chunk interrupted by an LO round
  __syncthreads();
  bool ended = false, to_lo = false;
  long long t_sample = 0, t_score = 0, t_replay = 0, n_scored = 0, n_skipped = 0;   // n_skipped: per warp
  const long long t_begin = clock64();
  while (true) {
    long long t0 = clock64();
    if (pass >= chunk_base + chunk_n) {
      // ---- phase A: one THREAD per pass: sample until a valid sample gives >= 1 admissible P3P pose ----
      // a pass consumes at least one iteration, so no more than max_iters - iter passes can still run: the chunk is
      // clipped to that (400 iterations = 160 + 160 + 80 passes), rounded up to whole rounds of the 20 warps
      chunk_base = pass;
      {
        const unsigned long long room = max_iters > iter ? max_iters - iter : 1ull;
        int c = room < (unsigned long long)CHUNK ? (int)room : CHUNK;
        c = ((c + WARPS - 1) / WARPS) * WARPS;
        chunk_n = c < CHUNK ? c : CHUNK;
      }
      if (tid < chunk_n) {
        PassRecord* rc = recs + tid;
        const int my_pass = chunk_base + tid;
        int fails = -1, nm = 0;
        const double* PU = ws.pts + (size_t)p * 7 * NMAX + 5 * NMAX;   // original pixel coordinates (u row, v row)
        double models[48];
        while (++fails < prm.max_unsuccessful) {
          int s[3];
          if (!unique_set(seed, stream0, (u64)my_pass, (u64)fails, N, 3, s)) continue;
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
      for (int k = warp; k < chunk_n; k += WARPS) {
        PassRecord* rc = recs + k;
        const int nm = rc->nm;
        int m = 0;
        for (; m + 1 < nm; m += 2) {                                 // two solutions per sweep over the points
          int inl[2], px[2];
          double val[2];
          n_skipped += cpref ? score_warp2<true>(sp, N, rc->models + 12 * m, rc->models + 12 * (m + 1), sq_trunc, bits, bits2,
                                                 lane, inl, px, cpref, val, best_inl)
                             : score_warp2<false>(sp, N, rc->models + 12 * m, rc->models + 12 * (m + 1), sq_trunc, bits, bits2,
                                                  lane, inl, px, cpref, val, best_inl);
          if (lane == 0) {
            rc->inl[m] = inl[0]; rc->pix[m] = px[0]; rc->val[m] = val[0];
            rc->inl[m + 1] = inl[1]; rc->pix[m + 1] = px[1]; rc->val[m + 1] = val[1];
          }
        }
        if (m < nm) {
          int inl, px;
          double val;
          n_skipped += score_warp(sp, N, rc->models + 12 * m, sq_trunc, bits, lane, &inl, &px, cpref, &val, best_inl);
          if (lane == 0) { rc->inl[m] = inl; rc->pix[m] = px; rc->val[m] = val; }
        }
      }
      __syncthreads();
      stage_chunk();
      __syncthreads();
      if (tid == 0) { int sc = 0; for (int k = 0; k < chunk_n; ++k) sc += s_nm[k]; n_scored += sc; }
      t_score += clock64() - t0; t0 = clock64();
    }
    // ---- in-order replay: warp 0 walks the chunk's records (scalar code, every lane the same), the other 19 warps wait;
    // run by all warps it cost 20x the issue slots for the same dependent chain (0.59 of the main phase's 2.9 Mcycles) ----
    if (warp == 0) {
    // Most passes change nothing but the iteration count.  Lane l looks at pass k0 + l: it tests, against the state at the
    // start of the batch, whether the loop would stop before the pass or a model of the pass would be accepted; the
    // passes before the first such lane only advance the counters, that pass then runs the scalar code below.
    int k = pass - chunk_base;
    while (k < chunk_n && !ended && !to_lo) {
      const int kl = k + lane;
      const bool have = kl < chunk_n;
      const unsigned long long step = have ? 1ull + (unsigned long long)s_fails[kl] : 0ull;
      unsigned long long incl = step;                              // inclusive prefix sum of the iteration steps
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const unsigned long long it0 = iter + incl - step;           // iteration count when pass kl starts
      bool event = false;
      if (have) {
        const unsigned long long lim = max_iteration < max_iters ? max_iteration : max_iters;
        if (!(min_iters > it0 || it0 < lim)) event = true;
        if (min_iters < it0 && (it0 > max_iteration || it0 > max_iters || prm.min_coverage < coverage)) event = true;
        const int nm = s_nm[kl];
        for (int m = 0; m < nm; ++m) {
          const double cv = s_inl[kl * 4 + m] + 1 < best_inl ? 0.0 : s_val[kl * 4 + m];
          if (best_value < cv) event = true;
        }
      }
      const unsigned int evm = __ballot_sync(0xffffffffu, event);
      const int nb = chunk_n - k < 32 ? chunk_n - k : 32;          // passes in this batch
      const int f = evm ? __ffs(evm) - 1 : nb;                     // uneventful passes at the head of the batch
      if (f > 0) {
        iter += __shfl_sync(0xffffffffu, incl, f - 1);
        pass += f; k += f;
      }
      if (!evm) continue;
      // ---- pass k: the sequential code ----
      {
      const unsigned long long lim = max_iteration < max_iters ? max_iteration : max_iters;
      if (!(min_iters > iter || iter < lim)) { ended = true; break; }
      if (min_iters < iter) {
        if (iter > max_iteration || iter > max_iters || prm.min_coverage < coverage) { ended = true; break; }
      }
      bool do_lo = false;
      ++iter;
      const PassRecord* rc = recs + k;
      iter += (unsigned long long)s_fails[k];
      const int nm = s_nm[k];
      for (int m = 0; m < nm; ++m) {
        int c_inl = s_inl[k * 4 + m];
        double c_val = s_val[k * 4 + m];
        if (c_inl + 1 < best_inl) { c_inl = 0; c_val = 0.0; }     // early-out of getScore, scoring_function.h:257-259
        if (best_value < c_val) {
          best_value = c_val; best_inl = c_inl;
          if (lane < 12) best_model[lane] = rc->models[12 * m + lane];
          do_lo = iter > (unsigned long long)prm.min_iters_before_lo && best_inl > 3;
          max_iteration = iteration_bound(1.0, best_inl, N);
          coverage = (double)best_value / (double)used_pixels;
        }
      }
      ++pass; ++k;
      if (do_lo) {
        lo_runs += 2;                                               // GCRANSAC.h:409 and :702
        if (gc_count + 1 < prm.max_graph_cuts) { to_lo = true; break; }
        ++gc_count;                                                 // budget exhausted: the while at :710 exits at once
      }
      }
    }
    if (lane == 0) {
      rs.iter = iter; rs.max_iteration = max_iteration; rs.best_value = best_value; rs.coverage = coverage;
      rs.pass = pass; rs.best_inl = best_inl; rs.lo_runs = lo_runs; rs.gc_count = gc_count;
      rs.flags = (ended ? 1 : 0) | (to_lo ? 2 : 0);
    }
    }
    __syncthreads();
    iter = rs.iter; max_iteration = rs.max_iteration; best_value = rs.best_value; coverage = rs.coverage;
    pass = rs.pass; best_inl = rs.best_inl; lo_runs = rs.lo_runs; gc_count = rs.gc_count;
    ended = rs.flags & 1; to_lo = rs.flags & 2;
    __syncthreads();
    t_replay += clock64() - t0;
    if (ended || to_lo) break;
  }
  if (lane == 0 && n_skipped) atomicAdd(reinterpret_cast<unsigned long long*>(&st->pts_skipped), (unsigned long long)n_skipped);
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
struct TrialRecord { double model[12]; double val; int ok, inl, pix, pad; };

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
      unique_set(st->seed, st->stream_base + 1ULL, (u64)gc, (u64)t, ni, sample_size, sel);
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
      double val;
      score_warp(sp, N, rc->model, sq_trunc, bits, lane, &in_, &px, st->cpref, &val);
      if (lane == 0) { rc->ok = 1; rc->inl = in_; rc->pix = px; rc->val = val; }
    }
    __syncwarp();
  }
  __syncthreads();
  if (tid == 0) {
    st->t_fit += t_fit;
    for (int t = 0; t < n_eval; ++t) st->n_scored_lo += recs[t].ok ? 1 : 0;
    bool updated = false;
    double mv = st->lo_value;
    int mi = st->lo_inl;
    if (sample_size < ni) {
      for (int t = 0; t < trials; ++t) {
        const TrialRecord* rc = recs + t;
        if (!rc->ok) continue;                                      // failed fit: `continue` (GCRANSAC.h:748-752)
        int s_inl = rc->inl;
        double s_val = rc->val;
        if (s_inl + 1 < mi) { s_inl = 0; s_val = 0.0; }
        if (mv < s_val) { updated = true; mv = s_val; mi = s_inl; for (int i = 0; i < 12; ++i) st->lo_model[i] = rc->model[i]; }
      }
    } else if (3 < ni) {
      const TrialRecord* rc = recs;                                 // identical model in every trial: first one decides
      if (rc->ok) {
        int s_inl = rc->inl;
        double s_val = rc->val;
        if (s_inl + 1 < mi) { s_inl = 0; s_val = 0.0; }
        if (mv < s_val) { updated = true; mv = s_val; mi = s_inl; for (int i = 0; i < 12; ++i) st->lo_model[i] = rc->model[i]; }
      }
    }
    st->lo_value = mv; st->lo_inl = mi;
    if (gc >= 1 && gc <= TRACE_ROUNDS) {
      int* tr = ws.trace + ((size_t)p * TRACE_ROUNDS + (gc - 1)) * TRACE_COLS;
      tr[0] = gc; tr[1] = ni; tr[2] = updated ? 1 : 0; tr[3] = (int)mv; tr[4] = mi;
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
    if (st->multi) { if (tid == 0) { st->n_final = 0; st->found = 0; } __syncthreads(); return; }
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
  WarpGroupN wgn;
  wgn.rank = tid & 31; wgn.size = 32; wgn.sh = fit_sh;
  constexpr int WARP_N_MAX = 512;      // above this the whole CTA fits (block reductions); below, warp 0 alone
  // Non-minimal fit on an index list: short lists (the no-consensus case: a few dozen inliers) are fitted by warp 0 alone,
  // which avoids two block barriers per reduction; long lists use the whole CTA.
  auto fit_list = [&](const unsigned short* list, int n, double* m2) -> bool {
    PointView v;
    v.un = sp.un; v.vn = sp.vn; v.x = sp.x; v.y = sp.y; v.z = sp.z; v.idx = list; v.n = n;
    if (n > WARP_N_MAX) return fit_nonminimal_group(g, v, m2);
    __syncthreads();
    if (tid < 32) {
      double mm[12];
      const bool ok = n > WARP_FIT_MAX ? fit_nonminimal_group(wgn, v, mm) : fit_nonminimal_group(wg, v, mm);
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
  const double best_value = st->best_value;
  const double* cpref = st->cpref;
  // Progressive-X: score value = pixels - (shared support)^2; the shared support is summed by warp 0 in score_warp's order
  auto compound_value = [&](const double* model, int pixels) -> double {
    if (!cpref) return (double)pixels;
    __syncthreads();
    if (tid < 32) {
      double shared = 0.0;
      for (int i = tid; i < N; i += 32)
        if (is_inlier(sp.un[i], sp.vn[i], sp.x[i], sp.y[i], sp.z[i], model, sq_trunc)) {
          double pref = 1.0 - sq_residual(sp.un[i], sp.vn[i], sp.x[i], sp.y[i], sp.z[i], model) / sq_trunc;
          if (!(pref > 0.0)) pref = 0.0;
          const double c = cpref[i];
          shared += c < pref ? c : pref;
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) shared += __shfl_xor_sync(0xffffffffu, shared, o);
      if (tid == 0) bcast[15] = shared;
    }
    __syncthreads();
    const double sh_ = bcast[15];
    __syncthreads();
    return (double)pixels - sh_ * sh_;
  };
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
      if (best_value < compound_value(cur, pxC)) {
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
  if (st->multi) {
    // Progressive-X proposal: the model and its inliers stay in the workspace (no final LM in this branch of
    // progressivex_python.cpp:136-221); the caller continues with validation / PEARL
    __syncthreads();
    if (tid < 12) ws.fin_model[(size_t)p * 12 + tid] = best_model[tid];
    unsigned short* fo = ws.fin_inl + (size_t)p * NMAX;
    for (int i = tid; i < nA; i += THREADS) fo[i] = listA[i];
    if (tid == 0) { st->n_final = nA; st->found = 1; st->phase = PH_DONE; st->n_scored_final += n_scored; }
    __syncthreads();
    return;
  }
  if (prm.apply_numerical_optimization && nA >= 6) {                                     // progressivex_python.cpp:257-312
    const double R[9] = {best_model[0], best_model[1], best_model[2], best_model[4], best_model[5], best_model[6],
                         best_model[8], best_model[9], best_model[10]};
    double param[6];
    matrix_to_rodrigues(R, param);
    param[3] = best_model[3]; param[4] = best_model[7]; param[5] = best_model[11];
    pv.idx = listA; pv.n = nA;
    if (nA > WARP_N_MAX) {
      lm_refine_group(g, pv, param);
    } else {
      __syncthreads();
      if (tid < 32) {
        if (nA > WARP_FIT_MAX) lm_refine_group(wgn, pv, param); else lm_refine_group(wg, pv, param);
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

// =====================================================================================================
// Progressive-X (multi-instance fitting): ProgressiveX::run with PEARL
//   /root/reference/external/progressive-x/src/pyprogressivex/include/progressive_x.h:397-649,651-794, PEARL.h:271-536,
//   scoring_function_with_compound_model.h:127-266; alpha-expansion with label costs:
//   .../graph-cut-ransac/src/pygcransac/include/GCoptimization.cpp:316-404,452-470,1003-1088,1131-1303.
// One persistent CTA per problem runs the whole outer loop: proposal (the GC-RANSAC state machine above with the
// compound score) -> validation (Tanimoto) -> PEARL (alpha-expansion by preflow-push, refits, rejections) -> compound
// model update -> termination test.  Restated against oracle/posefit.cpp (progx_run), which is pinned on the reference's
// own GCoptimization sources.
// =====================================================================================================
struct MultiParams {
  int max_model_number_for_pearl;   // maximum_model_number_to_optimize
  int min_point_number;             // minimum inliers of an instance and PEARL's label cost
  double confidence;                // conf (0.5)
  double max_tanimoto;              // 0.9
};

// runs the GC-RANSAC state machine of problem p (phases main / cut / trials) until it is ready for phase_final
__device__ inline void run_state_machine(const Workspace& ws, const epos_fit_params& prm, int p, unsigned char* smem_raw,
                                         bool& pts_loaded) {
  ProbState* st = ws.st + p;
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
}

// block-wide sum of one double per thread (fixed order: lanes by butterfly, warps sequentially) -> every thread
__device__ inline double block_sum(double v, double* sh /* WARPS + 1 doubles */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0.0; for (int w = 0; w < WARPS; ++w) s += sh[w]; sh[WARPS] = s; }
  __syncthreads();
  const double r = sh[WARPS];
  __syncthreads();
  return r;
}

struct PearlCtx {
  int N, L, KN;                 // sites, labels (instances + outlier), neighbour slots per site
  double T2, oml, lambda, cost;
  int* label;                   // [N]   (shared memory)
  const double* r2tab;          // [L-1][NMAX]
  const short* nbr; const unsigned char* owned; const unsigned char* mult; const int* rev_off; const int* rev_idx;
  double* cf; double* cr; double* fl;   // per owned slot (stride KN): capacity owner->nbr, nbr->owner, flow owner->nbr
  double* exc; double* af;      // [N]   (shared memory)
  int* dist; unsigned short* order; int* lstart;
  int* newlab;                  // [N] scratch (shared memory)
};

__device__ __forceinline__ double pearl_data_cost(const PearlCtx& c, int i, int l) {      // PEARL.h:81-134
  if (l == c.L - 1) return c.oml;
  const double r2 = c.r2tab[(size_t)l * NMAX + i];
  return r2 > c.T2 ? 2.0 * c.oml : c.oml * r2 / c.T2;
}

// One expansion move on label alpha (GCoptimization.cpp:1239-1303).  The binary energy of the move is minimised by a
// preflow-push max-flow over: the sites whose label differs from alpha, the neighbour arcs between them, and one
// auxiliary node per other label in use (label cost; GCoptimization.cpp:1131-1196: source -> aux capacity `cost`,
// aux -> every site of that label capacity `cost`).  A site keeps its label iff it can still reach the sink in the
// residual graph (what_segment == SINK, residuals below CUT_EPS count as saturated), otherwise it takes alpha.  The move
// is applied iff it lowers the energy by more than 1e-9 (the reference compares the flow value with the energy before
// the move; the oracle uses the same tolerance).  Returns true when the labeling changed.
__device__ __noinline__ bool pearl_expansion_move(const PearlCtx& c, int alpha, double* red_sh) {
  __shared__ int s_cnt[PEARL_MAX], s_qn, s_any, s_adist[PEARL_MAX], s_anext[PEARL_MAX], s_changed;
  __shared__ double s_aexc[PEARL_MAX];
  const int tid = threadIdx.x, N = c.N, L = c.L, KN = c.KN;
  if (tid < PEARL_MAX) s_cnt[tid] = 0;
  if (tid == 0) s_changed = 0;
  __syncthreads();
  for (int i = tid; i < N; i += THREADS) atomicAdd(&s_cnt[c.label[i]], 1);
  __syncthreads();
  if (s_cnt[alpha] == N) return false;                                // no active site
  const double lam = c.lambda;
  // ---- terminal capacities and arc capacities ----
  for (int i = tid; i < N; i += THREADS) {
    const int li = c.label[i];
    const bool act = li != alpha;
    double tr = act ? pearl_data_cost(c, i, li) - pearl_data_cost(c, i, alpha) : 0.0;   // source minus sink capacity
    for (int k = 0; k < KN; ++k) {
      const size_t e = (size_t)i * KN + k;
      double f = 0.0, r = 0.0;
      if (c.owned[(size_t)i * MAXNB + k]) {
        const int j = c.nbr[(size_t)i * MAXNB + k];
        const int lj = c.label[j];
        const double w = lam * (double)c.mult[(size_t)i * MAXNB + k];
        if (act && lj != alpha) {
          // add_term2(x = larger index, y = smaller index, 0, w, w, D) with D = w [l_x != l_y] (GCoptimization.cpp:391-398):
          // terminal(x) += D, arc x -> y capacity w, arc y -> x capacity w - D
          const double D = li != lj ? w : 0.0;
          if (i > j) { tr += D; f = w; r = w - D; } else { f = w - D; r = w; }
        } else if (act) {
          tr += w;                                                    // add_term1(i, 0, w): the neighbour already has alpha
        }
      }
      c.cf[e] = f; c.cr[e] = r; c.fl[e] = 0.0;
    }
    for (int a = c.rev_off[i]; a < c.rev_off[i + 1]; ++a) {           // pairs owned by the other endpoint
      const int e0 = c.rev_idx[a];
      const int x = e0 / MAXNB, k = e0 % MAXNB;
      const int lx = c.label[x];
      const double w = lam * (double)c.mult[(size_t)x * MAXNB + k];
      if (act && lx != alpha) { if (i > x && li != lx) tr += w; }
      else if (act) tr += w;
    }
    c.exc[i] = tr;
    c.af[i] = 0.0;
  }
  if (tid < PEARL_MAX) s_aexc[tid] = (tid < L && tid != alpha && s_cnt[tid] > 0 && c.cost > 0.0) ? c.cost : 0.0;
  __syncthreads();
  // ---- preflow-push in waves (as phase_cut), with the auxiliary label-cost nodes ----
  int depth_cap = 2;
  for (int round = 0; round < 8192; ++round) {
    if (tid == 0) { s_qn = 0; s_any = 0; }
    if (tid < PEARL_MAX) { s_adist[tid] = DIST_INF; s_anext[tid] = DIST_INF; }
    __syncthreads();
    for (int i = tid; i < N; i += THREADS) {
      const bool root = c.label[i] != alpha && c.exc[i] < -CUT_EPS;
      c.dist[i] = root ? 0 : DIST_INF;
      if (root) c.order[atomicAdd(&s_qn, 1)] = (unsigned short)i;
    }
    __syncthreads();
    int level = 0, begin = 0, end = s_qn;
    if (tid == 0) c.lstart[0] = 0;
    bool pending = false;   // an auxiliary node carries the current level and its predecessors are not expanded yet
    while ((begin < end || pending) && level < depth_cap) {
      // sites of this level: predecessors over neighbour arcs, and the auxiliary node of the site's label
      for (int q = begin + tid; q < end; q += THREADS) {
        const int v = c.order[q];
        for (int k = 0; k < KN; ++k)
          if (c.owned[(size_t)v * MAXNB + k]) {
            const int u = c.nbr[(size_t)v * MAXNB + k];
            const size_t e = (size_t)v * KN + k;
            if (c.dist[u] == DIST_INF && c.cr[e] + c.fl[e] > CUT_EPS && atomicCAS(&c.dist[u], DIST_INF, level + 1) == DIST_INF) {
              c.order[atomicAdd(&s_qn, 1)] = (unsigned short)u;
              if (c.exc[u] > 0.0) s_any = 1;
            }
          }
        for (int a = c.rev_off[v]; a < c.rev_off[v + 1]; ++a) {
          const int e0 = c.rev_idx[a];
          const int u = e0 / MAXNB;
          const size_t e = (size_t)u * KN + (e0 % MAXNB);
          if (c.dist[u] == DIST_INF && c.cf[e] - c.fl[e] > CUT_EPS && atomicCAS(&c.dist[u], DIST_INF, level + 1) == DIST_INF) {
            c.order[atomicAdd(&s_qn, 1)] = (unsigned short)u;
            if (c.exc[u] > 0.0) s_any = 1;
          }
        }
        const int lv = c.label[v];
        if (c.cost - c.af[v] > CUT_EPS && s_aexc[lv] >= 0.0) {          // arc aux(lv) -> v has residual capacity
          const int old = atomicCAS(&s_adist[lv], DIST_INF, level + 1);
          if (old == DIST_INF || old == level + 1) atomicMin(&s_anext[lv], v);
        }
      }
      __syncthreads();
      // auxiliary nodes labelled at this level: their predecessors are the sites of that label that hold aux flow
      for (int l = 0; l < L; ++l)
        if (s_adist[l] == level + 1) {
          if (tid == 0 && s_aexc[l] > 0.0) s_any = 1;
        } else if (s_adist[l] == level) {
          for (int i = tid; i < N; i += THREADS)
            if (c.label[i] == l && c.af[i] > CUT_EPS && atomicCAS(&c.dist[i], DIST_INF, level + 1) == DIST_INF) {
              c.order[atomicAdd(&s_qn, 1)] = (unsigned short)i;
              if (c.exc[i] > 0.0) s_any = 1;
            }
        }
      __syncthreads();
      begin = end; end = s_qn; ++level;
      if (tid == 0) c.lstart[level] = begin;
      pending = false;
      for (int l = 0; l < L; ++l) pending |= (s_adist[l] == level);
      __syncthreads();
    }
    // an auxiliary node labelled at the last level still has predecessors to visit: the BFS is exhausted only if none is
    const bool aux_pending = pending;
    const bool exhausted = begin >= end && !aux_pending;
    const int maxlevel = (begin >= end) ? level - 1 : level;
    if (tid == 0) c.lstart[level + 1] = end;
    __syncthreads();
    if (!s_any) {
      if (exhausted) break;
      depth_cap *= 4;
      continue;
    }
    const int top = aux_pending ? level : maxlevel;
    for (int d = top; d >= 1; --d) {
      const int qb = d <= maxlevel ? c.lstart[d] : 0, qe = d <= maxlevel ? c.lstart[d + 1] : 0;
      for (int q = qb + tid; q < qe; q += THREADS) {
        const int i = c.order[q];
        double ex = c.exc[i];
        if (!(ex > 0.0)) continue;
        for (int k = 0; k < KN && ex > 0.0; ++k)
          if (c.owned[(size_t)i * MAXNB + k]) {
            const int j = c.nbr[(size_t)i * MAXNB + k];
            if (c.dist[j] != d - 1) continue;
            const size_t e = (size_t)i * KN + k;
            const double res = c.cf[e] - c.fl[e];
            if (!(res > 0.0)) continue;
            if (ex >= res) { c.fl[e] = c.cf[e]; atomicAdd(&c.exc[j], res); ex -= res; }
            else { c.fl[e] += ex; atomicAdd(&c.exc[j], ex); ex = 0.0; }
          }
        for (int a = c.rev_off[i]; a < c.rev_off[i + 1] && ex > 0.0; ++a) {
          const int e0 = c.rev_idx[a];
          const int x = e0 / MAXNB;
          if (c.dist[x] != d - 1) continue;
          const size_t e = (size_t)x * KN + (e0 % MAXNB);
          const double res = c.cr[e] + c.fl[e];
          if (!(res > 0.0)) continue;
          if (ex >= res) { c.fl[e] = -c.cr[e]; atomicAdd(&c.exc[x], res); ex -= res; }
          else { c.fl[e] -= ex; atomicAdd(&c.exc[x], ex); ex = 0.0; }
        }
        const int li = c.label[i];
        if (ex > 0.0 && s_adist[li] == d - 1 && c.af[i] > 0.0) {        // back along the arc aux -> i
          const double res = c.af[i];
          const double amt = ex >= res ? res : ex;
          c.af[i] = ex >= res ? 0.0 : res - ex;
          atomicAdd(&s_aexc[li], amt);
          ex -= amt;
        }
        c.exc[i] = ex;
      }
      if (tid < L && s_adist[tid] == d && s_aexc[tid] > 0.0 && s_anext[tid] != DIST_INF) {
        const int j = s_anext[tid];                                    // a site of level d - 1 with residual capacity
        const double res = c.cost - c.af[j];
        const double amt = s_aexc[tid] >= res ? res : s_aexc[tid];
        c.af[j] = s_aexc[tid] >= res ? c.cost : c.af[j] + amt;
        atomicAdd(&c.exc[j], amt);
        s_aexc[tid] -= amt;
      }
      __syncthreads();
    }
  }
  // ---- new labeling + energy change ----
  double delta = 0.0;
  int nchg = 0;
  for (int i = tid; i < N; i += THREADS) {
    const int li = c.label[i];
    const int nl = (li != alpha && c.dist[i] == DIST_INF) ? alpha : li;
    c.newlab[i] = nl;
    if (nl != li) { delta += pearl_data_cost(c, i, nl) - pearl_data_cost(c, i, li); ++nchg; }
  }
  __syncthreads();
  for (int i = tid; i < N; i += THREADS)
    for (int k = 0; k < KN; ++k)
      if (c.owned[(size_t)i * MAXNB + k]) {
        const int j = c.nbr[(size_t)i * MAXNB + k];
        const double w = lam * (double)c.mult[(size_t)i * MAXNB + k];
        delta += w * ((c.newlab[i] != c.newlab[j] ? 1.0 : 0.0) - (c.label[i] != c.label[j] ? 1.0 : 0.0));
      }
  if (tid < PEARL_MAX) s_adist[tid] = 0;                                // reuse: label counts of the new labeling
  __syncthreads();
  for (int i = tid; i < N; i += THREADS) atomicAdd(&s_adist[c.newlab[i]], 1);
  if (nchg) s_changed = 1;
  __syncthreads();
  double lab_delta = 0.0;
  if (tid == 0)
    for (int l = 0; l < L; ++l) lab_delta += c.cost * ((s_adist[l] > 0 ? 1.0 : 0.0) - (s_cnt[l] > 0 ? 1.0 : 0.0));
  delta = block_sum(delta + lab_delta, red_sh);
  const bool apply = s_changed && delta < -1e-9;
  if (apply)
    for (int i = tid; i < N; i += THREADS) c.label[i] = c.newlab[i];
  __syncthreads();
  return apply;
}

// energy of the current labeling (GCoptimization.cpp:950-986)
__device__ inline double pearl_energy(const PearlCtx& c, double* red_sh) {
  __shared__ int s_used[PEARL_MAX];
  const int tid = threadIdx.x;
  if (tid < PEARL_MAX) s_used[tid] = 0;
  __syncthreads();
  double e = 0.0;
  for (int i = tid; i < c.N; i += THREADS) {
    const int li = c.label[i];
    s_used[li] = 1;
    e += pearl_data_cost(c, i, li);
    for (int k = 0; k < c.KN; ++k)
      if (c.owned[(size_t)i * MAXNB + k] && li != c.label[c.nbr[(size_t)i * MAXNB + k]])
        e += c.lambda * (double)c.mult[(size_t)i * MAXNB + k];
  }
  __syncthreads();
  if (tid == 0)
    for (int l = 0; l < c.L; ++l) if (s_used[l]) e += c.cost;
  return block_sum(e, red_sh);
}

constexpr size_t SMEM_PEARL = (size_t)NMAX * (4 + 4 + 8 + 8 + 4 + 2) + (NMAX + 2) * 4 + 64 + FIT_SCRATCH_DOUBLES * 8 +
                              (size_t)WARPS * FIT_NRED * 8;
static_assert(SMEM_PEARL <= SMEM_CUT_DYN, "PEARL shared memory must fit in the fit kernel's allocation");

// PEARL::run (PEARL.h:391-459) on the instances ws.models[0 .. n_models) of problem p.  Labels end in ws.labels.
__device__ __noinline__ void pearl_run(const Workspace& ws, const epos_fit_params& prm, const MultiParams& mp, int p,
                                       unsigned char* smem_raw, bool& pts_loaded, bool& have_labels) {
  __shared__ double red_sh[WARPS + 2];
  __shared__ int s_count[PEARL_MAX], scan_sh[WARPS + 1], s_flag, s_src[PEARL_MAX];
  const int tid = threadIdx.x;
  ProbState* st = ws.st + p;
  MultiState* ms = ws.ms + p;
  const int N = st->N;
  double* models = ws.models + (size_t)p * MAX_INSTANCES * 12;
  double* prefs = ws.pref + (size_t)p * PEARL_MAX * NMAX;
  int* glabels = ws.labels + (size_t)p * NMAX;
  const double* P5 = ws.pts + (size_t)p * 7 * NMAX;
  PearlCtx c;
  c.N = N; c.KN = prm.max_neighbors < MAXNB ? prm.max_neighbors : MAXNB;
  c.T2 = 9.0 / 4.0 * st->thr_n * st->thr_n;                           // PEARL.h:49 (not (1.5 thr)^2)
  c.lambda = prm.spatial_coherence_weight; c.oml = 1.0 - c.lambda; c.cost = (double)mp.min_point_number;
  c.r2tab = ws.r2tab + (size_t)p * PEARL_MAX * NMAX;
  c.nbr = ws.nbr + (size_t)p * NMAX * MAXNB; c.owned = ws.owned + (size_t)p * NMAX * MAXNB;
  c.mult = ws.mult + (size_t)p * NMAX * MAXNB;
  c.rev_off = ws.rev_off + (size_t)p * (NMAX + 1); c.rev_idx = ws.rev_idx + (size_t)p * NMAX * MAXNB;
  c.cf = ws.capf + (size_t)p * NMAX * MAXNB; c.cr = ws.cr + (size_t)p * NMAX * MAXNB; c.fl = ws.flow + (size_t)p * NMAX * MAXNB;
  // shared-memory carve-up (the point set is evicted; fits below read the points from the L2-resident workspace)
  unsigned char* sm = smem_raw;
  c.label = reinterpret_cast<int*>(sm); sm += NMAX * 4;
  c.newlab = reinterpret_cast<int*>(sm); sm += NMAX * 4;
  c.exc = reinterpret_cast<double*>(sm); sm += NMAX * 8;
  c.af = reinterpret_cast<double*>(sm); sm += NMAX * 8;
  c.dist = reinterpret_cast<int*>(sm); sm += NMAX * 4;
  c.lstart = reinterpret_cast<int*>(sm); sm += (NMAX + 2) * 4;
  c.order = reinterpret_cast<unsigned short*>(sm); sm += NMAX * 2;
  sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sm) + 15) & ~(uintptr_t)15);
  double* fit_sh = reinterpret_cast<double*>(sm); sm += FIT_SCRATCH_DOUBLES * 8;
  double* part = reinterpret_cast<double*>(sm); sm += WARPS * FIT_NRED * 8;
  pts_loaded = false;
  CtaGroup g;
  g.rank = tid; g.size = THREADS; g.sh = fit_sh; g.part = part;
  unsigned short* list = ws.lists + (size_t)p * NMAX;
  int iteration = 0;
  double energy = 0.0, previous_energy = -1.0;
  bool rejected = false, converged = false;
  __syncthreads();
  while (!converged && iteration++ < 50) {
    const bool init_prev = iteration > 1 && !rejected;
    int M = ms->n_models;
    __syncthreads();
    if (M > 0) {                                                      // labeling(), PEARL.h:461-536
      c.L = M + 1;
      double* r2 = const_cast<double*>(c.r2tab);
      for (int i = tid; i < N; i += THREADS)
        for (int l = 0; l < M; ++l)
          r2[(size_t)l * NMAX + i] = sq_residual(P5[i], P5[NMAX + i], P5[2 * NMAX + i], P5[3 * NMAX + i], P5[4 * NMAX + i],
                                                 models + 12 * l);
      for (int i = tid; i < N; i += THREADS) c.label[i] = (init_prev && have_labels) ? glabels[i] : 0;
      __syncthreads();
      for (int cycle = 1; cycle <= 1000; ++cycle) {                   // GCoptimization.cpp:1063-1075, standard cycles
        bool any = false;
        for (int a = 0; a < c.L; ++a) {
          any |= pearl_expansion_move(c, a, red_sh);
          if (tid == 0) ++ms->moves;
        }
        if (!any) break;
      }
      energy = pearl_energy(c, red_sh);
      for (int i = tid; i < N; i += THREADS) glabels[i] = c.label[i];
      have_labels = true;
      __syncthreads();
    }
    bool changed = false;
    rejected = false;
    M = ms->n_models;
    if (have_labels) {                                                // parameterEstimation, PEARL.h:313-389
      if (tid < PEARL_MAX) s_count[tid] = 0;
      __syncthreads();
      for (int k = 0; k < M; ++k) {
        // ordered member list of instance k
        int cnt = 0;
        const int per = PER_THREAD;
        for (int q = 0; q < per; ++q) { const int i = tid * per + q; cnt += (i < N && glabels[i] == k); }
        int total;
        int base = block_excl_scan<WARPS>(cnt, scan_sh, &total);
        for (int q = 0; q < per; ++q) { const int i = tid * per + q; if (i < N && glabels[i] == k) list[base++] = (unsigned short)i; }
        if (tid == 0) s_count[k] = total;
        __syncthreads();
        if (total < 3) continue;
        double before = 0.0;
        for (int q = tid; q < total; q += THREADS) {
          const int i = list[q];
          before += sqrt(sq_residual(P5[i], P5[NMAX + i], P5[2 * NMAX + i], P5[3 * NMAX + i], P5[4 * NMAX + i], models + 12 * k));
        }
        before = block_sum(before, red_sh);
        PointView v;
        v.un = P5; v.vn = P5 + NMAX; v.x = P5 + 2 * NMAX; v.y = P5 + 3 * NMAX; v.z = P5 + 4 * NMAX; v.idx = list; v.n = total;
        double m2[12];
        const bool ok = fit_nonminimal_group(g, v, m2);
        __syncthreads();
        if (!ok) continue;
        double after = 0.0;
        for (int q = tid; q < total; q += THREADS) {
          const int i = list[q];
          after += sqrt(sq_residual(P5[i], P5[NMAX + i], P5[2 * NMAX + i], P5[3 * NMAX + i], P5[4 * NMAX + i], m2));
        }
        after = block_sum(after, red_sh);
        if (after < before) {
          __syncthreads();
          if (tid < 12) models[12 * k + tid] = m2[tid];
          changed = true;
          __syncthreads();
        }
      }
      // outliers of this labeling (before rejections)
      int oc = 0;
      for (int i = tid; i < N; i += THREADS) oc += glabels[i] >= M;
      oc = (int)(block_sum((double)oc, red_sh) + 0.5);
      // rejectInstances, PEARL.h:271-311 (last to first; removed instances hand their points to the outliers)
      if (tid == 0) {
        int nm = 0;
        for (int k = 0; k < M; ++k) {
          if (s_count[k] < mp.min_point_number) oc += s_count[k]; else s_src[nm++] = k;
        }
        s_flag = nm;
        ms->n_models = nm;
        ms->n_outliers = oc;
      }
      __syncthreads();
      const int nm = s_flag;
      rejected = nm != M;
      if (rejected)
        for (int q = 0; q < nm; ++q) {                                // survivors move down in order (src >= q)
          const int src = s_src[q];
          if (src != q) {
            if (tid < 12) models[12 * q + tid] = models[12 * src + tid];
            for (int i = tid; i < N; i += THREADS) prefs[(size_t)q * NMAX + i] = prefs[(size_t)src * NMAX + i];
          }
          __syncthreads();
        }
      __syncthreads();
    }
    if (!rejected && !changed && fabs(energy - previous_energy) < 1e-5 && iteration > 1) converged = true;
    previous_energy = energy;
  }
  if (tid == 0) ms->pearl_iterations += iteration > 50 ? 50 : iteration;
  __syncthreads();
}

// thread 0: state of a fresh GC-RANSAC run (proposal `it`) on problem p
__device__ inline void proposal_reset(ProbState* st, int it, const double* cpref) {
  st->phase = PH_MAIN; st->iter = 0; st->pass = 0; st->best_value = 0.0; st->best_inl = 0; st->coverage = 0.0;
  st->lo_runs = 0; st->gc_count = 0; st->lo_final = 0; st->lo_stage = 0; st->ni = 0; st->found = 0;
  st->lo_value = 0.0; st->lo_inl = 0; st->chunk_base = 0; st->chunk_n = 0; st->n_final = 0;
  st->max_iteration = ~0ULL;
  for (int i = 0; i < 12; ++i) { st->best_model[i] = 0.0; st->lo_model[i] = 0.0; }
  st->cpref = cpref; st->stream_base = 2ULL * (unsigned long long)it; st->multi = 1;
}

// Neighbourhood graph of spedUpFitting (progressive_x.h:288-289: FlannNeighborhoodGraph over ALL 7 columns of the
// remaining rows with a hard-coded radius of 20): deterministic stand-in as in prep_kernel -- the max_neighbors nearest
// rows within the radius, f32 L2 over (u_n, v_n, x, y, z, u, v), ties by index -- plus edge ownership and reverse
// adjacency.  Called by the whole CTA (THREADS threads) on the current (compacted) point set of problem p.
__device__ __noinline__ void build_graph7(const Workspace& ws, const epos_fit_params& prm, int p, int N, unsigned char* smem_raw) {
  constexpr int NT = THREADS, DIMS = 7, UROW = 5, VROW = 6, MAXCELL = 4096;
  __shared__ float s_red[4][NT / 32];
  __shared__ float s_box[4];
  __shared__ int scan_sh[NT / 32 + 1];
  const int tid = threadIdx.x;
  const double* P5 = ws.pts + (size_t)p * 7 * NMAX;
  float* q = reinterpret_cast<float*>(smem_raw);                                   // [7][NMAX]
  int* cell_start = reinterpret_cast<int*>(smem_raw + (size_t)DIMS * NMAX * 4);    // [MAXCELL + 8]
  unsigned short* sorted = reinterpret_cast<unsigned short*>(cell_start + MAXCELL + 8);   // [NMAX]
  int* cursor = reinterpret_cast<int*>(sorted + NMAX);                             // [MAXCELL]
  __syncthreads();
  for (int i = tid; i < N; i += NT)
    for (int d = 0; d < DIMS; ++d) q[d * NMAX + i] = (float)P5[(size_t)d * NMAX + i];
  __syncthreads();
  const int KN = prm.max_neighbors < MAXNB ? prm.max_neighbors : MAXNB;
  const float rad = 20.0f, r2 = rad * rad;
  short* nbr = ws.nbr + (size_t)p * NMAX * MAXNB;
  {
    float mnu = INFINITY, mnv = INFINITY, mxu = -INFINITY, mxv = -INFINITY;
    for (int i = tid; i < N; i += NT) {
      mnu = fminf(mnu, q[UROW * NMAX + i]); mxu = fmaxf(mxu, q[UROW * NMAX + i]);
      mnv = fminf(mnv, q[VROW * NMAX + i]); mxv = fmaxf(mxv, q[VROW * NMAX + i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mnu = fminf(mnu, __shfl_xor_sync(0xffffffffu, mnu, o)); mxu = fmaxf(mxu, __shfl_xor_sync(0xffffffffu, mxu, o));
      mnv = fminf(mnv, __shfl_xor_sync(0xffffffffu, mnv, o)); mxv = fmaxf(mxv, __shfl_xor_sync(0xffffffffu, mxv, o));
    }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = mnu; s_red[1][tid >> 5] = mxu; s_red[2][tid >> 5] = mnv; s_red[3][tid >> 5] = mxv; }
    __syncthreads();
    if (tid == 0) {
      for (int k = 1; k < NT / 32; ++k) {
        s_red[0][0] = fminf(s_red[0][0], s_red[0][k]); s_red[1][0] = fmaxf(s_red[1][0], s_red[1][k]);
        s_red[2][0] = fminf(s_red[2][0], s_red[2][k]); s_red[3][0] = fmaxf(s_red[3][0], s_red[3][k]);
      }
      float cs = rad * 1.0001f + 1e-6f;
      int gw = 1, gh = 1;
      if (N > 0)
        for (;;) {
          gw = (int)floorf((s_red[1][0] - s_red[0][0]) / cs) + 1;
          gh = (int)floorf((s_red[3][0] - s_red[2][0]) / cs) + 1;
          if (gw > 0 && gh > 0 && (long long)gw * gh <= MAXCELL) break;
          cs *= 2.0f;
          if (!(cs < 1e30f)) { gw = gh = 1; break; }
        }
      s_box[0] = s_red[0][0]; s_box[1] = s_red[2][0]; s_box[2] = cs; s_box[3] = __int_as_float(gw | (gh << 16));
    }
    __syncthreads();
  }
  const float bu = s_box[0], bv = s_box[1], ics = 1.0f / s_box[2];
  const int gw = __float_as_int(s_box[3]) & 0xffff, gh = __float_as_int(s_box[3]) >> 16;
  const int ncell = gw * gh;
  auto cell_of = [&](int i, int* cx, int* cy) {
    int x = (int)floorf((q[UROW * NMAX + i] - bu) * ics), y = (int)floorf((q[VROW * NMAX + i] - bv) * ics);
    *cx = x < 0 ? 0 : (x >= gw ? gw - 1 : x);
    *cy = y < 0 ? 0 : (y >= gh ? gh - 1 : y);
  };
  for (int c = tid; c <= ncell; c += NT) cell_start[c] = 0;
  __syncthreads();
  for (int i = tid; i < N; i += NT) { int cx, cy; cell_of(i, &cx, &cy); atomicAdd(&cell_start[cy * gw + cx + 1], 1); }
  __syncthreads();
  {
    constexpr int per = (MAXCELL + NT - 1) / NT;
    int loc = 0;
    for (int k = 0; k < per; ++k) { const int c = tid * per + k + 1; if (c <= ncell) loc += cell_start[c]; }
    int total;
    int base = block_excl_scan<NT / 32>(loc, scan_sh, &total);
    for (int k = 0; k < per; ++k) { const int c = tid * per + k + 1; if (c <= ncell) { base += cell_start[c]; cell_start[c] = base; } }
  }
  __syncthreads();
  for (int c = tid; c < ncell; c += NT) cursor[c] = cell_start[c + 1];
  __syncthreads();
  for (int i = tid; i < N; i += NT) { int cx, cy; cell_of(i, &cx, &cy); sorted[atomicSub(&cursor[cy * gw + cx], 1) - 1] = (unsigned short)i; }
  __syncthreads();
  for (int i = tid; i < N; i += NT) {
    float bd[MAXNB];
    int bj[MAXNB];
#pragma unroll
    for (int k = 0; k < MAXNB; ++k) { bd[k] = INFINITY; bj[k] = -1; }
    float a[DIMS];
#pragma unroll
    for (int d = 0; d < DIMS; ++d) a[d] = q[d * NMAX + i];
    int cx, cy;
    cell_of(i, &cx, &cy);
    for (int yy = max(cy - 1, 0); yy <= min(cy + 1, gh - 1); ++yy) {
      const int c0 = yy * gw + max(cx - 1, 0), c1 = yy * gw + min(cx + 1, gw - 1);
      for (int s_ = cell_start[c0]; s_ < cell_start[c1 + 1]; ++s_) {
        const int j = sorted[s_];
        float dsum = 0.f;
#pragma unroll
        for (int d = 0; d < DIMS; ++d) { const float e = a[d] - q[d * NMAX + j]; dsum = __fmaf_rn(e, e, dsum); }
        if (j == i || !(dsum <= r2)) continue;
        float cd = dsum; int cj = j;
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
  // edge ownership + reverse adjacency (as prep_kernel)
  unsigned char* owned = ws.owned + (size_t)p * NMAX * MAXNB;
  int* rev_off = ws.rev_off + (size_t)p * (NMAX + 1);
  int* rev_idx = ws.rev_idx + (size_t)p * NMAX * MAXNB;
  int* indeg = reinterpret_cast<int*>(smem_raw);                                   // q is dead now
  for (int i = tid; i <= N; i += NT) indeg[i] = 0;
  __syncthreads();
  for (int i = tid; i < N; i += NT)
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
    constexpr int per = (NMAX + NT - 1) / NT;
    int loc[per];
    int s = 0;
    for (int k = 0; k < per; ++k) { const int i = tid * per + k; loc[k] = i < N ? indeg[i] : 0; s += loc[k]; }
    int total;
    int base = block_excl_scan<NT / 32>(s, scan_sh, &total);
    for (int k = 0; k < per; ++k) { const int i = tid * per + k; if (i < N) rev_off[i] = base; base += loc[k]; }
    if (tid == 0) rev_off[N] = total;
    __syncthreads();
    for (int i = tid; i < N; i += NT) indeg[i] = rev_off[i];
    __syncthreads();
    for (int i = tid; i < N; i += NT)
      for (int k = 0; k < MAXNB; ++k)
        if (owned[(size_t)i * MAXNB + k]) {
          const int j = nbr[(size_t)i * MAXNB + k];
          rev_idx[atomicAdd(&indeg[j], 1)] = i * MAXNB + k;
        }
    __syncthreads();
    for (int i = tid; i < N; i += NT) {
      const int b = rev_off[i], e = rev_off[i + 1];
      for (int a2 = b + 1; a2 < e; ++a2) {
        const int v = rev_idx[a2];
        int c2 = a2 - 1;
        while (c2 >= b && rev_idx[c2] > v) { rev_idx[c2 + 1] = rev_idx[c2]; --c2; }
        rev_idx[c2 + 1] = v;
      }
    }
  }
  __syncthreads();
}

// spedUpFitting (progressive_x.h:265-391): sequential GC-RANSAC proposals (plain EPOS score, confidence 1) on the points
// the accepted proposals have not explained yet; no PEARL; the labeling stays zero and the scores stay zero as in the
// reference.  max_model_number = -1 ("all instances"): the reference's loop never terminates (size_t counter against an
// int holding -1); defined here as in the oracle: stop when no model is found, when a proposal has fewer than
// min_point_number inliers, or when the unseen-inlier test of ProgressiveX::run fires; at most MAX_INSTANCES instances.
__device__ __noinline__ void sped_up_fitting(const Workspace& ws, const epos_fit_params& prm, const MultiParams& mp, int p,
                                             int max_model_number, unsigned char* smem_raw, const int* __restrict__ offsets,
                                             double* __restrict__ poses, int* __restrict__ labeling) {
  __shared__ int scan_sh[WARPS + 1], s_stop;
  const int tid = threadIdx.x;
  ProbState* st = ws.st + p;
  MultiState* ms = ws.ms + p;
  double* models = ws.models + (size_t)p * MAX_INSTANCES * 12;
  double* P5 = ws.pts + (size_t)p * 7 * NMAX;
  unsigned short* pixg = ws.pix + (size_t)p * NMAX;
  const bool unbounded = max_model_number < 0;
  const int limit = unbounded ? MAX_INSTANCES : (max_model_number < MAX_INSTANCES ? max_model_number : MAX_INSTANCES);
  bool pts_loaded = false;
  if (tid == 0) ms->sped_up = 1;
  for (int it = 0; it < limit; ++it) {
    __syncthreads();
    const int N = st->N;
    build_graph7(ws, prm, p, N, smem_raw);
    pts_loaded = false;
    if (tid == 0) { proposal_reset(st, it, nullptr); ++ms->proposals; }
    __syncthreads();
    run_state_machine(ws, prm, p, smem_raw, pts_loaded);
    phase_final(ws, prm, p, smem_raw, pts_loaded, offsets, poses, labeling);
    __syncthreads();
    if (!st->found) { if (unbounded) break; continue; }
    const int nin = st->n_final;
    const int slot = ms->n_models;
    __syncthreads();
    if (tid < 12) models[12 * slot + tid] = ws.fin_model[(size_t)p * 12 + tid];
    if (tid == 0) { ms->total_iterations += (long long)st->iter; ms->n_models = slot + 1; ++ms->accepted; ms->scores[slot] = 0.0; }
    __syncthreads();
    if (nin < mp.min_point_number) { if (unbounded) break; continue; }
    // ---- remove the proposal's inliers: ordered compaction of the 7 point rows and the pixel ids ----
    int* newidx = reinterpret_cast<int*>(smem_raw);                 // [NMAX]  -1 = removed
    double* buf = reinterpret_cast<double*>(smem_raw + NMAX * 4);   // [NMAX]
    const unsigned short* fi = ws.fin_inl + (size_t)p * NMAX;
    for (int i = tid; i < N; i += THREADS) newidx[i] = 0;
    __syncthreads();
    for (int i = tid; i < nin; i += THREADS) newidx[fi[i]] = -1;
    __syncthreads();
    {
      const int per = PER_THREAD;
      int cnt = 0;
      for (int q = 0; q < per; ++q) { const int i = tid * per + q; cnt += (i < N && newidx[i] == 0); }
      int total;
      int base = block_excl_scan<WARPS>(cnt, scan_sh, &total);
      for (int q = 0; q < per; ++q) { const int i = tid * per + q; if (i < N && newidx[i] == 0) newidx[i] = base++; }
      __syncthreads();
      for (int r = 0; r < 7; ++r) {
        for (int i = tid; i < N; i += THREADS) buf[i] = P5[(size_t)r * NMAX + i];
        __syncthreads();
        for (int i = tid; i < N; i += THREADS) if (newidx[i] >= 0) P5[(size_t)r * NMAX + newidx[i]] = buf[i];
        __syncthreads();
      }
      unsigned short* sb = reinterpret_cast<unsigned short*>(buf);
      for (int i = tid; i < N; i += THREADS) sb[i] = pixg[i];
      __syncthreads();
      for (int i = tid; i < N; i += THREADS) if (newidx[i] >= 0) pixg[newidx[i]] = sb[i];
      __syncthreads();
      if (tid == 0) {
        st->N = total;
        s_stop = 0;
        if (unbounded) {
          const double ratio = pow(1.0 - pow(1.0 - mp.confidence, 1.0 / (double)ms->total_iterations), 1.0 / 3.0);
          s_stop = llround((double)total * ratio) < (long long)mp.min_point_number;
        }
      }
      __syncthreads();
    }
    if (s_stop) break;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(THREADS, 1)
progx_kernel(Workspace ws, epos_fit_params prm, MultiParams mp, const int* __restrict__ offsets,
             const int* __restrict__ max_models, double* __restrict__ poses, int* __restrict__ labeling,
             double* __restrict__ multi_poses, double* __restrict__ multi_scores, int* __restrict__ multi_counts) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ double red_sh[WARPS + 2];
  __shared__ int s_stop;
  const int p = blockIdx.x, tid = threadIdx.x;
  ProbState* st = ws.st + p;
  MultiState* ms = ws.ms + p;
  const int N = st->N, off = offsets[p];
  const int max_model_number = max_models[p];
  double* models = ws.models + (size_t)p * MAX_INSTANCES * 12;
  double* prefs = ws.pref + (size_t)p * PEARL_MAX * NMAX;
  double* compound = ws.compound + (size_t)p * NMAX;
  int* glabels = ws.labels + (size_t)p * NMAX;
  const double* P5 = ws.pts + (size_t)p * 7 * NMAX;
  double* out_poses = multi_poses + (size_t)p * MAX_INSTANCES * 12;
  double* out_scores = multi_scores + (size_t)p * MAX_INSTANCES;
  if (tid == 0) {
    ms->total_iterations = 0; ms->n_models = 0; ms->unaccepted = 0; ms->first_events = 0; ms->n_outliers = 0;
    ms->proposals = 0; ms->accepted = 0; ms->pearl_iterations = 0; ms->moves = 0; ms->done = 0;
    ms->max_models = max_model_number; ms->sped_up = 0;
    multi_counts[p] = 0;
  }
  for (int i = tid; i < N; i += THREADS) { compound[i] = 0.0; labeling[off + i] = 0; }
  // multiplicity of the owned neighbour pairs: PEARL enters a pair once per LISTING (PEARL.h:517-520)
  {
    const short* nbr = ws.nbr + (size_t)p * NMAX * MAXNB;
    const unsigned char* owned = ws.owned + (size_t)p * NMAX * MAXNB;
    unsigned char* mult = ws.mult + (size_t)p * NMAX * MAXNB;
    for (int i = tid; i < N; i += THREADS)
      for (int k = 0; k < MAXNB; ++k) {
        unsigned char m = 0;
        if (owned[(size_t)i * MAXNB + k]) {
          const int j = nbr[(size_t)i * MAXNB + k];
          m = 1;
          for (int q = 0; q < MAXNB; ++q) m += (nbr[(size_t)j * MAXNB + q] == i) ? 1 : 0;
        }
        mult[(size_t)i * MAXNB + k] = m;
      }
  }
  __syncthreads();
  if (!st->valid || st->phase == PH_DONE) return;                      // fewer than 6 correspondences, singular K, ...
  if (max_model_number == 0 || max_model_number == 1 || max_model_number < -1) {
    if (tid == 0) multi_counts[p] = -1;                                // not a multi-instance bound: rejected
    return;
  }
  if (max_model_number < 0 || max_model_number > mp.max_model_number_for_pearl) {          // progressive_x.h:417-425
    sped_up_fitting(ws, prm, mp, p, max_model_number, smem_raw, offsets, poses, labeling);
    const int M = ms->n_models;
    for (int i = tid; i < 12 * M; i += THREADS) out_poses[i] = models[i];
    for (int i = tid; i < M; i += THREADS) out_scores[i] = 0.0;        // RANSACStatistics::score is never written
    if (tid == 0) {
      multi_counts[p] = M;
      ms->done = 1;
      double* rec = poses + (size_t)p * EPOS_POSE_RECORD_DOUBLES;
      for (int i = 0; i < 12; ++i) rec[i] = M > 0 ? models[i] : 0.0;
      rec[12] = 0.0; rec[13] = (double)ms->total_iterations; rec[14] = (double)M; rec[15] = (double)ms->proposals;
    }
    return;
  }
  const double T2 = 9.0 / 4.0 * st->thr_n * st->thr_n;                 // progressive_x.h:679
  bool pts_loaded = false, have_labels = false;
  for (int it = 0; it <= 100; ++it) {                                  // progressive_x.h:427-437
    __syncthreads();
    if (tid == 0) { proposal_reset(st, it, ms->n_models > 0 ? compound : nullptr); ++ms->proposals; }
    __syncthreads();
    run_state_machine(ws, prm, p, smem_raw, pts_loaded);
    phase_final(ws, prm, p, smem_raw, pts_loaded, offsets, poses, labeling);
    __syncthreads();
    if (!st->found) continue;
    const int nin = st->n_final;
    if (tid == 0) ms->total_iterations += (long long)st->iter;
    // ---- isPutativeModelValid (progressive_x.h:723-749) ----
    bool valid = nin >= (3 > mp.min_point_number ? 3 : mp.min_point_number);
    const int slot = ms->n_models;                                     // where the proposal goes if accepted
    __syncthreads();
    if (valid) {
      const double* fm = ws.fin_model + (size_t)p * 12;
      double dot = 0.0, na = 0.0, nc = 0.0;
      for (int i = tid; i < N; i += THREADS) {
        double v = 1.0 - sq_residual(P5[i], P5[NMAX + i], P5[2 * NMAX + i], P5[3 * NMAX + i], P5[4 * NMAX + i], fm) / T2;
        if (!(v > 0.0)) v = 0.0;
        prefs[(size_t)slot * NMAX + i] = v;
        const double cc = compound[i];
        dot += v * cc; na += v * v; nc += cc * cc;
      }
      dot = block_sum(dot, red_sh); na = block_sum(na, red_sh); nc = block_sum(nc, red_sh);
      const double tanimoto = dot / (na + nc - dot);
      if (mp.max_tanimoto < tanimoto) valid = false;
    }
    if (!valid) {
      if (tid == 0) { ++ms->unaccepted; s_stop = ms->unaccepted == 10; }
      __syncthreads();
      if (s_stop) break;
      continue;
    }
    // ---- accept: compound instance optimisation (progressive_x.h:505-547) ----
    if (tid < 12) models[12 * slot + tid] = ws.fin_model[(size_t)p * 12 + tid];
    if (tid == 0) { ms->n_models = slot + 1; ++ms->accepted; }
    __syncthreads();
    if (slot == 0) {
      const unsigned short* fi = ws.fin_inl + (size_t)p * NMAX;
      for (int i = tid; i < N; i += THREADS) glabels[i] = 1;
      __syncthreads();
      for (int i = tid; i < nin; i += THREADS) glabels[fi[i]] = 0;
      if (tid == 0) ++ms->first_events;
      have_labels = false;                 // PEARL's own engine does not exist yet (labels above are ProgressiveX's)
      __syncthreads();
    } else {
      pearl_run(ws, prm, mp, p, smem_raw, pts_loaded, have_labels);
    }
    // ---- updateCompoundModel (progressive_x.h:755-794) ----
    const int M = ms->n_models;
    __syncthreads();
    if (M > 0) {
      for (int i = tid; i < N; i += THREADS) {
        double mx = 0.0;
        for (int k = 0; k < M; ++k) mx = fmax(mx, prefs[(size_t)k * NMAX + i]);
        compound[i] = mx;
      }
      for (int k = 0; k < M; ++k) {
        double sc = 0.0;
        for (int i = tid; i < N; i += THREADS) sc += prefs[(size_t)k * NMAX + i];
        sc = block_sum(sc, red_sh);
        if (tid == 0) ms->scores[k] = sc;
      }
    }
    // ---- termination (progressive_x.h:589-615, 651-673) ----
    if (tid == 0) {
      const long long covered = M == 1 ? (long long)ms->first_events : (long long)N - (long long)ms->n_outliers;
      const double ratio = pow(1.0 - pow(1.0 - mp.confidence, 1.0 / (double)ms->total_iterations), 1.0 / 3.0);
      const long long unseen = llround((double)((long long)N - covered) * ratio);
      s_stop = (unseen < (long long)mp.min_point_number) || (M >= max_model_number);
    }
    __syncthreads();
    if (s_stop) break;
  }
  __syncthreads();
  // ---- outputs ----
  const int M = ms->n_models;
  for (int i = tid; i < 12 * M; i += THREADS) out_poses[i] = models[i];
  for (int i = tid; i < M; i += THREADS) out_scores[i] = ms->scores[i];
  if (ms->accepted > 0)
    for (int i = tid; i < N; i += THREADS) labeling[off + i] = glabels[i];
  if (tid == 0) {
    multi_counts[p] = M;
    ms->done = 1;
    double* rec = poses + (size_t)p * EPOS_POSE_RECORD_DOUBLES;      // record = first instance (valid = number of instances)
    for (int i = 0; i < 12; ++i) rec[i] = M > 0 ? models[i] : 0.0;
    int n0 = 0;
    if (M > 0) for (int i = 0; i < N; ++i) n0 += glabels[i] == 0;
    rec[12] = (double)n0; rec[13] = (double)ms->total_iterations; rec[14] = (double)M; rec[15] = (double)ms->proposals;
  }
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
static_assert(sizeof(PassRecord) <= 464 && CHUNK == 160, "workspace_layout reserves 160 x 464 bytes of pass records");
constexpr size_t SMEM_MAIN = SMEM_POINTS + 2 * WARPS * (NMAX / 32) * 4 + 12 * 8 + CHUNK * (4 * 8 + 4 * 4 + 4 + 4) + 64;
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
//                  models scored over all N points in main / LO trials / final, points the early-out skipped.
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
    o[16] = h[i].n_scored_main; o[17] = h[i].n_scored_lo; o[18] = h[i].n_scored_final; o[19] = h[i].pts_skipped;
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
  EPOS_CUDA(cudaFuncSetAttribute(progx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FIT));
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

size_t epos_fit_multi_workspace_bytes(int P) {
  if (P <= 0) return 0;
  return workspace_layout(P, nullptr, nullptr, true);
}

int epos_fit_max_instances(void) { return MAX_INSTANCES; }

int epos_fit_poses_multi(const double* coord_2d, const double* coord_3d, const int32_t* offsets, const int32_t* counts, int P,
                         const double* K, const uint64_t* seeds, const epos_fit_params* params,
                         const epos_multi_params* mparams, const int32_t* max_models, double* poses, int32_t* labeling,
                         double* multi_poses, double* multi_scores, int32_t* multi_counts, void* workspace,
                         size_t workspace_bytes, void* stream) {
  EPOS_CHECK_ARG(coord_2d && coord_3d && offsets && counts && K && seeds && params && mparams && max_models && poses &&
                 labeling && multi_poses && multi_scores && multi_counts && workspace);
  EPOS_CHECK_ARG(P > 0);
  EPOS_CHECK_ARG(params->max_neighbors >= 0 && params->max_neighbors <= MAXNB);
  EPOS_CHECK_ARG(params->max_lo_trials >= 0 && params->max_lo_trials <= MAX_TRIALS);
  EPOS_CHECK_ARG(params->max_graph_cuts >= 1 && params->max_graph_cuts <= 64);
  EPOS_CHECK_ARG(params->spatial_coherence_weight > 0.0);
  EPOS_CHECK_ARG(mparams->max_model_number_for_pearl >= 2 && mparams->max_model_number_for_pearl < PEARL_MAX);
  EPOS_CHECK_ARG(mparams->min_point_number >= 1 && mparams->confidence > 0.0 && mparams->confidence < 1.0);
  EPOS_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  if (workspace_bytes < workspace_layout(P, nullptr, nullptr, true)) {
    set_error("epos_fit_poses_multi: workspace too small (%zu < %zu)", workspace_bytes, workspace_layout(P, nullptr, nullptr, true));
    return EPOS_ERR_INVALID_ARG;
  }
  int rc = set_smem_attrs();
  if (rc) return rc;
  Workspace ws;
  workspace_layout(P, workspace, &ws, true);
  cudaStream_t s = (cudaStream_t)stream;
  MultiParams mp;
  mp.max_model_number_for_pearl = mparams->max_model_number_for_pearl; mp.min_point_number = mparams->min_point_number;
  mp.confidence = mparams->confidence; mp.max_tanimoto = mparams->max_tanimoto_similarity;
  prep_kernel<<<P, PT, SMEM_PREP, s>>>(ws, coord_2d, coord_3d, offsets, counts, K,
                                            reinterpret_cast<const unsigned long long*>(seeds), *params, labeling, poses);
  EPOS_LAUNCH_CHECK();
  progx_kernel<<<P, THREADS, SMEM_FIT, s>>>(ws, *params, mp, offsets, max_models, poses, labeling, multi_poses, multi_scores,
                                            multi_counts);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

// host copy of the per-problem Progressive-X counters of the last epos_fit_poses_multi (synchronous):
// out [P][8] i64: proposals, accepted, total RANSAC iterations, PEARL iterations, expansion moves, instances, unaccepted, 0
int epos_fit_multi_debug_state(const void* workspace, int P, long long* out) {
  EPOS_CHECK_ARG(workspace && out && P > 0);
  Workspace ws;
  workspace_layout(P, const_cast<void*>(workspace), &ws, true);
  MultiState* h = (MultiState*)malloc((size_t)P * sizeof(MultiState));
  if (!h) return EPOS_ERR_CUDA;
  cudaError_t e = cudaMemcpy(h, ws.ms, (size_t)P * sizeof(MultiState), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { free(h); set_error("epos_fit_multi_debug_state: %s", cudaGetErrorString(e)); return EPOS_ERR_CUDA; }
  for (int i = 0; i < P; ++i) {
    long long* o = out + (size_t)i * 8;
    o[0] = h[i].proposals; o[1] = h[i].accepted; o[2] = h[i].total_iterations; o[3] = h[i].pearl_iterations;
    o[4] = h[i].moves; o[5] = h[i].n_models; o[6] = h[i].unaccepted; o[7] = 0;
  }
  free(h);
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
