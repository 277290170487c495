// Group-cooperative non-minimal PnP solver in f64: the behaviour of cv::solvePnP(SOLVEPNP_ITERATIVE) with K = I as
// called by the reference at
//   .../pygcransac/include/solver_epnp_lm.h:139-146   (useExtrinsicGuess = false: DLT / planar init + LM)
//   external/progressive-x/src/pyprogressivex/src/progressivex_python.cpp:292-299 (useExtrinsicGuess = true: LM only)
// i.e. OpenCV's cvFindExtrinsicCameraParams2 + CvLevMarq(6, 2n, {EPS+ITER, 20, FLT_EPSILON}).  A "group" is either one
// warp (LO inner fits on <= 21 points, one fit per warp) or the whole CTA (fits on the full inlier set).
#pragma once
#include "pose_math.cuh"

namespace epos {
namespace pose {

constexpr int FIT_SCRATCH_DOUBLES = 12 * 12 * 2 + 12 + 28;   // A, V, w, reduction output
constexpr int WARP_FIT_MAX = 21;                             // points a warp-level fit handles (one per lane, 14 doubles each)
constexpr int FIT_NRED = 28;

struct WarpGroup {
  int rank, size;
  double* sh;            // FIT_SCRATCH_DOUBLES doubles of shared memory private to this warp
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  // v[n] per-thread partial sums -> out[n] (shared, visible to the whole group after return)
  __device__ inline void reduce(double* v, int n, double* out) const {
    for (int k = 0; k < n; ++k) {
      double x = v[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (rank == 0) out[k] = x;
    }
    __syncwarp();
  }
  __device__ __forceinline__ bool leader_warp() const { return true; }
  __device__ __forceinline__ int lane() const { return rank; }
};

// A warp working on MORE points than it has lanes: same interface as WarpGroup, but a distinct type so that the generic
// (lane-strided) lm_eval applies instead of the one-point-per-lane specialisation.  Used for the final fits on a few dozen
// to a few hundred inliers, where a single warp without block barriers beats the whole CTA with two barriers per
// reduction.
struct WarpGroupN : WarpGroup {};

struct CtaGroup {
  int rank, size;
  double* sh;            // FIT_SCRATCH_DOUBLES doubles
  double* part;          // (blockDim.x / 32) * FIT_NRED doubles
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ inline void reduce(double* v, int n, double* out) const {
    const int w = rank >> 5, l = rank & 31, nw = size >> 5;
    for (int k = 0; k < n; ++k) {
      double x = v[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (l == 0) part[w * FIT_NRED + k] = x;
    }
    __syncthreads();
    if (rank < n) {
      double s = 0.0;
      for (int i = 0; i < nw; ++i) s += part[i * FIT_NRED + rank];
      out[rank] = s;
    }
    __syncthreads();
  }
  __device__ __forceinline__ bool leader_warp() const { return (rank >> 5) == 0; }
  __device__ __forceinline__ int lane() const { return rank & 31; }
};

// Points are addressed through an index list into SoA arrays (un, vn, x, y, z).
struct PointView {
  const double* un; const double* vn; const double* x; const double* y; const double* z;
  const unsigned short* idx;   // may be nullptr: identity
  int n;
  __device__ __forceinline__ void get(int i, double* X, double* uv) const {
    const int j = idx ? (int)idx[i] : i;
    X[0] = x[j]; X[1] = y[j]; X[2] = z[j];
    uv[0] = un[j]; uv[1] = vn[j];
  }
};

template <class G>
__device__ inline void group_eig(const G& g, int n, double* A, double* V, double* w) {
  g.sync();
  if (g.leader_warp()) jacobi_eig_warp(n, A, V, w, g.lane());
  g.sync();
}

// normalised-DLT homography dst ~ H src, src = planar object coords (through Rt/Tt), dst = image points
template <class G>
__device__ inline bool homography_group(const G& g, const PointView& pv, const double* Rt, const double* Tt, double* H) {
  double* A = g.sh; double* V = g.sh + 144; double* w = g.sh + 288; double* red = g.sh + 300;
  const int n = pv.n;
  double acc[9];
  // centroids
  for (int k = 0; k < 4; ++k) acc[k] = 0.0;
  for (int i = g.rank; i < n; i += g.size) {
    double X[3], uv[2];
    pv.get(i, X, uv);
    acc[0] += Rt[0] * X[0] + Rt[1] * X[1] + Rt[2] * X[2] + Tt[0];
    acc[1] += Rt[3] * X[0] + Rt[4] * X[1] + Rt[5] * X[2] + Tt[1];
    acc[2] += uv[0]; acc[3] += uv[1];
  }
  g.reduce(acc, 4, red);
  const double cs0 = red[0] / n, cs1 = red[1] / n, cd0 = red[2] / n, cd1 = red[3] / n;
  g.sync();
  acc[0] = acc[1] = 0.0;
  for (int i = g.rank; i < n; i += g.size) {
    double X[3], uv[2];
    pv.get(i, X, uv);
    const double sx = Rt[0] * X[0] + Rt[1] * X[1] + Rt[2] * X[2] + Tt[0] - cs0;
    const double sy = Rt[3] * X[0] + Rt[4] * X[1] + Rt[5] * X[2] + Tt[1] - cs1;
    acc[0] += sqrt(sx * sx + sy * sy);
    acc[1] += sqrt((uv[0] - cd0) * (uv[0] - cd0) + (uv[1] - cd1) * (uv[1] - cd1));
  }
  g.reduce(acc, 2, red);
  double ss = red[0], sd = red[1];
  g.sync();
  if (ss == 0.0 || sd == 0.0) return false;
  ss = sqrt(2.0) * n / ss; sd = sqrt(2.0) * n / sd;
  // A = sum r1 r1^T + r2 r2^T (9x9), one row-block (9 entries) per sweep
  for (int a = 0; a < 9; ++a) {
    for (int b = 0; b < 9; ++b) acc[b] = 0.0;
    for (int i = g.rank; i < n; i += g.size) {
      double X[3], uv[2];
      pv.get(i, X, uv);
      const double x = (Rt[0] * X[0] + Rt[1] * X[1] + Rt[2] * X[2] + Tt[0] - cs0) * ss;
      const double y = (Rt[3] * X[0] + Rt[4] * X[1] + Rt[5] * X[2] + Tt[1] - cs1) * ss;
      const double Xd = (uv[0] - cd0) * sd, Yd = (uv[1] - cd1) * sd;
      const double r1[9] = {x, y, 1, 0, 0, 0, -Xd * x, -Xd * y, -Xd};
      const double r2[9] = {0, 0, 0, x, y, 1, -Yd * x, -Yd * y, -Yd};
      for (int b = 0; b < 9; ++b) acc[b] += r1[a] * r1[b] + r2[a] * r2[b];
    }
    g.reduce(acc, 9, red);
    if (g.rank < 9) A[a * 9 + g.rank] = red[g.rank];
    g.sync();
  }
  group_eig(g, 9, A, V, w);
  int k = 0;
  for (int i = 1; i < 9; ++i)
    if (w[i] < w[k]) k = i;
  double Hn[9];
  for (int i = 0; i < 9; ++i) Hn[i] = V[i * 9 + k];
  g.sync();
  const double Ts[9] = {ss, 0, -ss * cs0, 0, ss, -ss * cs1, 0, 0, 1};
  const double Tdi[9] = {1 / sd, 0, cd0, 0, 1 / sd, cd1, 0, 0, 1};
  double tmp[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) tmp[i * 3 + j] = Hn[i * 3] * Ts[j] + Hn[i * 3 + 1] * Ts[3 + j] + Hn[i * 3 + 2] * Ts[6 + j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) H[i * 3 + j] = Tdi[i * 3] * tmp[j] + Tdi[i * 3 + 1] * tmp[3 + j] + Tdi[i * 3 + 2] * tmp[6 + j];
  if (H[8] != 0.0) {
    const double h8 = H[8];
    for (int i = 0; i < 9; ++i) H[i] /= h8;
  }
  bool fin = true;
  for (int i = 0; i < 9; ++i) fin &= isfinite(H[i]);
  return fin;
}

// Initial pose (cvFindExtrinsicCameraParams2 without extrinsic guess).  param (registers, identical in every thread).
template <class G>
__device__ inline bool pnp_init_group(const G& g, const PointView& pv, double* param) {
  double* A = g.sh; double* V = g.sh + 144; double* w = g.sh + 288; double* red = g.sh + 300;
  const int n = pv.n;
  double acc[10];
  acc[0] = acc[1] = acc[2] = 0.0;
  for (int i = g.rank; i < n; i += g.size) {
    double X[3], uv[2];
    pv.get(i, X, uv);
    acc[0] += X[0]; acc[1] += X[1]; acc[2] += X[2];
  }
  g.reduce(acc, 3, red);
  const double Mc[3] = {red[0] / n, red[1] / n, red[2] / n};
  g.sync();
  for (int k = 0; k < 6; ++k) acc[k] = 0.0;
  for (int i = g.rank; i < n; i += g.size) {
    double X[3], uv[2];
    pv.get(i, X, uv);
    const double a = X[0] - Mc[0], b = X[1] - Mc[1], c = X[2] - Mc[2];
    acc[0] += a * a; acc[1] += a * b; acc[2] += a * c; acc[3] += b * b; acc[4] += b * c; acc[5] += c * c;
  }
  g.reduce(acc, 6, red);
  double MM[9] = {red[0], red[1], red[2], red[1], red[3], red[4], red[2], red[4], red[5]};
  g.sync();
  double V3[9], w3[3];
  jacobi_eig3(MM, V3, w3);
  int o0 = 0, o1 = 1, o2 = 2;                         // indices by decreasing eigenvalue
  if (w3[o0] < w3[o1]) { int t = o0; o0 = o1; o1 = t; }
  if (w3[o1] < w3[o2]) { int t = o1; o1 = o2; o2 = t; }
  if (w3[o0] < w3[o1]) { int t = o0; o0 = o1; o1 = t; }
  double R[9], t[3];
  if (w3[o2] / w3[o1] < 1e-3) {
    const int ord[3] = {o0, o1, o2};
    double Rt[9];
    for (int r = 0; r < 3; ++r)
      for (int k = 0; k < 3; ++k) Rt[r * 3 + k] = V3[k * 3 + ord[r]];
    if (Rt[2] * Rt[2] + Rt[5] * Rt[5] < 1e-10)
      for (int i = 0; i < 9; ++i) Rt[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (det3(Rt) < 0)
      for (int i = 0; i < 9; ++i) Rt[i] = -Rt[i];
    double Tt[3];
    for (int r = 0; r < 3; ++r) Tt[r] = -(Rt[r * 3] * Mc[0] + Rt[r * 3 + 1] * Mc[1] + Rt[r * 3 + 2] * Mc[2]);
    double H[9];
    if (n >= 4 && homography_group(g, pv, Rt, Tt, H)) {
      double h1[3] = {H[0], H[3], H[6]}, h2[3] = {H[1], H[4], H[7]}, h3[3] = {H[2], H[5], H[8]};
      const double n1 = sqrt(dot3(h1, h1)), n2 = sqrt(dot3(h2, h2));
      for (int k = 0; k < 3; ++k) { h1[k] /= fmax(n1, DBL_EPSILON); h2[k] /= fmax(n2, DBL_EPSILON); }
      double tt[3];
      for (int k = 0; k < 3; ++k) tt[k] = h3[k] * (2.0 / fmax(n1 + n2, DBL_EPSILON));
      cross3(h1, h2, h3);
      const double Hm[9] = {h1[0], h2[0], h3[0], h1[1], h2[1], h3[1], h1[2], h2[2], h3[2]};
      double rv[3], Hr[9];
      matrix_to_rodrigues(Hm, rv);
      rodrigues_to_matrix(rv, Hr, nullptr);
      for (int r = 0; r < 3; ++r) t[r] = Hr[r * 3] * Tt[0] + Hr[r * 3 + 1] * Tt[1] + Hr[r * 3 + 2] * Tt[2] + tt[r];
      for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k) R[r * 3 + k] = Hr[r * 3] * Rt[k] + Hr[r * 3 + 1] * Rt[3 + k] + Hr[r * 3 + 2] * Rt[6 + k];
    } else {
      for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
      t[0] = t[1] = t[2] = 0.0;
    }
  } else {
    if (n < 6) return false;
    // L^T L = [[S1, 0, Sx], [0, S1, Sy], [Sx, Sy, Sq]] with S* = sum w * Mh Mh^T, Mh = (X,1), x = -u, y = -v
    for (int sweep = 0; sweep < 4; ++sweep) {
      for (int k = 0; k < 10; ++k) acc[k] = 0.0;
      for (int i = g.rank; i < n; i += g.size) {
        double X[3], uv[2];
        pv.get(i, X, uv);
        const double x = -uv[0], y = -uv[1];
        const double wgt = sweep == 0 ? 1.0 : (sweep == 1 ? x : (sweep == 2 ? y : x * x + y * y));
        const double M[4] = {X[0], X[1], X[2], 1.0};
        int k = 0;
        for (int a = 0; a < 4; ++a)
          for (int b = a; b < 4; ++b) acc[k++] += wgt * M[a] * M[b];
      }
      g.reduce(acc, 10, red);
      if (g.rank == 0) {
        int k = 0;
        for (int a = 0; a < 4; ++a)
          for (int b = a; b < 4; ++b) {
            const double v = red[k++];
            if (sweep == 0) {
              A[a * 12 + b] = A[b * 12 + a] = v;
              A[(4 + a) * 12 + 4 + b] = A[(4 + b) * 12 + 4 + a] = v;
              A[a * 12 + 4 + b] = A[b * 12 + 4 + a] = 0.0;
              A[(4 + a) * 12 + b] = A[(4 + b) * 12 + a] = 0.0;
            } else if (sweep == 1) {
              A[a * 12 + 8 + b] = A[b * 12 + 8 + a] = v;
              A[(8 + a) * 12 + b] = A[(8 + b) * 12 + a] = v;
            } else if (sweep == 2) {
              A[(4 + a) * 12 + 8 + b] = A[(4 + b) * 12 + 8 + a] = v;
              A[(8 + a) * 12 + 4 + b] = A[(8 + b) * 12 + 4 + a] = v;
            } else {
              A[(8 + a) * 12 + 8 + b] = A[(8 + b) * 12 + 8 + a] = v;
            }
          }
      }
      g.sync();
    }
    group_eig(g, 12, A, V, w);
    int k = 0;
    for (int i = 1; i < 12; ++i)
      if (w[i] < w[k]) k = i;
    double RRt[12];
    for (int i = 0; i < 12; ++i) RRt[i] = V[i * 12 + k];
    g.sync();
    double RR[9] = {RRt[0], RRt[1], RRt[2], RRt[4], RRt[5], RRt[6], RRt[8], RRt[9], RRt[10]};
    double tt[3] = {RRt[3], RRt[7], RRt[11]};
    if (det3(RR) < 0) {
      for (int i = 0; i < 9; ++i) RR[i] = -RR[i];
      for (int i = 0; i < 3; ++i) tt[i] = -tt[i];
    }
    double sc = 0;
    for (int i = 0; i < 9; ++i) sc += RR[i] * RR[i];
    sc = sqrt(sc);
    if (!(sc > DBL_EPSILON)) return false;
    if (!polar_rotation(RR, R)) return false;
    double nr = 0;
    for (int i = 0; i < 9; ++i) nr += R[i] * R[i];
    nr = sqrt(nr);
    for (int i = 0; i < 3; ++i) t[i] = tt[i] * (nr / sc);
  }
  matrix_to_rodrigues(R, param);
  param[3] = t[0]; param[4] = t[1]; param[5] = t[2];
  bool fin = true;
  for (int i = 0; i < 6; ++i) fin &= isfinite(param[i]);
  return fin;
}

// One evaluation of the reprojection residuals (and optionally J^T J, J^T e) at `param`; results in red[0..27]:
// [0..20] upper triangle of JtJ (row-major), [21..26] JtErr, [27] sum err^2.
template <class G>
__device__ inline void lm_eval(const G& g, const PointView& pv, const double* param, bool with_jac, double* red) {
  double R[9], dR[27];
  rodrigues_to_matrix(param, R, with_jac ? dR : nullptr);
  double acc[FIT_NRED];
  const int nacc = with_jac ? FIT_NRED : 1;
  for (int k = 0; k < FIT_NRED; ++k) acc[k] = 0.0;
  for (int i = g.rank; i < pv.n; i += g.size) {
    double p[3], uv[2];
    pv.get(i, p, uv);
    const double Y0 = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + param[3];
    const double Y1 = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + param[4];
    const double Y2 = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + param[5];
    const double iz = Y2 != 0.0 ? 1.0 / Y2 : 1.0;
    const double x = Y0 * iz, y = Y1 * iz;
    const double ex = x - uv[0], ey = y - uv[1];
    if (with_jac) {
      double j0[6], j1[6];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double* d = dR + 9 * k;
        const double dY0 = d[0] * p[0] + d[1] * p[1] + d[2] * p[2];
        const double dY1 = d[3] * p[0] + d[4] * p[1] + d[5] * p[2];
        const double dY2 = d[6] * p[0] + d[7] * p[1] + d[8] * p[2];
        j0[k] = iz * (dY0 - x * dY2);
        j1[k] = iz * (dY1 - y * dY2);
      }
      j0[3] = iz; j0[4] = 0.0; j0[5] = -x * iz;
      j1[3] = 0.0; j1[4] = iz; j1[5] = -y * iz;
      int k = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b) acc[k++] += j0[a] * j0[b] + j1[a] * j1[b];
#pragma unroll
      for (int a = 0; a < 6; ++a) acc[21 + a] += j0[a] * ex + j1[a] * ey;
      acc[27] += ex * ex + ey * ey;
    } else {
      acc[0] += ex * ex + ey * ey;
    }
  }
  g.reduce(acc, nacc, red);
}

// Warp-level evaluation for n <= WARP_FIT_MAX points (one point per lane): every lane stores its 14 per-point
// quantities (j0[6], j1[6], ex, ey) in the warp's scratch (the Jacobi A/V/w area, idle during LM), then lanes 0..27 each
// form one of the 28 sums.
template <>
__device__ inline void lm_eval<WarpGroup>(const WarpGroup& g, const PointView& pv, const double* param, bool with_jac,
                                          double* red) {
  double R[9], dR[27];
  rodrigues_to_matrix(param, R, with_jac ? dR : nullptr);
  double* rows = g.sh;                         // [n][14] <= 294 doubles (A + V + w = 300)
  const int n = pv.n, lane = g.rank;
  double e2 = 0.0;
  if (lane < n) {
    double p[3], uv[2];
    pv.get(lane, p, uv);
    const double Y0 = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + param[3];
    const double Y1 = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + param[4];
    const double Y2 = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + param[5];
    const double iz = Y2 != 0.0 ? 1.0 / Y2 : 1.0;
    const double x = Y0 * iz, y = Y1 * iz;
    const double ex = x - uv[0], ey = y - uv[1];
    e2 = ex * ex + ey * ey;
    if (with_jac) {
      double* r = rows + lane * 14;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double* d = dR + 9 * k;
        const double dY0 = d[0] * p[0] + d[1] * p[1] + d[2] * p[2];
        const double dY1 = d[3] * p[0] + d[4] * p[1] + d[5] * p[2];
        const double dY2 = d[6] * p[0] + d[7] * p[1] + d[8] * p[2];
        r[k] = iz * (dY0 - x * dY2);
        r[6 + k] = iz * (dY1 - y * dY2);
      }
      r[3] = iz; r[4] = 0.0; r[5] = -x * iz;
      r[9] = 0.0; r[10] = iz; r[11] = -y * iz;
      r[12] = ex; r[13] = ey;
    }
  }
  if (!with_jac) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e2 += __shfl_xor_sync(0xffffffffu, e2, o);
    if (lane == 0) red[0] = e2;
    __syncwarp();
    return;
  }
  __syncwarp();
  if (lane < FIT_NRED) {
    // sum_i r[ia0] r[ib0] + r[ia1] r[ib1]: JtJ(a,b) -> (a, b | 6+a, 6+b); JtErr(a) -> (a, 12 | 6+a, 13); err^2 -> (12,12|13,13)
    int ia0, ib0, ia1, ib1;
    if (lane < 21) {
      int k = lane, a = 0, b = 0;
      for (a = 0; a < 6; ++a) { const int len = 6 - a; if (k < len) { b = a + k; break; } k -= len; }
      ia0 = a; ib0 = b; ia1 = 6 + a; ib1 = 6 + b;
    } else if (lane < 27) {
      ia0 = lane - 21; ib0 = 12; ia1 = 6 + lane - 21; ib1 = 13;
    } else {
      ia0 = 12; ib0 = 12; ia1 = 13; ib1 = 13;
    }
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
      const double* r = rows + i * 14;
      s += r[ia0] * r[ib0] + r[ia1] * r[ib1];
    }
    red[lane] = s;
  }
  __syncwarp();
}

__device__ inline bool lm_step(const double* red, const double* prev, int lambdaLg10, double* param) {
  double A[36], b[6], x[6];
  const double lambda = exp(lambdaLg10 * log(10.0));
  int k = 0;
  for (int a = 0; a < 6; ++a)
    for (int c = a; c < 6; ++c) { A[a * 6 + c] = red[k]; A[c * 6 + a] = red[k]; ++k; }
  for (int a = 0; a < 6; ++a) { b[a] = red[21 + a]; A[a * 6 + a] *= 1.0 + lambda; }
  if (!solve_linear6(A, b, x)) return false;
  for (int i = 0; i < 6; ++i) param[i] = prev[i] - x[i];
  return true;
}

// CvLevMarq loop of cvFindExtrinsicCameraParams2 (max 20 iterations, eps FLT_EPSILON).  param in/out (registers).
template <class G>
__device__ inline void lm_refine_group(const G& g, const PointView& pv, double* param) {
  double* red = g.sh + 300;
  const int max_iter = 20;
  const double eps = (double)FLT_EPSILON;
  int lambdaLg10 = -3, iters = 0;
  double prevErrNorm = DBL_MAX;
  for (;;) {
    lm_eval(g, pv, param, true, red);
    double jt[FIT_NRED];
    for (int k = 0; k < FIT_NRED; ++k) jt[k] = red[k];
    g.sync();
    double prev[6];
    for (int i = 0; i < 6; ++i) prev[i] = param[i];
    if (!lm_step(jt, prev, lambdaLg10, param)) {
      for (int i = 0; i < 6; ++i) param[i] = prev[i];
      return;
    }
    if (iters == 0) prevErrNorm = sqrt(jt[27]);
    double errNorm;
    for (;;) {
      lm_eval(g, pv, param, false, red);
      errNorm = sqrt(red[0]);
      g.sync();
      if (errNorm > prevErrNorm && ++lambdaLg10 <= 16) {
        if (!lm_step(jt, prev, lambdaLg10, param)) {
          for (int i = 0; i < 6; ++i) param[i] = prev[i];
          return;
        }
        continue;
      }
      break;
    }
    lambdaLg10 = lambdaLg10 - 1 > -16 ? lambdaLg10 - 1 : -16;
    double dn = 0, pn = 0;
    for (int i = 0; i < 6; ++i) { dn += (param[i] - prev[i]) * (param[i] - prev[i]); pn += prev[i] * prev[i]; }
    const double rel = sqrt(dn) / (pn > 0 ? sqrt(pn) : DBL_MIN);
    if (++iters >= max_iter || rel < eps) return;
    prevErrNorm = errNorm;
  }
}

// EPnPLM::estimateModel + filter of estimateModelNonminimal (perspective_n_point_estimator.h:240-268).
// model: row-major 3x4 (registers, identical in every thread of the group).
template <class G>
__device__ inline bool fit_nonminimal_group(const G& g, const PointView& pv, double* model) {
  if (pv.n < 4) return false;                      // cv::solvePnP needs >= 4 points (n == 3 only with a guess)
  double param[6];
  if (!pnp_init_group(g, pv, param)) return false;
  lm_refine_group(g, pv, param);
  bool fin = true;
  for (int i = 0; i < 6; ++i) fin &= isfinite(param[i]);
  if (!fin) return false;
  double R[9];
  rodrigues_to_matrix(param, R, nullptr);
  if (param[5] < 0.0 || det3(R) < -0.95) return false;
  for (int r = 0; r < 3; ++r) {
    model[r * 4] = R[r * 3]; model[r * 4 + 1] = R[r * 3 + 1]; model[r * 4 + 2] = R[r * 3 + 2]; model[r * 4 + 3] = param[3 + r];
  }
  return true;
}

}  // namespace pose
}  // namespace epos
