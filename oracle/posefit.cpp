// ORACLE -- test infrastructure, never linked or imported by the product (epos_b200/).
//
// CPU restatement (plain C++17, no third-party dependency) of the single-instance branch of
// pyprogressivex.find6DPoses:
//   /root/reference/external/progressive-x/src/pyprogressivex/src/bindings.cpp:9-118
//   /root/reference/external/progressive-x/src/pyprogressivex/src/progressivex_python.cpp:36-134,222-336
// and of the GC-RANSAC machinery under
//   /root/reference/external/progressive-x/graph-cut-ransac/src/pygcransac/include/  (paths below relative to it)
//     GCRANSAC.h:206-530 (run), :533-657 (iteratedLeastSquaresFitting), :679-809 (graphCutLocalOptimization),
//     :812-920 (labeling); scoring_function.h:178-268 (EPOSScoringFunction);
//     perspective_n_point_estimator.h:56-271; solver_p3p.h:111-371 (Kneip P3P); solver_epnp_lm.h:91-161;
//     uniform_sampler.h:84-102, uniform_random_generator.h:43-120; flann_neighborhood_graph.h:86-125;
//     energy.h:204-253 (add_term1 / add_term2 reparameterisation); settings.h:68-88.
//
// Third-party arithmetic that is NOT under /root/reference and is restated here from its published algorithm:
//   * OpenCV 3.4.2 cv::solvePnP(SOLVEPNP_ITERATIVE) = cvFindExtrinsicCameraParams2 (DLT or planar-homography
//     initialisation + CvLevMarq, <= 20 iterations, eps = FLT_EPSILON) and cv::Rodrigues.  Pinned by
//     tests/golden/cv2_solvepnp.json (outputs of the cv2 4.13 wheel) and the 16-correspondence notebook vector.
//     Deviations: the planar branch initialises from a normalised-DLT homography without OpenCV's extra LM polish
//     of H (the pose LM that follows converges to the same minimum); a non-planar fit with < 6 points fails
//     instead of returning an ill-defined null-space vector.
//   * Eigen::PolynomialSolver<double,4>::realRoots (solver_p3p.h:175-177): restated as Ferrari + Newton polish,
//     real roots in ascending order (Eigen's order is an implementation detail of its QR iteration: unpinned).
//   * Boykov-Kolmogorov max-flow (graph.h / maxflow.cpp): the labeling only needs "which nodes can still reach
//     the sink in the residual graph of a maximum flow" (graph.h:112-115,478-488), which is the same for every
//     maximum flow; restated with Dinic + reverse BFS and validated against the reference's own BK sources
//     compiled into oracle/_ref (tests/test_oracle_pose.py).
//   * std::mt19937 seeded from std::random_device (uniform_random_generator.h:51-54) has no reproducible stream;
//     replaced by the counter-based generator documented in DESIGN.md ("RANSAC random stream") that the CUDA
//     kernels implement independently.
//   * cv::FlannBasedMatcher(KDTree(4), checks=6).radiusMatch is randomised/approximate (<= 5 neighbours per point
//     measured); determinised as "the max_neighbors nearest points within the radius,
//     f32 L2 in the 5-D space (u, v, s*x, s*y, s*z), ties by index".  Neighbour lists can also be injected.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <queue>
#include <set>
#include <utility>
#include <vector>

namespace ora {

typedef unsigned long long u64;

struct Params {
  double threshold;                 // px
  double spatial_coherence_weight;  // lambda
  double neighborhood_ball_radius;
  double scaling_from_millimeters;
  double min_triangle_area;
  double min_coverage;
  double confidence;                // proposal_engine_conf (1.0 in EPOS)
  int max_iters, min_iters, min_iters_before_lo, max_lo_trials, max_graph_cuts, max_lsq_iters, max_unsuccessful,
      max_neighbors, apply_numerical_optimization;
};

// ---------------------------------------------------------------------------------------------------
// Counter-based random stream: value = mix(mix(mix(mix(seed ^ stream*C) + a) + b) + c)
// ---------------------------------------------------------------------------------------------------
static inline u64 mix64(u64 z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static inline u64 rng_u64(u64 seed, u64 stream, u64 a, u64 b, u64 c) {
  return mix64(mix64(mix64(mix64(seed ^ (stream * 0xD6E8FEB86659FD93ULL)) + a) + b) + c);
}
static inline u64 rng_index(u64 seed, u64 stream, u64 a, u64 b, u64 c, u64 n) {
  return (u64)(((unsigned __int128)rng_u64(seed, stream, a, b, c) * (unsigned __int128)n) >> 64);
}
// k distinct indices in [0, n): draw c = 0,1,2,...; a draw equal to an already accepted one is discarded
// (uniform_random_generator.h:85-97).
static bool unique_set(u64 seed, u64 stream, u64 a, u64 b, int n, int k, int* out) {
  if (k > n) return false;
  u64 c = 0;
  for (int i = 0; i < k;) {
    int v = (int)rng_index(seed, stream, a, b, c++, (u64)n);
    bool dup = false;
    for (int j = 0; j < i; ++j) dup |= (out[j] == v);
    if (!dup) out[i++] = v;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------
// small linear algebra
// ---------------------------------------------------------------------------------------------------
static inline void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void normalize3(double* a) {
  double n = std::sqrt(dot3(a, a));
  a[0] /= n; a[1] /= n; a[2] /= n;
}
static inline double det3(const double* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
static bool inv3(const double* m, double* o) {
  double d = det3(m);
  if (d == 0.0) return false;
  double id = 1.0 / d;
  o[0] = (m[4] * m[8] - m[5] * m[7]) * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = (m[5] * m[6] - m[3] * m[8]) * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = (m[3] * m[7] - m[4] * m[6]) * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return true;
}

// Jacobi eigen-decomposition of a symmetric n x n matrix (row-major A, destroyed); V columns = eigenvectors.
// Rotations are applied in the round-robin ("tournament") parallel ordering: each of the m-1 steps of a sweep
// (m = n rounded up to even) rotates n/2 disjoint index pairs at once, A <- J^T A J with J the product of the step's
// rotations, all angles taken from A before the step.  (Stands in for the SVDs inside cvFindExtrinsicCameraParams2 /
// cvFindHomography; the CUDA kernels use the same ordering so that both sides take the same path.)
static void jacobi_eig(int n, double* A, double* V, double* w) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  const int m = (n + 1) & ~1;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i == j ? diag : off) += A[i * n + j] * A[i * n + j];
    if (off <= 1e-300 || off < 1e-28 * diag) break;
    for (int step = 0; step < m - 1; ++step) {
      int P[8], Q[8], np = 0;
      double C[8], S[8];
      for (int k = 0; k < m / 2; ++k) {
        int a = k == 0 ? m - 1 : (step + k) % (m - 1);
        int b = k == 0 ? step : (step - k + (m - 1)) % (m - 1);
        int p = a < b ? a : b, q = a < b ? b : a;
        if (q >= n) continue;
        double apq = A[p * n + q], c = 1.0, s_ = 0.0;
        if (apq != 0.0) {
          double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
          double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
          c = 1.0 / std::sqrt(t * t + 1.0);
          s_ = t * c;
        }
        P[np] = p; Q[np] = q; C[np] = c; S[np] = s_; ++np;
      }
      for (int j = 0; j < np; ++j)                                  // A <- A J, V <- V J
        for (int k = 0; k < n; ++k) {
          double akp = A[k * n + P[j]], akq = A[k * n + Q[j]];
          A[k * n + P[j]] = C[j] * akp - S[j] * akq;
          A[k * n + Q[j]] = S[j] * akp + C[j] * akq;
          double vkp = V[k * n + P[j]], vkq = V[k * n + Q[j]];
          V[k * n + P[j]] = C[j] * vkp - S[j] * vkq;
          V[k * n + Q[j]] = S[j] * vkp + C[j] * vkq;
        }
      for (int j = 0; j < np; ++j)                                  // A <- J^T A
        for (int k = 0; k < n; ++k) {
          double apk = A[P[j] * n + k], aqk = A[Q[j] * n + k];
          A[P[j] * n + k] = C[j] * apk - S[j] * aqk;
          A[Q[j] * n + k] = S[j] * apk + C[j] * aqk;
        }
    }
  }
  for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
}

// Gaussian elimination with partial pivoting, n <= 8.  Returns false if singular.
static bool solve_linear(int n, double* A, double* b, double* x) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(A[r * n + c]) > std::fabs(A[piv * n + c])) piv = r;
    if (A[piv * n + c] == 0.0) return false;
    if (piv != c) {
      for (int k = 0; k < n; ++k) std::swap(A[c * n + k], A[piv * n + k]);
      std::swap(b[c], b[piv]);
    }
    for (int r = c + 1; r < n; ++r) {
      double f = A[r * n + c] / A[c * n + c];
      for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < n; ++k) s -= A[r * n + k] * x[k];
    x[r] = s / A[r * n + r];
  }
  return true;
}

// Orthogonal polar factor U V^T of a 3x3 matrix with positive determinant (Newton iteration
// R <- (R + R^-T)/2): what cvSVD + U V^T produce in cvFindExtrinsicCameraParams2.
static bool polar_rotation(const double* M, double* R) {
  std::memcpy(R, M, 9 * sizeof(double));
  for (int it = 0; it < 100; ++it) {
    double inv[9];
    if (!inv3(R, inv)) return false;
    double diff = 0.0;
    double Rn[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        Rn[i * 3 + j] = 0.5 * (R[i * 3 + j] + inv[j * 3 + i]);
        diff += (Rn[i * 3 + j] - R[i * 3 + j]) * (Rn[i * 3 + j] - R[i * 3 + j]);
      }
    std::memcpy(R, Rn, sizeof(Rn));
    if (diff < 1e-30) break;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------
// cv::Rodrigues (both directions) and d R / d rvec
// ---------------------------------------------------------------------------------------------------
static void rodrigues_to_matrix(const double* r, double* R, double* dRdr /* 3 x 9 or null */) {
  double theta = std::sqrt(dot3(r, r));
  if (theta < std::numeric_limits<double>::epsilon()) {
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (dRdr) {
      std::memset(dRdr, 0, 27 * sizeof(double));
      dRdr[5] = -1; dRdr[7] = 1; dRdr[9 + 2] = 1; dRdr[9 + 6] = -1; dRdr[18 + 1] = -1; dRdr[18 + 3] = 1;
    }
    return;
  }
  double c = std::cos(theta), s = std::sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
  double k[3] = {r[0] * it, r[1] * it, r[2] * it};
  double kkt[9] = {k[0] * k[0], k[0] * k[1], k[0] * k[2], k[0] * k[1], k[1] * k[1], k[1] * k[2],
                   k[0] * k[2], k[1] * k[2], k[2] * k[2]};
  double kx[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) R[i] = c * I[i] + c1 * kkt[i] + s * kx[i];
  if (dRdr) {
    // d/dr_i of (c I + c1 k k^T + s [k]x) with k = r/theta
    const double dkkt[27] = {2 * k[0], k[1], k[2], k[1], 0, 0, k[2], 0, 0,
                             0, k[0], 0, k[0], 2 * k[1], k[2], 0, k[2], 0,
                             0, 0, k[0], 0, 0, k[1], k[0], k[1], 2 * k[2]};
    const double dkx[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
    for (int i = 0; i < 3; ++i) {
      double ki = k[i];
      double a0 = -s * ki, a1 = (s - 2 * c1 * it) * ki, a2 = c1 * it, a3 = (c - s * it) * ki, a4 = s * it;
      for (int j = 0; j < 9; ++j)
        dRdr[i * 9 + j] = a0 * I[j] + a1 * kkt[j] + a2 * dkkt[i * 9 + j] + a3 * kx[j] + a4 * dkx[i * 9 + j];
    }
  }
}

static void matrix_to_rodrigues(const double* Rin, double* r) {
  double R[9];
  if (!polar_rotation(Rin, R)) std::memcpy(R, Rin, sizeof(R));   // cvRodrigues2 re-orthogonalises with an SVD
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  double theta = std::acos(c);
  if (s < 1e-5) {
    if (c > 0) { r[0] = r[1] = r[2] = 0.0; return; }
    double t;
    t = (R[0] + 1) * 0.5; rx = std::sqrt(std::max(t, 0.0));
    t = (R[4] + 1) * 0.5; ry = std::sqrt(std::max(t, 0.0)) * (R[1] < 0 ? -1.0 : 1.0);
    t = (R[8] + 1) * 0.5; rz = std::sqrt(std::max(t, 0.0)) * (R[2] < 0 ? -1.0 : 1.0);
    if (std::fabs(rx) < std::fabs(ry) && std::fabs(rx) < std::fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
    theta /= std::sqrt(rx * rx + ry * ry + rz * rz);
    r[0] = rx * theta; r[1] = ry * theta; r[2] = rz * theta;
    return;
  }
  double vth = theta / (2.0 * s);
  r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
}

// ---------------------------------------------------------------------------------------------------
// Real roots of c[4] x^4 + c[3] x^3 + c[2] x^2 + c[1] x + c[0] (Ferrari + Newton polish), ascending.
// ---------------------------------------------------------------------------------------------------
static double cubic_largest_real_root(double A, double B, double C) {   // x^3 + A x^2 + B x + C
  double a3 = A / 3.0;
  double p = B - A * a3, q = 2.0 * a3 * a3 * a3 - a3 * B + C;           // t^3 + p t + q, x = t - A/3
  double disc = q * q / 4.0 + p * p * p / 27.0;
  double t;
  if (disc > 0) {
    double sq = std::sqrt(disc);
    t = std::cbrt(-q / 2.0 + sq) + std::cbrt(-q / 2.0 - sq);
  } else if (p == 0.0) {
    t = std::cbrt(-q);
  } else {
    double m = 2.0 * std::sqrt(-p / 3.0);
    double arg = 3.0 * q / (p * m);
    arg = arg > 1.0 ? 1.0 : (arg < -1.0 ? -1.0 : arg);
    t = m * std::cos(std::acos(arg) / 3.0);                             // largest of the three real roots
  }
  double x = t - a3;
  for (int i = 0; i < 3; ++i) {                                         // Newton polish
    double f = ((x + A) * x + B) * x + C, df = (3.0 * x + 2.0 * A) * x + B;
    if (df == 0.0) break;
    double xn = x - f / df;
    if (!std::isfinite(xn)) break;
    x = xn;
  }
  return x;
}

static int quadratic_real(double b, double c, double* r) {              // x^2 + b x + c
  double disc = b * b - 4.0 * c;
  if (!(disc >= 0.0)) return 0;
  double sq = std::sqrt(disc);
  double q = -0.5 * (b + (b >= 0 ? sq : -sq));
  r[0] = q;
  r[1] = (q != 0.0) ? c / q : 0.0;
  return 2;
}

static int solve_quartic_real(const double* c, double* roots) {
  if (c[4] == 0.0 || !std::isfinite(c[4])) return 0;
  double a = c[3] / c[4], b = c[2] / c[4], cc = c[1] / c[4], d = c[0] / c[4];
  if (!(std::isfinite(a) && std::isfinite(b) && std::isfinite(cc) && std::isfinite(d))) return 0;
  double a4 = a / 4.0;
  double p = b - 6.0 * a4 * a4;
  double q = cc - 2.0 * b * a4 + 8.0 * a4 * a4 * a4;
  double r = d - cc * a4 + b * a4 * a4 - 3.0 * a4 * a4 * a4 * a4;
  double y[4];
  int n = 0;
  // resolvent: m^3 + p m^2 + (p^2/4 - r) m - q^2/8 = 0, take the largest real root (>= 0)
  double m = cubic_largest_real_root(p, p * p / 4.0 - r, -q * q / 8.0);
  if (m > 0.0 && std::fabs(q) > 0.0) {
    double s = std::sqrt(2.0 * m);
    n += quadratic_real(-s, p / 2.0 + m + q / (2.0 * s), y + n);
    n += quadratic_real(s, p / 2.0 + m - q / (2.0 * s), y + n);
  } else {                                                              // biquadratic y^4 + p y^2 + r
    double z[2];
    int nz = quadratic_real(p, r, z);
    for (int i = 0; i < nz; ++i)
      if (z[i] >= 0.0) { double s = std::sqrt(z[i]); y[n++] = s; y[n++] = -s; }
  }
  for (int i = 0; i < n; ++i) {
    double x = y[i] - a4;
    for (int it = 0; it < 3; ++it) {
      double f = (((c[4] * x + c[3]) * x + c[2]) * x + c[1]) * x + c[0];
      double df = ((4.0 * c[4] * x + 3.0 * c[3]) * x + 2.0 * c[2]) * x + c[1];
      if (df == 0.0) break;
      double xn = x - f / df;
      if (!std::isfinite(xn)) break;
      x = xn;
    }
    roots[i] = x;
  }
  std::sort(roots, roots + n);
  return n;
}

// ---------------------------------------------------------------------------------------------------
// Kneip P3P (solver_p3p.h:239-371) + admissibility filter (perspective_n_point_estimator.h:110-131).
// pts: N x 7 rows (u_n, v_n, x, y, z, u, v).  models: up to 4 row-major 3x4 [R|t].
// ---------------------------------------------------------------------------------------------------
static int p3p(const double* pts, const int* idx, double* models) {
  double f[3][3], P[3][3];
  for (int i = 0; i < 3; ++i) {
    const double* row = pts + 7 * (size_t)idx[i];
    f[i][0] = row[0]; f[i][1] = row[1]; f[i][2] = 1.0;
    normalize3(f[i]);
    P[i][0] = row[2]; P[i][1] = row[3]; P[i][2] = row[4];
  }
  double e1[3], e2[3], cr[3];
  auto edges = [&]() {
    for (int k = 0; k < 3; ++k) { e1[k] = P[1][k] - P[0][k]; e2[k] = P[2][k] - P[0][k]; }
  };
  edges();
  cross3(e1, e2, cr);
  if (dot3(cr, cr) < 1e-6) return 0;                                   // collinear world points
  double T[3][3], f2c[3];
  auto camera_frame = [&]() {
    for (int k = 0; k < 3; ++k) T[0][k] = f[0][k];
    cross3(f[0], f[1], T[2]);
    normalize3(T[2]);
    cross3(T[2], T[0], T[1]);
    for (int k = 0; k < 3; ++k) f2c[k] = dot3(T[k], f[2]);
  };
  camera_frame();
  if (f2c[2] > 0) {
    for (int k = 0; k < 3; ++k) { std::swap(f[0][k], f[1][k]); std::swap(P[0][k], P[1][k]); }
    camera_frame();
    edges();
  }
  if (std::fabs(f2c[2]) < std::numeric_limits<double>::epsilon()) return 0;
  double Nw[3][3], P2w[3];
  double d12 = std::sqrt(dot3(e1, e1));
  for (int k = 0; k < 3; ++k) Nw[0][k] = e1[k] / d12;
  cross3(Nw[0], e2, Nw[2]);
  normalize3(Nw[2]);
  cross3(Nw[2], Nw[0], Nw[1]);
  for (int k = 0; k < 3; ++k) P2w[k] = dot3(Nw[k], e2);
  const double f1 = f2c[0] / f2c[2], f2 = f2c[1] / f2c[2], p1 = P2w[0], p2 = P2w[1];
  const double cos_beta = dot3(f[0], f[1]);
  double b = 1.0 / (1.0 - cos_beta * cos_beta) - 1.0;
  b = cos_beta < 0 ? -std::sqrt(b) : std::sqrt(b);
  const double F1 = f1 * f1, F2 = f2 * f2, P1 = p1 * p1, P2 = p2 * p2, D = d12, D2 = d12 * d12, b2 = b * b;
  double c[5];
  c[4] = -P2 * P2 * (F2 + F1 + 1.0);
  c[3] = 2.0 * p2 * P2 * D * (b * (1.0 + F2) - f1 * f2);
  c[2] = P2 * (-F2 * P1 - F2 * D2 * b2 - F2 * D2 + F2 * P2 + P2 * F1 + 2.0 * p1 * D + 2.0 * f1 * f2 * p1 * D * b -
               P1 * F1 + 2.0 * p1 * F2 * D - D2 * b2 - 2.0 * P1);
  c[1] = 2.0 * p2 * D * (P1 * b + f1 * f2 * P2 - F2 * P2 * b - p1 * D * b);
  c[0] = -2.0 * f2 * P2 * f1 * p1 * D * b + F2 * P2 * D2 + 2.0 * p1 * P1 * D - P1 * D2 + F2 * P2 * P1 - P1 * P1 -
         2.0 * F2 * P2 * p1 * D + P2 * F1 * P1 + F2 * P2 * D2 * b2;
  double roots[4];
  int nr = solve_quartic_real(c, roots);
  int nm = 0;
  for (int i = 0; i < nr; ++i) {
    double ct = roots[i] > 1.0 ? 1.0 : (roots[i] < -1.0 ? -1.0 : roots[i]);
    double cot_a = (-f1 * p1 / f2 - ct * p2 + D * b) / (-f1 * ct * p2 / f2 + p1 - D);
    double st = std::sqrt(1.0 - ct * ct);
    double sa = std::sqrt(1.0 / (cot_a * cot_a + 1.0));
    double ca = std::sqrt(1.0 - sa * sa);
    if (cot_a < 0) ca = -ca;
    double g = sa * b + ca;
    double cnu[3] = {D * ca * g, ct * D * sa * g, st * D * sa * g};
    double C[3];
    for (int k = 0; k < 3; ++k) C[k] = P[0][k] + Nw[0][k] * cnu[0] + Nw[1][k] * cnu[1] + Nw[2][k] * cnu[2];
    const double Q[3][3] = {{-ca, -sa * ct, -sa * st}, {sa, -ca * ct, -ca * st}, {0.0, -st, ct}};
    // R = T^T Q N
    double QN[3][3], R[9];
    for (int r_ = 0; r_ < 3; ++r_)
      for (int k = 0; k < 3; ++k) QN[r_][k] = Q[r_][0] * Nw[0][k] + Q[r_][1] * Nw[1][k] + Q[r_][2] * Nw[2][k];
    for (int r_ = 0; r_ < 3; ++r_)
      for (int k = 0; k < 3; ++k) R[r_ * 3 + k] = T[0][r_] * QN[0][k] + T[1][r_] * QN[1][k] + T[2][r_] * QN[2][k];
    double t[3];
    for (int r_ = 0; r_ < 3; ++r_) t[r_] = -(R[r_ * 3] * C[0] + R[r_ * 3 + 1] * C[1] + R[r_ * 3 + 2] * C[2]);
    bool finite = true;
    for (int k = 0; k < 9; ++k) finite &= std::isfinite(R[k]);
    for (int k = 0; k < 3; ++k) finite &= std::isfinite(t[k]);
    if (!finite) continue;
    if (t[2] < 0.0 || det3(R) < -0.95) continue;                        // perspective_n_point_estimator.h:121-127
    double* m = models + 12 * nm++;
    for (int r_ = 0; r_ < 3; ++r_) {
      m[r_ * 4 + 0] = R[r_ * 3]; m[r_ * 4 + 1] = R[r_ * 3 + 1]; m[r_ * 4 + 2] = R[r_ * 3 + 2]; m[r_ * 4 + 3] = t[r_];
    }
  }
  return nm;
}

// ---------------------------------------------------------------------------------------------------
// cv::solvePnP(SOLVEPNP_ITERATIVE) with K = I, no distortion.
// ---------------------------------------------------------------------------------------------------
static void project_residuals(int n, const double* X, const double* uv, const double* param, double* err,
                              double* J /* 2n x 6 or null */) {
  double R[9], dR[27];
  rodrigues_to_matrix(param, R, J ? dR : nullptr);
  const double* t = param + 3;
  for (int i = 0; i < n; ++i) {
    const double* p = X + 3 * i;
    double Y0 = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + t[0];
    double Y1 = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + t[1];
    double Y2 = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + t[2];
    double iz = Y2 != 0.0 ? 1.0 / Y2 : 1.0;                             // cvProjectPoints2: z = z ? 1./z : 1
    double x = Y0 * iz, y = Y1 * iz;
    err[2 * i] = x - uv[2 * i];
    err[2 * i + 1] = y - uv[2 * i + 1];
    if (J) {
      double* j0 = J + (size_t)(2 * i) * 6;
      double* j1 = j0 + 6;
      for (int k = 0; k < 3; ++k) {
        const double* d = dR + 9 * k;
        double dY0 = d[0] * p[0] + d[1] * p[1] + d[2] * p[2];
        double dY1 = d[3] * p[0] + d[4] * p[1] + d[5] * p[2];
        double dY2 = d[6] * p[0] + d[7] * p[1] + d[8] * p[2];
        j0[k] = iz * (dY0 - x * dY2);
        j1[k] = iz * (dY1 - y * dY2);
      }
      j0[3] = iz; j0[4] = 0.0; j0[5] = -x * iz;
      j1[3] = 0.0; j1[4] = iz; j1[5] = -y * iz;
    }
  }
}

static bool lm_step(const double* JtJ, const double* JtErr, const double* prev, int lambdaLg10, double* param) {
  double A[36], b[6], x[6];
  double lambda = std::exp(lambdaLg10 * std::log(10.0));
  std::memcpy(A, JtJ, sizeof(A));
  std::memcpy(b, JtErr, sizeof(b));
  for (int i = 0; i < 6; ++i) A[i * 6 + i] *= 1.0 + lambda;
  if (!solve_linear(6, A, b, x)) return false;
  for (int i = 0; i < 6; ++i) param[i] = prev[i] - x[i];
  return true;
}

// CvLevMarq(6, 2n, TermCriteria(EPS+ITER, 20, FLT_EPSILON), completeSymmFlag=true) driven as in
// cvFindExtrinsicCameraParams2.
static void lm_refine(int n, const double* X, const double* uv, double* param) {
  const int max_iter = 20;
  const double eps = (double)std::numeric_limits<float>::epsilon();
  std::vector<double> J((size_t)2 * n * 6), err((size_t)2 * n);
  int lambdaLg10 = -3, iters = 0;
  double prevErrNorm = std::numeric_limits<double>::max();
  for (;;) {
    project_residuals(n, X, uv, param, err.data(), J.data());           // state CALC_J
    double JtJ[36] = {0}, JtErr[6] = {0}, prev[6];
    for (int r = 0; r < 2 * n; ++r) {
      const double* j = &J[(size_t)r * 6];
      for (int a = 0; a < 6; ++a) {
        JtErr[a] += j[a] * err[r];
        for (int b_ = 0; b_ < 6; ++b_) JtJ[a * 6 + b_] += j[a] * j[b_];
      }
    }
    std::memcpy(prev, param, sizeof(prev));
    if (!lm_step(JtJ, JtErr, prev, lambdaLg10, param)) { std::memcpy(param, prev, sizeof(prev)); return; }
    if (iters == 0) {
      double s = 0;
      for (double e : err) s += e * e;
      prevErrNorm = std::sqrt(s);
    }
    double errNorm;
    for (;;) {                                                          // state CHECK_ERR
      project_residuals(n, X, uv, param, err.data(), nullptr);
      double s = 0;
      for (double e : err) s += e * e;
      errNorm = std::sqrt(s);
      if (errNorm > prevErrNorm && ++lambdaLg10 <= 16) {
        if (!lm_step(JtJ, JtErr, prev, lambdaLg10, param)) { std::memcpy(param, prev, sizeof(prev)); return; }
        continue;
      }
      break;
    }
    lambdaLg10 = std::max(lambdaLg10 - 1, -16);
    double dn = 0, pn = 0;
    for (int i = 0; i < 6; ++i) { dn += (param[i] - prev[i]) * (param[i] - prev[i]); pn += prev[i] * prev[i]; }
    double rel = std::sqrt(dn) / (pn > 0 ? std::sqrt(pn) : std::numeric_limits<double>::min());
    if (++iters >= max_iter || rel < eps) return;
    prevErrNorm = errNorm;
  }
}

// normalised-DLT homography: dst ~ H src (n >= 4)
static bool homography_dlt(int n, const double* src, const double* dst, double* H) {
  double cs[2] = {0, 0}, cd[2] = {0, 0};
  for (int i = 0; i < n; ++i) { cs[0] += src[2 * i]; cs[1] += src[2 * i + 1]; cd[0] += dst[2 * i]; cd[1] += dst[2 * i + 1]; }
  cs[0] /= n; cs[1] /= n; cd[0] /= n; cd[1] /= n;
  double ss = 0, sd = 0;
  for (int i = 0; i < n; ++i) {
    ss += std::sqrt((src[2 * i] - cs[0]) * (src[2 * i] - cs[0]) + (src[2 * i + 1] - cs[1]) * (src[2 * i + 1] - cs[1]));
    sd += std::sqrt((dst[2 * i] - cd[0]) * (dst[2 * i] - cd[0]) + (dst[2 * i + 1] - cd[1]) * (dst[2 * i + 1] - cd[1]));
  }
  if (ss == 0 || sd == 0) return false;
  ss = std::sqrt(2.0) * n / ss; sd = std::sqrt(2.0) * n / sd;
  double A[81] = {0};
  for (int i = 0; i < n; ++i) {
    double x = (src[2 * i] - cs[0]) * ss, y = (src[2 * i + 1] - cs[1]) * ss;
    double X = (dst[2 * i] - cd[0]) * sd, Y = (dst[2 * i + 1] - cd[1]) * sd;
    double r1[9] = {x, y, 1, 0, 0, 0, -X * x, -X * y, -X};
    double r2[9] = {0, 0, 0, x, y, 1, -Y * x, -Y * y, -Y};
    for (int a = 0; a < 9; ++a)
      for (int b = 0; b < 9; ++b) A[a * 9 + b] += r1[a] * r1[b] + r2[a] * r2[b];
  }
  double V[81], w[9];
  jacobi_eig(9, A, V, w);
  int k = 0;
  for (int i = 1; i < 9; ++i)
    if (w[i] < w[k]) k = i;
  double Hn[9];
  for (int i = 0; i < 9; ++i) Hn[i] = V[i * 9 + k];
  // H = Td^-1 Hn Ts
  double Ts[9] = {ss, 0, -ss * cs[0], 0, ss, -ss * cs[1], 0, 0, 1};
  double Tdi[9] = {1 / sd, 0, cd[0], 0, 1 / sd, cd[1], 0, 0, 1};
  double tmp[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) tmp[i * 3 + j] = Hn[i * 3] * Ts[j] + Hn[i * 3 + 1] * Ts[3 + j] + Hn[i * 3 + 2] * Ts[6 + j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) H[i * 3 + j] = Tdi[i * 3] * tmp[j] + Tdi[i * 3 + 1] * tmp[3 + j] + Tdi[i * 3 + 2] * tmp[6 + j];
  if (H[8] != 0.0)
    for (int i = 0; i < 9; ++i) H[i] /= H[8];
  for (int i = 0; i < 9; ++i)
    if (!std::isfinite(H[i])) return false;
  return true;
}

static bool pnp_init(int n, const double* X, const double* uv, double* param) {
  double Mc[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) Mc[k] += X[3 * i + k];
  for (int k = 0; k < 3; ++k) Mc[k] /= n;
  double MM[9] = {0};
  for (int i = 0; i < n; ++i)
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) MM[a * 3 + b] += (X[3 * i + a] - Mc[a]) * (X[3 * i + b] - Mc[b]);
  double V[9], w[3], MMc[9];
  std::memcpy(MMc, MM, sizeof(MM));
  jacobi_eig(3, MMc, V, w);
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int a, int b) { return w[a] > w[b]; });
  double W[3] = {w[order[0]], w[order[1]], w[order[2]]};
  double R[9], t[3];
  if (W[2] / W[1] < 1e-3) {
    // planar structure: rows of Rt = eigenvectors by decreasing eigenvalue (V^T of the SVD)
    double Rt[9];
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) Rt[r * 3 + k] = V[k * 3 + order[r]];
    if (Rt[2] * Rt[2] + Rt[5] * Rt[5] < 1e-10) { for (int i = 0; i < 9; ++i) Rt[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    if (det3(Rt) < 0) for (int i = 0; i < 9; ++i) Rt[i] = -Rt[i];
    double Tt[3];
    for (int r = 0; r < 3; ++r) Tt[r] = -(Rt[r * 3] * Mc[0] + Rt[r * 3 + 1] * Mc[1] + Rt[r * 3 + 2] * Mc[2]);
    std::vector<double> Mxy((size_t)2 * n);
    for (int i = 0; i < n; ++i) {
      const double* s = X + 3 * i;
      Mxy[2 * i] = Rt[0] * s[0] + Rt[1] * s[1] + Rt[2] * s[2] + Tt[0];
      Mxy[2 * i + 1] = Rt[3] * s[0] + Rt[4] * s[1] + Rt[5] * s[2] + Tt[1];
    }
    double H[9];
    if (n >= 4 && homography_dlt(n, Mxy.data(), uv, H)) {
      double h1[3] = {H[0], H[3], H[6]}, h2[3] = {H[1], H[4], H[7]}, h3[3] = {H[2], H[5], H[8]};
      double n1 = std::sqrt(dot3(h1, h1)), n2 = std::sqrt(dot3(h2, h2));
      const double de = std::numeric_limits<double>::epsilon();
      for (int k = 0; k < 3; ++k) { h1[k] /= std::max(n1, de); h2[k] /= std::max(n2, de); }
      double tt[3];
      for (int k = 0; k < 3; ++k) tt[k] = h3[k] * (2.0 / std::max(n1 + n2, de));
      cross3(h1, h2, h3);
      double Hm[9] = {h1[0], h2[0], h3[0], h1[1], h2[1], h3[1], h1[2], h2[2], h3[2]};
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
    if (n < 6) return false;                                            // DLT needs 6 points (deviation, see header)
    double LL[144] = {0};
    for (int i = 0; i < n; ++i) {
      const double* M = X + 3 * i;
      double x = -uv[2 * i], y = -uv[2 * i + 1];
      double r1[12] = {M[0], M[1], M[2], 1, 0, 0, 0, 0, x * M[0], x * M[1], x * M[2], x};
      double r2[12] = {0, 0, 0, 0, M[0], M[1], M[2], 1, y * M[0], y * M[1], y * M[2], y};
      for (int a = 0; a < 12; ++a)
        for (int b = 0; b < 12; ++b) LL[a * 12 + b] += r1[a] * r1[b] + r2[a] * r2[b];
    }
    double LV[144], LW[12];
    jacobi_eig(12, LL, LV, LW);
    int k = 0;
    for (int i = 1; i < 12; ++i)
      if (LW[i] < LW[k]) k = i;
    double RRt[12];
    for (int i = 0; i < 12; ++i) RRt[i] = LV[i * 12 + k];
    double RR[9] = {RRt[0], RRt[1], RRt[2], RRt[4], RRt[5], RRt[6], RRt[8], RRt[9], RRt[10]};
    double tt[3] = {RRt[3], RRt[7], RRt[11]};
    if (det3(RR) < 0) { for (int i = 0; i < 9; ++i) RR[i] = -RR[i]; for (int i = 0; i < 3; ++i) tt[i] = -tt[i]; }
    double sc = 0;
    for (int i = 0; i < 9; ++i) sc += RR[i] * RR[i];
    sc = std::sqrt(sc);
    if (!(sc > std::numeric_limits<double>::epsilon())) return false;
    if (!polar_rotation(RR, R)) return false;
    double nr = 0;
    for (int i = 0; i < 9; ++i) nr += R[i] * R[i];
    nr = std::sqrt(nr);
    for (int i = 0; i < 3; ++i) t[i] = tt[i] * (nr / sc);
  }
  matrix_to_rodrigues(R, param);
  param[3] = t[0]; param[4] = t[1]; param[5] = t[2];
  for (int i = 0; i < 6; ++i)
    if (!std::isfinite(param[i])) return false;
  return true;
}

static bool solvepnp_iterative(int n, const double* X, const double* uv, bool use_guess, double* param) {
  if (n < 4 && !(n == 3 && use_guess)) return false;
  if (!use_guess && !pnp_init(n, X, uv, param)) return false;
  lm_refine(n, X, uv, param);
  for (int i = 0; i < 6; ++i)
    if (!std::isfinite(param[i])) return false;
  return true;
}

// EPnPLM::estimateModel + the filter of estimateModelNonminimal (perspective_n_point_estimator.h:240-268).
static bool fit_nonminimal(const double* pts, const int* idx, int n, double* model) {
  if (n < 3) return false;
  std::vector<double> X((size_t)3 * n), uv((size_t)2 * n);
  for (int i = 0; i < n; ++i) {
    const double* row = pts + 7 * (size_t)idx[i];
    uv[2 * i] = row[0]; uv[2 * i + 1] = row[1];
    X[3 * i] = row[2]; X[3 * i + 1] = row[3]; X[3 * i + 2] = row[4];
  }
  double param[6];
  if (!solvepnp_iterative(n, X.data(), uv.data(), false, param)) return false;
  double R[9];
  rodrigues_to_matrix(param, R, nullptr);
  if (param[5] < 0.0 || det3(R) < -0.95) return false;
  for (int r = 0; r < 3; ++r) {
    model[r * 4] = R[r * 3]; model[r * 4 + 1] = R[r * 3 + 1]; model[r * 4 + 2] = R[r * 3 + 2]; model[r * 4 + 3] = param[3 + r];
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------
// residual + EPOS score
// ---------------------------------------------------------------------------------------------------
static inline double sq_residual(const double* row, const double* m) {
  // explicit fma chain so that the CUDA kernels (which restate the same chain) agree bit for bit
  double px = std::fma(m[0], row[2], std::fma(m[1], row[3], std::fma(m[2], row[4], m[3])));
  double py = std::fma(m[4], row[2], std::fma(m[5], row[3], std::fma(m[6], row[4], m[7])));
  double pz = std::fma(m[8], row[2], std::fma(m[9], row[3], std::fma(m[10], row[4], m[11])));
  double du = px / pz - row[0], dv = py / pz - row[1];
  return std::fma(du, du, dv * dv);
}

// Score::value is a double as in the reference (scoring_function.h:44-66): the pixel count in the single-instance
// path, pixel count minus (shared support)^2 when a compound model is present (Progressive-X).
struct Score { double value = 0; long long inliers = 0; };

struct Problem {
  // Progressive-X proposal engine only: preference vector of the compound model (NULL = plain EPOS score) and the
  // random-stream pair of this proposal (streams stream_base / stream_base + 1: main sampler / LO sampler)
  const double* cpref = nullptr;
  u64 stream_base = 0;
  int N = 0;
  std::vector<double> pts;          // N x 7
  std::vector<int> pixel_id;        // dense rank of ((int)u, (int)v)
  int used_pixels = 0;
  double thr_n = 0, sq_trunc = 0;   // normalised threshold, (1.5 thr_n)^2
  std::vector<std::vector<int>> nbr;
};

// getScore (scoring_function.h:220-267).  The early-out `N - i + inl < best.inl` can only fire at the last point
// examined, i.e. iff inl_total + 1 < best_inl (N - i + inl(i) is non-increasing in i); the result is then Score().
static Score get_score(const Problem& pb, const double* model, long long best_inl, std::vector<int>* inliers) {
  Score s;
  if (inliers) inliers->clear();
  std::vector<unsigned char> seen((size_t)pb.used_pixels, 0);
  long long pixels = 0;
  for (int i = 0; i < pb.N; ++i) {
    double r2 = sq_residual(&pb.pts[(size_t)7 * i], model);
    if (r2 < pb.sq_trunc) {
      if (inliers) inliers->push_back(i);
      if (!seen[pb.pixel_id[i]]) { seen[pb.pixel_id[i]] = 1; ++pixels; }
      ++s.inliers;
    }
  }
  if (s.inliers + 1 < best_inl) { if (inliers) inliers->clear(); return Score(); }
  s.value = (double)pixels;
  if (pb.cpref) {
    // EPOSScoringFunctionWithCompoundModel::getScore (scoring_function_with_compound_model.h:216-263): the support
    // shared with the compound model, sum_i min(compound_pref_i, pref_i) with pref_i = max(0, 1 - r_i^2 / T) for inliers,
    // squared (exponent_of_shared_score is an int set to 2, progressive_x.h:196,663), is subtracted from the pixel count.
    // Summation order (the reference adds sequentially; the CUDA warp cannot): point i goes to partial sum i mod 32 in
    // ascending order, the 32 partial sums are combined by a butterfly (offsets 16, 8, 4, 2, 1).
    double part[32] = {0};
    for (int i = 0; i < pb.N; ++i) {
      double r2 = sq_residual(&pb.pts[(size_t)7 * i], model);
      if (r2 < pb.sq_trunc) {
        double pref = 1.0 - r2 / pb.sq_trunc;
        if (pref < 0.0) pref = 0.0;
        part[i & 31] += std::min(pb.cpref[i], pref);
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      double t[32];
      for (int l = 0; l < 32; ++l) t[l] = part[l] + part[l ^ o];
      std::memcpy(part, t, sizeof(t));
    }
    s.value -= part[0] * part[0];
  }
  return s;
}

// ---------------------------------------------------------------------------------------------------
// neighbourhood graph (deterministic stand-in for FLANN, see header)
// ---------------------------------------------------------------------------------------------------
static void build_neighbors(Problem& pb, double radius, double scaling, int max_nbr) {
  const int N = pb.N;
  std::vector<float> q((size_t)5 * N);
  for (int i = 0; i < N; ++i) {
    const double* r = &pb.pts[(size_t)7 * i];
    q[5 * i] = (float)r[5]; q[5 * i + 1] = (float)r[6];
    q[5 * i + 2] = (float)(r[2] * scaling); q[5 * i + 3] = (float)(r[3] * scaling); q[5 * i + 4] = (float)(r[4] * scaling);
  }
  const float r2 = (float)radius * (float)radius;
  pb.nbr.assign(N, std::vector<int>());
  if (N == 0) return;
  // Uniform grid over (u, v) with cells at least one radius wide: a point within the 5-D radius is within the radius
  // in (u, v), hence in the 3 x 3 block of cells around the query.  The candidate SET is the one an all-pairs scan
  // finds and (distance, index) pairs are ordered the same way, so the lists are identical to the O(N^2) search
  // (the reference builds a KD-tree, flann_neighborhood_graph.h:86-125: sub-quadratic as well).
  float mnu = q[0], mxu = q[0], mnv = q[1], mxv = q[1];
  bool finite = true;
  for (int i = 0; i < N; ++i) {
    mnu = std::min(mnu, q[5 * i]); mxu = std::max(mxu, q[5 * i]);
    mnv = std::min(mnv, q[5 * i + 1]); mxv = std::max(mxv, q[5 * i + 1]);
    finite = finite && std::isfinite(q[5 * i]) && std::isfinite(q[5 * i + 1]);
  }
  float cs = (float)radius * 1.0001f + 1e-6f;
  long long gw = 1, gh = 1;
  if (finite) {
    for (;;) {
      gw = (long long)std::floor((mxu - mnu) / cs) + 1;
      gh = (long long)std::floor((mxv - mnv) / cs) + 1;
      if (gw * gh <= (1 << 22)) break;
      cs *= 2.0f;
    }
  }
  auto cell = [&](int i, long long* cx, long long* cy) {
    if (!finite) { *cx = *cy = 0; return; }
    long long x = (long long)std::floor((q[5 * i] - mnu) / cs), y = (long long)std::floor((q[5 * i + 1] - mnv) / cs);
    *cx = std::min(std::max(x, 0LL), gw - 1); *cy = std::min(std::max(y, 0LL), gh - 1);
  };
  std::vector<std::vector<int>> cells((size_t)(gw * gh));
  for (int i = 0; i < N; ++i) { long long cx, cy; cell(i, &cx, &cy); cells[(size_t)(cy * gw + cx)].push_back(i); }
  std::vector<std::pair<float, int>> cand;
  for (int i = 0; i < N; ++i) {
    cand.clear();
    long long cx, cy;
    cell(i, &cx, &cy);
    for (long long yy = std::max(cy - 1, 0LL); yy <= std::min(cy + 1, gh - 1); ++yy)
      for (long long xx = std::max(cx - 1, 0LL); xx <= std::min(cx + 1, gw - 1); ++xx)
        for (int j : cells[(size_t)(yy * gw + xx)]) {
          if (j == i) continue;
          float d = 0.f;
          for (int k = 0; k < 5; ++k) { float e = q[5 * i + k] - q[5 * j + k]; d = std::fmaf(e, e, d); }
          if (d <= r2) cand.emplace_back(d, j);
        }
    size_t keep = std::min(cand.size(), (size_t)max_nbr);
    std::partial_sort(cand.begin(), cand.begin() + keep, cand.end());
    for (size_t k = 0; k < keep; ++k) pb.nbr[i].push_back(cand[k].second);
  }
}

// All-pairs form of the same search (tests/test_oracle_pose.py checks the two agree).
static void build_neighbors_bruteforce(Problem& pb, double radius, double scaling, int max_nbr) {
  const int N = pb.N;
  std::vector<float> q((size_t)5 * N);
  for (int i = 0; i < N; ++i) {
    const double* r = &pb.pts[(size_t)7 * i];
    q[5 * i] = (float)r[5]; q[5 * i + 1] = (float)r[6];
    q[5 * i + 2] = (float)(r[2] * scaling); q[5 * i + 3] = (float)(r[3] * scaling); q[5 * i + 4] = (float)(r[4] * scaling);
  }
  const float r2 = (float)radius * (float)radius;
  pb.nbr.assign(N, std::vector<int>());
  std::vector<std::pair<float, int>> cand;
  for (int i = 0; i < N; ++i) {
    cand.clear();
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      float d = 0.f;
      for (int k = 0; k < 5; ++k) { float e = q[5 * i + k] - q[5 * j + k]; d = std::fmaf(e, e, d); }
      if (d <= r2) cand.emplace_back(d, j);
    }
    size_t keep = std::min(cand.size(), (size_t)max_nbr);
    std::partial_sort(cand.begin(), cand.begin() + keep, cand.end());
    for (size_t k = 0; k < keep; ++k) pb.nbr[i].push_back(cand[k].second);
  }
}

// ---------------------------------------------------------------------------------------------------
// graph-cut labeling (GCRANSAC.h:812-920, energy.h:204-253): returns the SINK segment (inliers)
// ---------------------------------------------------------------------------------------------------
constexpr double CUT_EPS = 1e-9;     // capacities are O(lambda) = O(0.1)

struct Dinic {
  struct Arc { int to; double cap; };
  std::vector<Arc> arcs;
  std::vector<std::vector<int>> adj;
  std::vector<int> level, it;
  explicit Dinic(int n) : adj(n), level(n), it(n) {}
  void add(int u, int v, double c, double rc) {
    adj[u].push_back((int)arcs.size()); arcs.push_back({v, c});
    adj[v].push_back((int)arcs.size()); arcs.push_back({u, rc});
  }
  bool bfs(int s, int t) {
    std::fill(level.begin(), level.end(), -1);
    std::queue<int> q;
    level[s] = 0; q.push(s);
    while (!q.empty()) {
      int u = q.front(); q.pop();
      for (int a : adj[u])
        if (arcs[a].cap > 0 && level[arcs[a].to] < 0) { level[arcs[a].to] = level[u] + 1; q.push(arcs[a].to); }
    }
    return level[t] >= 0;
  }
  double dfs(int u, int t, double f) {
    if (u == t) return f;
    for (int& i = it[u]; i < (int)adj[u].size(); ++i) {
      int a = adj[u][i];
      int v = arcs[a].to;
      if (arcs[a].cap > 0 && level[v] == level[u] + 1) {
        double d = dfs(v, t, std::min(f, arcs[a].cap));
        if (d > 0) {
          if (d >= arcs[a].cap) arcs[a].cap = 0.0; else arcs[a].cap -= d;   // bottleneck arc ends exactly at 0
          arcs[a ^ 1].cap += d;
          return d;
        }
      }
    }
    return 0.0;
  }
  void run(int s, int t) {
    while (bfs(s, t)) {
      std::fill(it.begin(), it.end(), 0);
      while (dfs(s, t, std::numeric_limits<double>::infinity()) > 0) {}
    }
  }
};

// Builds terminal capacities tr[i] (source minus sink) and the n-link list in the reference's order.
struct CutGraph {
  std::vector<double> tr;
  std::vector<int> ex, ey;
  std::vector<double> cxy, cyx;
  std::vector<double> u0, u1, e00;   // raw energy terms as passed to add_term1 / add_term2 (for the BK cross-check)
};

static void build_cut_graph(const Problem& pb, const double* model, double lambda, CutGraph& g) {
  const int N = pb.N;
  const double T = pb.sq_trunc, oml = 1.0 - lambda;
  std::vector<double> d(N);
  g.tr.assign(N, 0.0);
  g.ex.clear(); g.ey.clear(); g.cxy.clear(); g.cyx.clear();
  g.u0.assign(N, 0.0); g.u1.assign(N, 0.0); g.e00.clear();
  for (int i = 0; i < N; ++i) {
    double r2 = sq_residual(&pb.pts[(size_t)7 * i], model);
    double q = r2 / T;
    d[i] = q < 0.0 ? 0.0 : (q > 1.0 ? 1.0 : q);
    if (!(q == q)) d[i] = 0.0 < q ? 1.0 : 0.0;                          // NaN residual: clamp() is UB; treat as outlier
    double e = 1.0 - d[i];
    if (r2 <= T) { g.u0[i] = oml * e; g.u1[i] = 0.0; }                  // add_term1(i, oml*e, 0)
    else { g.u0[i] = 0.0; g.u1[i] = oml * (1.0 - e); }                  // add_term1(i, 0, oml*(1-e))
    g.tr[i] += g.u1[i] - g.u0[i];                                       // add_tweights(i, source = E1, sink = E0)
  }
  if (lambda > 0) {
    std::set<std::pair<int, int>> used;
    for (int i = 0; i < N; ++i)
      for (int j : pb.nbr[i]) {
        if (j == i || j < 0) continue;
        std::pair<int, int> key(std::min(i, j), std::max(i, j));
        if (!used.insert(key).second) continue;
        double e00 = 0.5 * (d[i] + d[j]);
        double A = e00 * lambda, B = lambda, C = lambda, D = 0.0 * lambda;
        g.tr[i] += D - A;                                               // add_tweights(x, D, A)
        B -= A; C -= D;
        g.ex.push_back(i); g.ey.push_back(j); g.cxy.push_back(B); g.cyx.push_back(C);   // B, C >= 0 always here
        g.e00.push_back(A);
      }
  }
}

static void labeling(const Problem& pb, const double* model, double lambda, std::vector<int>& inliers) {
  const int N = pb.N;
  CutGraph g;
  build_cut_graph(pb, model, lambda, g);
  Dinic dn(N + 2);
  const int S = N, Tn = N + 1;
  for (int i = 0; i < N; ++i) {
    if (g.tr[i] > 0) dn.add(S, i, g.tr[i], 0.0);
    else if (g.tr[i] < 0) dn.add(i, Tn, -g.tr[i], 0.0);
  }
  for (size_t e = 0; e < g.ex.size(); ++e) dn.add(g.ex[e], g.ey[e], g.cxy[e], g.cyx[e]);
  dn.run(S, Tn);
  // SINK segment = nodes that can still reach the sink through residual arcs.  A residual capacity below CUT_EPS
  // counts as saturated: when an outlier's terminal capacity equals the total capacity of its arcs in exact arithmetic
  // (e.g. exactly (1-lambda)/lambda incident edges towards inliers) both labelings have the same energy and the
  // answer of ANY floating-point max-flow (the reference's BK included) hinges on the rounding of its own sequence of
  // subtractions; the tolerance makes the answer the exact-arithmetic one (saturated -> SOURCE) for every algorithm.
  std::vector<char> in_t(N + 2, 0);
  std::queue<int> q;
  in_t[Tn] = 1; q.push(Tn);
  while (!q.empty()) {
    int v = q.front(); q.pop();
    for (int a : dn.adj[v]) {
      int u = dn.arcs[a].to;                                            // arc a: v->u ; its pair a^1: u->v
      if (!in_t[u] && dn.arcs[a ^ 1].cap > CUT_EPS) { in_t[u] = 1; q.push(u); }
    }
  }
  inliers.clear();
  for (int i = 0; i < N; ++i)
    if (in_t[i]) inliers.push_back(i);
}

// ---------------------------------------------------------------------------------------------------
// GC-RANSAC
// ---------------------------------------------------------------------------------------------------
struct Stats { int iterations = 0, graph_cuts = 0, lo_runs = 0, passes = 0, found = 0; };

static size_t iteration_bound(double confidence, long long inl, int N) {
  if (confidence == 1.0) return std::numeric_limits<size_t>::max();
  double q = std::pow((double)inl / N, 3.0);
  double l2 = std::log(1 - q);
  if (std::fabs(l2) < std::numeric_limits<double>::epsilon()) return std::numeric_limits<size_t>::max();
  double it = std::log(1.0 - confidence) / l2;
  return (size_t)it + 1;
}

static bool valid_sample(const Problem& pb, const int* s, double min_area) {
  const double* a = &pb.pts[(size_t)7 * s[0]];
  const double* b = &pb.pts[(size_t)7 * s[1]];
  const double* c = &pb.pts[(size_t)7 * s[2]];
  double area = 0.5 * std::fabs((b[5] - a[5]) * (c[6] - a[6]) - (c[5] - a[5]) * (b[6] - a[6]));
  return area > min_area;
}

// One pass of the main loop's model generation (GCRANSAC.h:296-321) for pass index `pass`.
static int generate_models(const Problem& pb, const Params& P, u64 seed, int pass, double* models, int* fails_out,
                           int* sample_out) {
  int fails = -1, nm = 0;
  while (++fails < P.max_unsuccessful) {
    int s[3];
    if (!unique_set(seed, pb.stream_base, (u64)pass, (u64)fails, pb.N, 3, s)) continue;
    if (!valid_sample(pb, s, P.min_triangle_area)) continue;
    nm = p3p(pb.pts.data(), s, models);
    if (nm > 0) { if (sample_out) { sample_out[0] = s[0]; sample_out[1] = s[1]; sample_out[2] = s[2]; } break; }
  }
  *fails_out = fails;
  return nm;
}

// debugging aid: rows of 72 ints per LO round, same layout as the CUDA library's epos_fit_debug_trace
static std::vector<int> g_trace;

static void local_optimization(const Problem& pb, const Params& P, u64 seed, Stats& st, double* best_model,
                               Score& best_score) {
  Score max_score = best_score;
  double lo_model[12];
  std::memcpy(lo_model, best_model, sizeof(lo_model));
  std::vector<int> inliers, tmp;
  ++st.lo_runs;
  while (++st.graph_cuts < P.max_graph_cuts) {
    bool updated = false;
    labeling(pb, lo_model, P.spatial_coherence_weight, inliers);
    const int ni = (int)inliers.size();
    const int sample_size = std::min(21, ni);
    std::vector<int> row(72, 0);
    for (int t = 0; t < 20; ++t) row[5 + 3 * t] = -1;
    for (int trial = 0; trial < P.max_lo_trials; ++trial) {
      double model[12];
      if (trial < 20) row[5 + 3 * trial] = 0;
      if (sample_size < ni) {
        int sel[21], idx[21];
        unique_set(seed, pb.stream_base + 1, (u64)st.graph_cuts, (u64)trial, ni, sample_size, sel);
        for (int k = 0; k < sample_size; ++k) idx[k] = inliers[sel[k]];
        if (!fit_nonminimal(pb.pts.data(), idx, sample_size, model)) continue;
      } else if (3 < ni) {
        if (!fit_nonminimal(pb.pts.data(), inliers.data(), ni, model)) break;
      } else {
        break;
      }
      Score s = get_score(pb, model, max_score.inliers, nullptr);
      if (trial < 20) {
        Score raw = get_score(pb, model, 0, nullptr);
        row[5 + 3 * trial] = 1; row[6 + 3 * trial] = (int)raw.inliers; row[7 + 3 * trial] = (int)raw.value;
      }
      if (max_score.value < s.value) {
        updated = true;
        max_score = s;
        std::memcpy(lo_model, model, sizeof(lo_model));
      }
    }
    row[0] = st.graph_cuts; row[1] = ni; row[2] = updated ? 1 : 0; row[3] = (int)max_score.value; row[4] = (int)max_score.inliers;
    if (g_trace.size() > 72 * 4096) g_trace.clear();
    g_trace.insert(g_trace.end(), row.begin(), row.end());
    if (!updated) break;
  }
  if (best_score.value < max_score.value) {
    best_score = max_score;
    std::memcpy(best_model, lo_model, sizeof(lo_model));
  }
}

// Returns 1 if a model was found.  model: row-major 3x4.  inliers: final inlier list.
static int gcransac_run(const Problem& pb, const Params& P, u64 seed, double* model_out, std::vector<int>& inliers,
                        Stats& st) {
  const int N = pb.N;
  Score best;
  double best_model[12] = {0};
  size_t max_iteration = iteration_bound(P.confidence, 1, N);
  double coverage = 0.0;
  size_t iter = 0;
  int pass = 0;
  while ((size_t)P.min_iters > iter || iter < std::min(max_iteration, (size_t)P.max_iters)) {
    if ((size_t)P.min_iters < iter) {
      if (iter > max_iteration) break;
      if (iter > (size_t)P.max_iters) break;
      if (P.min_coverage < coverage) break;
    }
    bool do_lo = false;
    ++iter;
    double models[48];
    int fails = 0;
    int nm = generate_models(pb, P, seed, pass, models, &fails, nullptr);
    iter += (size_t)fails;
    for (int m = 0; m < nm; ++m) {
      Score s = get_score(pb, models + 12 * m, best.inliers, nullptr);
      if (best.value < s.value) {
        best = s;
        std::memcpy(best_model, models + 12 * m, sizeof(best_model));
        do_lo = iter > (size_t)P.min_iters_before_lo && best.inliers > 3;
        max_iteration = iteration_bound(P.confidence, best.inliers, N);
        coverage = (double)best.value / (double)pb.used_pixels;
      }
    }
    if (do_lo) {
      ++st.lo_runs;
      local_optimization(pb, P, seed, st, best_model, best);
      max_iteration = iteration_bound(P.confidence, best.inliers, N);
      coverage = (double)best.value / (double)pb.used_pixels;
    }
    ++pass;
  }
  st.iterations = (int)iter;
  st.passes = pass;
  if (best.inliers <= 3) return 0;
  if (st.lo_runs == 0) {
    ++st.lo_runs;
    local_optimization(pb, P, seed, st, best_model, best);
  }
  // final inlier set = inliers of the best model (GCRANSAC.h:470-478)
  get_score(pb, best_model, 0, &inliers);
  // iterated least squares (GCRANSAC.h:480-508, 533-657)
  bool refit_applied = false;
  {
    double model[12];
    std::memcpy(model, best_model, sizeof(model));
    std::vector<int> inl = inliers, tmp;
    int iterations = 0;
    if ((int)inl.size() > 3) {
      while (++iterations < P.max_lsq_iters) {
        double m2[12];
        if (!fit_nonminimal(pb.pts.data(), inl.data(), (int)inl.size(), m2)) break;
        Score s = get_score(pb, m2, 0, &tmp);
        if ((int)tmp.size() < 3) break;
        if (s.inliers <= (long long)inl.size()) break;
        std::memcpy(model, m2, sizeof(model));
        inl.swap(tmp);
      }
      if (iterations > 1) {
        Score s = get_score(pb, model, 0, &tmp);
        if (best.value < s.value) {
          refit_applied = true;
          std::memcpy(best_model, model, sizeof(model));
          inliers.swap(tmp);
        }
      }
    }
  }
  if (!refit_applied) {
    double m2[12];
    if (fit_nonminimal(pb.pts.data(), inliers.data(), (int)inliers.size(), m2)) std::memcpy(best_model, m2, sizeof(m2));
  }
  std::memcpy(model_out, best_model, sizeof(best_model));
  st.found = 1;
  return 1;
}

static void make_problem(int N, const double* x2d, const double* x3d, const double* K, const Params& P, Problem& pb) {
  pb.N = N;
  pb.pts.resize((size_t)7 * N);
  double Kinv[9];
  inv3(K, Kinv);
  std::map<std::pair<int, int>, int> pix;
  pb.pixel_id.resize(N);
  for (int i = 0; i < N; ++i) {
    double* r = &pb.pts[(size_t)7 * i];
    double u = x2d[2 * i], v = x2d[2 * i + 1];
    r[5] = u; r[6] = v;
    r[2] = x3d[3 * i]; r[3] = x3d[3 * i + 1]; r[4] = x3d[3 * i + 2];
    r[0] = Kinv[0] * u + Kinv[1] * v + Kinv[2];
    r[1] = Kinv[3] * u + Kinv[4] * v + Kinv[5];
    auto key = std::make_pair((int)u, (int)v);
    auto it = pix.find(key);
    if (it == pix.end()) it = pix.emplace(key, (int)pix.size()).first;
    pb.pixel_id[i] = it->second;
  }
  pb.used_pixels = (int)pix.size();
  pb.thr_n = P.threshold / (0.5 * (K[0] + K[4]));
  double tt = 1.5 * pb.thr_n;
  pb.sq_trunc = tt * tt;
}

static int find6dposes(int N, const double* x2d, const double* x3d, const double* K, const Params& P, u64 seed,
                       const int* nbr_offsets, const int* nbr_index, double* pose, int* labeling_out, Stats& st) {
  Problem pb;
  make_problem(N, x2d, x3d, K, P, pb);
  if (nbr_offsets) {
    pb.nbr.assign(N, std::vector<int>());
    for (int i = 0; i < N; ++i) pb.nbr[i].assign(nbr_index + nbr_offsets[i], nbr_index + nbr_offsets[i + 1]);
  } else {
    build_neighbors(pb, P.neighborhood_ball_radius, P.scaling_from_millimeters, P.max_neighbors);
  }
  std::vector<int> inliers;
  double model[12];
  for (int i = 0; i < N; ++i) labeling_out[i] = 0;
  if (!gcransac_run(pb, P, seed, model, inliers, st)) return 0;
  // final LM refinement (progressivex_python.cpp:257-312)
  if (P.apply_numerical_optimization && inliers.size() >= 6) {
    const int n = (int)inliers.size();
    std::vector<double> X((size_t)3 * n), uv((size_t)2 * n);
    for (int i = 0; i < n; ++i) {
      const double* r = &pb.pts[(size_t)7 * inliers[i]];
      uv[2 * i] = r[0]; uv[2 * i + 1] = r[1];
      X[3 * i] = r[2]; X[3 * i + 1] = r[3]; X[3 * i + 2] = r[4];
    }
    double R[9] = {model[0], model[1], model[2], model[4], model[5], model[6], model[8], model[9], model[10]};
    double param[6];
    matrix_to_rodrigues(R, param);
    param[3] = model[3]; param[4] = model[7]; param[5] = model[11];
    if (solvepnp_iterative(n, X.data(), uv.data(), true, param)) {
      rodrigues_to_matrix(param, R, nullptr);
      for (int r = 0; r < 3; ++r) {
        model[r * 4] = R[r * 3]; model[r * 4 + 1] = R[r * 3 + 1]; model[r * 4 + 2] = R[r * 3 + 2];
        model[r * 4 + 3] = param[3 + r];
      }
    }
  }
  for (int i : inliers) labeling_out[i] = 1;
  std::memcpy(pose, model, sizeof(model));
  return 1;
}


// ===================================================================================================
// Progressive-X multi-instance fitting (max_model_number != 1):
//   /root/reference/external/progressive-x/src/pyprogressivex/src/progressivex_python.cpp:136-221 (dispatch, settings)
//   .../include/progressive_x.h:397-649 (run), :265-391 (spedUpFitting), :651-673 (getPredictedUnseenInliers),
//       :675-721 (initialize), :723-749 (isPutativeModelValid), :755-794 (updateCompoundModel)
//   .../include/scoring_function_with_compound_model.h:127-266 (restated inside get_score above)
//   .../include/PEARL.h:57-134 (energy functors), :391-536 (run / labeling), :313-389 (parameterEstimation),
//       :271-311 (rejectInstances); progx_model.h:66-84 (preference vector)
//   alpha-expansion with label costs: .../graph-cut-ransac/src/pygcransac/include/GCoptimization.cpp:1003-1088
//       (expansion, standard cycles), :1239-1303 (alpha_expansion), :316-404 (active sites, data / smooth terms),
//       :1131-1196 (label costs), :452-470 (applyNewLabeling), :259-279,950-986 (energies); energy.h:204-253,324-328;
//       graph.h:388-397 (add_tweights).  Validated against the reference's own GCoptimization sources compiled into
//       oracle/_ref (tests/test_oracle_progx.py).
// ===================================================================================================
struct MultiParams {
  int max_model_number;            // > 1, or -1 = "all instances" (DETECTION)
  int max_model_number_for_pearl;  // maximum_model_number_to_optimize (EPOS: 5)
  double confidence;               // conf (required_progx_confidence, 0.5)
  double max_tanimoto;             // 0.9
  int min_point_number;            // 6: minimum inliers of an instance AND the label cost of PEARL
};
constexpr int MAX_INSTANCES = 32;  // cap of the "all instances" mode (the reference's loop is unbounded there)

// The binary energy of one expansion move (energy.h / graph.h), solved with Dinic; var(i) = 1 iff node i belongs to the
// SINK segment = can still reach the sink in the residual graph (what_segment with default SOURCE).
struct MoveEnergy {
  std::vector<double> tr;
  double flow_const = 0.0;
  struct E { int x, y; double cap, rev; };
  std::vector<E> edges;
  int add_variable() { tr.push_back(0.0); return (int)tr.size() - 1; }
  void add_tweights(int i, double cap_source, double cap_sink) {          // graph.h:388-397
    double delta = tr[i];
    if (delta > 0) cap_source += delta; else cap_sink -= delta;
    flow_const += (cap_source < cap_sink) ? cap_source : cap_sink;
    tr[i] = cap_source - cap_sink;
  }
  void add_term1(int x, double A, double B) { add_tweights(x, B, A); }    // energy.h:204-209
  void add_term2(int x, int y, double A, double B, double C, double D) {  // energy.h:211-253
    add_tweights(x, D, A);
    B -= A; C -= D;
    if (B < 0) { add_tweights(x, 0, B); add_tweights(y, 0, -B); edges.push_back({x, y, 0.0, B + C}); }
    else if (C < 0) { add_tweights(x, 0, -C); add_tweights(y, 0, C); edges.push_back({x, y, B + C, 0.0}); }
    else edges.push_back({x, y, B, C});
  }
  double minimize(std::vector<char>& var) {
    const int n = (int)tr.size();
    Dinic dn(n + 2);
    const int S = n, T = n + 1;
    for (int i = 0; i < n; ++i) {
      if (tr[i] > 0) dn.add(S, i, tr[i], 0.0);
      else if (tr[i] < 0) dn.add(i, T, -tr[i], 0.0);
    }
    for (const E& e : edges) dn.add(e.x, e.y, e.cap, e.rev);
    // flow value = capacity that left the source
    double before = 0.0, after = 0.0;
    for (int a : dn.adj[S]) before += dn.arcs[a].cap;
    dn.run(S, T);
    for (int a : dn.adj[S]) after += dn.arcs[a].cap;
    var.assign(n, 0);
    std::vector<char> in_t(n + 2, 0);
    std::queue<int> q;
    in_t[T] = 1; q.push(T);
    while (!q.empty()) {
      int v = q.front(); q.pop();
      for (int a : dn.adj[v]) {
        int u = dn.arcs[a].to;
        if (!in_t[u] && dn.arcs[a ^ 1].cap > CUT_EPS) { in_t[u] = 1; q.push(u); }
      }
    }
    for (int i = 0; i < n; ++i) var[i] = in_t[i];
    return flow_const + (before - after);
  }
};

struct AlphaExpansion {
  int n = 0, L = 0;
  std::vector<int> label;
  std::vector<std::vector<std::pair<int, double>>> nb;   // neighbour arrays in the order finalizeNeighbors() produces
  std::vector<double> D;                                  // data costs [n][L]
  double lambda = 0.0, label_cost = 0.0;
  std::vector<int> counts;
  std::vector<double> lab_cost;
  int moves = 0;

  double sc(int a, int b) const { return a != b ? lambda : 0.0; }          // PEARL.h:57-79
  void update_info() {
    counts.assign(L, 0);
    lab_cost.resize(n);
    for (int i = 0; i < n; ++i) { ++counts[label[i]]; lab_cost[i] = D[(size_t)i * L + label[i]]; }
  }
  double compute_energy() const {                                           // GCoptimization.cpp:950-986,259-279
    double data = 0.0, smooth = 0.0, lab = 0.0;
    for (int i = 0; i < n; ++i) data += lab_cost[i];
    for (int i = 0; i < n; ++i)
      for (const auto& e : nb[i])
        if (e.first < i) smooth += e.second * sc(label[i], label[e.first]);
    for (int l = L - 1; l >= 0; --l)                                        // m_labelcostsAll is prepended: last label first
      if (counts[l]) lab += label_cost;
    return data + smooth + lab;
  }
  bool expansion_move(int alpha) {                                          // GCoptimization.cpp:1239-1303
    std::vector<int> active;
    for (int i = 0; i < n; ++i) if (label[i] != alpha) active.push_back(i);
    const int size = (int)active.size();
    if (size == 0) return false;
    ++moves;
    std::vector<int> lookup(n, -1);
    for (int i = 0; i < size; ++i) lookup[active[i]] = i;
    MoveEnergy e;
    e.tr.assign(size, 0.0);
    double before = 0.0;
    for (int i = 0; i < size; ++i) {                                        // :328-334
      const int site = active[i];
      before += lab_cost[site];
      e.add_term1(i, D[(size_t)site * L + alpha], lab_cost[site]);
    }
    for (int i = size - 1; i >= 0; --i) {                                   // :338-404
      const int site = active[i];
      for (const auto& en : nb[site]) {
        const int ns = en.first; const double w = en.second;
        if (lookup[ns] == -1) {
          const double e0 = sc(alpha, label[ns]), e1 = sc(label[site], label[ns]);
          before += e1 * w;
          e.add_term1(i, e0 * w, e1 * w);
        } else if (ns < site) {
          const double e00 = sc(alpha, alpha), e01 = sc(alpha, label[ns]), e10 = sc(label[site], alpha),
                       e11 = sc(label[site], label[ns]);
          before += e11 * w;
          e.add_term2(i, lookup[ns], e00 * w, e01 * w, e10 * w, e11 * w);
        }
      }
    }
    double alpha_correction = 0.0;                                          // :1131-1196
    if (label_cost > 0.0) {
      std::vector<int> aux(L, -1);
      aux[alpha] = -2;
      if (!counts[alpha]) alpha_correction += label_cost;
      for (int i = 0; i < size; ++i) {
        const int l = label[active[i]];
        if (aux[l] == -2) continue;
        if (aux[l] == -1) {
          aux[l] = e.add_variable();
          e.add_term1(aux[l], 0.0, label_cost);
          before += label_cost;
        }
        e.add_term2(i, aux[l], 0.0, 0.0, label_cost, 0.0);
      }
    }
    std::vector<char> var;
    const double after = e.minimize(var) + alpha_correction;
    // The reference applies the move iff after < before (flow value vs energy before the move).  When the move cannot
    // improve anything the two are equal in exact arithmetic and the comparison is decided by rounding; a tolerance
    // makes that case "no move" for every implementation (the CUDA kernel compares the energies of the two labelings).
    const bool improves = before - after > 1e-9;
    if (improves) {                                                         // :452-470
      for (int i = 0; i < size; ++i)
        if (var[i] == 0) {
          const int site = active[i];
          --counts[label[site]]; ++counts[alpha];
          label[site] = alpha;
          lab_cost[site] = D[(size_t)site * L + alpha];
        }
    }
    return improves;
  }
  double expansion(int max_iterations) {                                    // :1003-1088, standard cycles
    update_info();
    double new_energy = compute_energy(), old_energy;
    for (int cycle = 1; cycle <= max_iterations; ++cycle) {
      old_energy = new_energy;
      for (int a = 0; a < L; ++a) expansion_move(a);
      new_energy = compute_energy();
      if (new_energy == old_energy) break;
    }
    return new_energy;
  }
};

// Neighbour arrays as GCoptimizationGeneralGraph builds them from PEARL.h:517-520: setNeighbors(i, j) for every
// listing j of i (i ascending) PREPENDS an entry to both sites' lists (GCoptimization.cpp:1683-1708); a mutual pair is
// therefore entered twice and weighs twice.
static void gco_neighbors(const std::vector<std::vector<int>>& nbr, std::vector<std::vector<std::pair<int, double>>>& nb) {
  const int n = (int)nbr.size();
  nb.assign(n, {});
  for (int i = 0; i < n; ++i)
    for (int j : nbr[i])
      if (j != i && j >= 0) { nb[i].insert(nb[i].begin(), {j, 1.0}); nb[j].insert(nb[j].begin(), {i, 1.0}); }
}

struct Instance {
  double m[12];
  std::vector<double> pref;        // preference vector as of the moment the instance was accepted (progx_model.h:66-84);
                                   // the reference does NOT refresh it when PEARL refits the instance
};

struct Pearl {
  std::vector<int> labels;         // alpha_expansion_engine->whatLabel
  bool have_engine = false;
  std::vector<int> outliers;
  std::vector<std::vector<int>> per_instance;
  int iterations = 0, moves = 0;

  void run(const Problem& pb, const std::vector<std::vector<std::pair<int, double>>>& nb, std::vector<Instance>& models,
           double lambda, int min_inliers) {
    const int N = pb.N;
    const double T2 = 9.0 / 4.0 * pb.thr_n * pb.thr_n, oml = 1.0 - lambda;   // PEARL.h:49 (not (1.5 thr)^2)
    int iteration = 0;
    double energy = 0.0, previous_energy = -1.0;
    bool rejected = false, converged = false;
    while (!converged && iteration++ < 50) {
      const bool init_prev = iteration > 1 && !rejected;
      if (!models.empty()) {                                               // labeling(), PEARL.h:461-536
        AlphaExpansion ax;
        ax.n = N; ax.L = (int)models.size() + 1; ax.lambda = lambda; ax.label_cost = (double)min_inliers;
        ax.nb = nb;
        ax.D.resize((size_t)N * ax.L);
        for (int i = 0; i < N; ++i) {
          for (int l = 0; l + 1 < ax.L; ++l) {                              // dataEnergyFunctor, PEARL.h:81-134
            const double r2 = sq_residual(&pb.pts[(size_t)7 * i], models[l].m);
            ax.D[(size_t)i * ax.L + l] = r2 > T2 ? 2.0 * oml : oml * r2 / T2;
          }
          ax.D[(size_t)i * ax.L + ax.L - 1] = oml;
        }
        if (init_prev && have_engine) ax.label = labels; else ax.label.assign(N, 0);
        energy = ax.expansion(1000);
        labels = ax.label;
        have_engine = true;
        moves += ax.moves;
      }
      bool changed = false;
      rejected = false;
      if (have_engine) {                                                   // parameterEstimation, PEARL.h:313-389
        const int M = (int)models.size();
        per_instance.assign(M, {});
        outliers.clear();
        for (int i = 0; i < N; ++i) {
          if (labels[i] < M) per_instance[labels[i]].push_back(i); else outliers.push_back(i);
        }
        for (int k = 0; k < M; ++k) {
          const std::vector<int>& inl = per_instance[k];
          if ((int)inl.size() < 3) continue;
          double before = 0.0, after = 0.0;
          for (int i : inl) before += std::sqrt(sq_residual(&pb.pts[(size_t)7 * i], models[k].m));
          double m2[12];
          if (!fit_nonminimal(pb.pts.data(), inl.data(), (int)inl.size(), m2)) continue;
          for (int i : inl) after += std::sqrt(sq_residual(&pb.pts[(size_t)7 * i], m2));
          if (after < before) { std::memcpy(models[k].m, m2, sizeof(m2)); changed = true; }
        }
      }
      for (int k = (int)models.size() - 1; k >= 0; --k)                    // rejectInstances, PEARL.h:271-311
        if ((int)per_instance[k].size() < min_inliers) {
          outliers.insert(outliers.end(), per_instance[k].begin(), per_instance[k].end());
          per_instance.erase(per_instance.begin() + k);
          models.erase(models.begin() + k);
          rejected = true;
        }
      if (!rejected && !changed && std::fabs(energy - previous_energy) < 1e-5 && iteration > 1) converged = true;
      previous_energy = energy;
    }
    iterations += iteration;
  }
};

struct MultiStats { int proposals = 0, accepted = 0, ransac_iterations = 0, pearl_iterations = 0, moves = 0, sped_up = 0; };

static void preference_vector(const Problem& pb, const double* model, double T2, std::vector<double>& pref) {
  pref.resize(pb.N);
  for (int i = 0; i < pb.N; ++i) {
    const double v = 1.0 - sq_residual(&pb.pts[(size_t)7 * i], model) / T2;
    pref[i] = v > 0.0 ? v : 0.0;                                           // MAX(0, v): a NaN residual gives 0
  }
}

// 7-column neighbourhood of spedUpFitting (progressive_x.h:288-289: FlannNeighborhoodGraph on the 7-column data with a
// hard-coded radius of 20): deterministic stand-in = the max_nbr nearest rows within the radius, f32 L2 over
// (u_n, v_n, x, y, z, u, v), ties by index.
static void build_neighbors7(Problem& pb, double radius, int max_nbr) {
  const int N = pb.N;
  std::vector<float> q((size_t)7 * N);
  for (size_t k = 0; k < q.size(); ++k) q[k] = (float)pb.pts[k];
  const float r2 = (float)radius * (float)radius;
  pb.nbr.assign(N, std::vector<int>());
  std::vector<std::pair<float, int>> cand;
  // (u, v) are columns 5, 6: a uniform grid on them bounds the search exactly as in build_neighbors
  if (N == 0) return;
  float mnu = q[5], mxu = q[5], mnv = q[6], mxv = q[6];
  bool finite = true;
  for (int i = 0; i < N; ++i) {
    mnu = std::min(mnu, q[7 * i + 5]); mxu = std::max(mxu, q[7 * i + 5]);
    mnv = std::min(mnv, q[7 * i + 6]); mxv = std::max(mxv, q[7 * i + 6]);
    finite = finite && std::isfinite(q[7 * i + 5]) && std::isfinite(q[7 * i + 6]);
  }
  float cs = (float)radius * 1.0001f + 1e-6f;
  long long gw = 1, gh = 1;
  if (finite)
    for (;;) {
      gw = (long long)std::floor((mxu - mnu) / cs) + 1; gh = (long long)std::floor((mxv - mnv) / cs) + 1;
      if (gw * gh <= (1 << 22)) break;
      cs *= 2.0f;
    }
  auto cell = [&](int i, long long* cx, long long* cy) {
    if (!finite) { *cx = *cy = 0; return; }
    long long x = (long long)std::floor((q[7 * i + 5] - mnu) / cs), y = (long long)std::floor((q[7 * i + 6] - mnv) / cs);
    *cx = std::min(std::max(x, 0LL), gw - 1); *cy = std::min(std::max(y, 0LL), gh - 1);
  };
  std::vector<std::vector<int>> cells((size_t)(gw * gh));
  for (int i = 0; i < N; ++i) { long long cx, cy; cell(i, &cx, &cy); cells[(size_t)(cy * gw + cx)].push_back(i); }
  for (int i = 0; i < N; ++i) {
    cand.clear();
    long long cx, cy;
    cell(i, &cx, &cy);
    for (long long yy = std::max(cy - 1, 0LL); yy <= std::min(cy + 1, gh - 1); ++yy)
      for (long long xx = std::max(cx - 1, 0LL); xx <= std::min(cx + 1, gw - 1); ++xx)
        for (int j : cells[(size_t)(yy * gw + xx)]) {
          if (j == i) continue;
          float d = 0.f;
          for (int k = 0; k < 7; ++k) { float e = q[7 * i + k] - q[7 * j + k]; d = std::fmaf(e, e, d); }
          if (d <= r2) cand.emplace_back(d, j);
        }
    size_t keep = std::min(cand.size(), (size_t)max_nbr);
    std::partial_sort(cand.begin(), cand.begin() + keep, cand.end());
    for (size_t k = 0; k < keep; ++k) pb.nbr[i].push_back(cand[k].second);
  }
}

// spedUpFitting (progressive_x.h:265-391): sequential GC-RANSAC (plain EPOS score, confidence 1) with removal of the
// inliers of every accepted proposal; no PEARL; the labeling stays all zero (the reference never writes it).
static void sped_up_fitting(const Problem& all, const Params& P, const MultiParams& MP, u64 seed,
                            std::vector<Instance>& models, MultiStats& ms) {
  ms.sped_up = 1;
  std::vector<int> indices(all.N);
  for (int i = 0; i < all.N; ++i) indices[i] = i;
  const bool unbounded = MP.max_model_number < 0;
  const int limit = unbounded ? MAX_INSTANCES : MP.max_model_number;
  for (int it = 0; it < limit; ++it) {
    Problem cur;
    cur.N = (int)indices.size();
    cur.pts.resize((size_t)7 * cur.N);
    std::map<std::pair<int, int>, int> pix;
    cur.pixel_id.resize(cur.N);
    for (int r = 0; r < cur.N; ++r) {
      std::memcpy(&cur.pts[(size_t)7 * r], &all.pts[(size_t)7 * indices[r]], 7 * sizeof(double));
      auto key = std::make_pair((int)cur.pts[(size_t)7 * r + 5], (int)cur.pts[(size_t)7 * r + 6]);
      auto itp = pix.find(key);
      if (itp == pix.end()) itp = pix.emplace(key, (int)pix.size()).first;
      cur.pixel_id[r] = itp->second;
    }
    cur.used_pixels = all.used_pixels;        // settings.proposal_engine_settings.used_pixels is never refreshed
    cur.thr_n = all.thr_n; cur.sq_trunc = all.sq_trunc;
    cur.stream_base = 2 * (u64)it;
    build_neighbors7(cur, 20.0, P.max_neighbors);
    Params Pl = P;
    Pl.confidence = 1.0;
    Stats st;
    std::vector<int> inliers;
    Instance inst;
    ++ms.proposals;
    const int found = gcransac_run(cur, Pl, seed, inst.m, inliers, st);
    if (!found) { if (unbounded) break; continue; }
    ms.ransac_iterations += st.iterations;
    models.push_back(inst);
    ++ms.accepted;
    if ((int)inliers.size() < MP.min_point_number) { if (unbounded) break; continue; }
    std::vector<char> mask(cur.N, 0);
    for (int i : inliers) mask[i] = 1;
    std::vector<int> rest;
    for (int r = 0; r < cur.N; ++r) if (!mask[r]) rest.push_back(indices[r]);
    indices.swap(rest);
    if (unbounded) {
      // "all instances" (max_model_number = -1): the reference's loop bound is (size_t)-1, i.e. it never terminates
      // (progressive_x.h:280 compares a size_t counter with an int holding -1).  Defined behaviour here: stop when no
      // model is found, when a proposal has fewer than min_point_number inliers, or when the termination test of
      // ProgressiveX::run (:589-611: predicted number of unseen inliers below min_point_number) fires on the points
      // that remain; at most MAX_INSTANCES instances.
      const double ratio = std::pow(1.0 - std::pow(1.0 - MP.confidence, 1.0 / (double)ms.ransac_iterations), 1.0 / 3.0);
      if (std::llround((double)indices.size() * ratio) < MP.min_point_number) break;
    }
  }
}

// ProgressiveX::run (progressive_x.h:397-649).  Returns the instances; labeling[i] = instance of point i (outlier label
// = number of instances; with a single instance 0 = inlier, 1 = outlier), scores[k] = sum of instance k's preferences.
static void progx_run(Problem& pb, const Params& P, const MultiParams& MP, u64 seed, std::vector<Instance>& models,
                      std::vector<int>& labeling, std::vector<double>& scores, MultiStats& ms) {
  const int N = pb.N;
  labeling.assign(N, 0);
  scores.clear();
  models.clear();
  if (MP.max_model_number < 0 || MP.max_model_number_for_pearl < MP.max_model_number) {   // :417-425 (size_t compare)
    sped_up_fitting(pb, P, MP, seed, models, ms);
    scores.assign(models.size(), 0.0);                                      // RANSACStatistics::score is never written
    return;
  }
  const double T2 = 9.0 / 4.0 * pb.thr_n * pb.thr_n;                        // :679
  std::vector<double> compound(N, 0.0);
  std::vector<std::vector<std::pair<int, double>>> nb;
  gco_neighbors(pb.nbr, nb);
  Pearl pearl;
  Params Pl = P;
  Pl.confidence = 1.0;                                                     // settings.proposal_engine_confidence (:66,695)
  long long total_iterations = 0;
  int unaccepted = 0, first_model_events = 0;
  for (int it = 0;; ++it) {                                                // the loop condition of :427 is always true
    if (it > 100) break;
    ++ms.proposals;
    pb.cpref = models.empty() ? nullptr : compound.data();
    pb.stream_base = 2 * (u64)it;
    Stats st;
    std::vector<int> inliers;
    Instance inst;
    const int found = gcransac_run(pb, Pl, seed, inst.m, inliers, st);
    pb.cpref = nullptr; pb.stream_base = 0;
    if (!found) continue;
    total_iterations += st.iterations;
    // isPutativeModelValid (:723-749)
    bool valid = (int)inliers.size() >= std::max(3, MP.min_point_number);
    if (valid) {
      preference_vector(pb, inst.m, T2, inst.pref);
      double dot = 0.0, na = 0.0, nc = 0.0;
      for (int i = 0; i < N; ++i) { dot += inst.pref[i] * compound[i]; na += inst.pref[i] * inst.pref[i]; nc += compound[i] * compound[i]; }
      const double tanimoto = dot / (na + nc - dot);
      if (MP.max_tanimoto < tanimoto) valid = false;
    }
    if (!valid) {
      ++unaccepted;
      if (unaccepted == 10) break;                                         // max_proposal_number_without_change
      continue;
    }
    models.push_back(inst);
    ++ms.accepted;
    if (models.size() == 1) {                                              // :520-529
      ++first_model_events;
      std::fill(labeling.begin(), labeling.end(), 1);
      for (int i : inliers) labeling[i] = 0;
    } else {                                                               // :531-547
      pearl.run(pb, nb, models, P.spatial_coherence_weight, MP.min_point_number);
      labeling = pearl.labels;
    }
    // updateCompoundModel (:755-794)
    if (!models.empty()) {
      std::fill(compound.begin(), compound.end(), 0.0);
      scores.assign(models.size(), 0.0);
      for (size_t k = 0; k < models.size(); ++k)
        for (int i = 0; i < N; ++i) {
          compound[i] = std::max(compound[i], models[k].pref[i]);
          scores[k] += models[k].pref[i];
        }
    }
    // predicted unseen inliers (:589-615, 651-673)
    const long long covered = models.size() == 1 ? (long long)first_model_events : (long long)N - (long long)pearl.outliers.size();
    const double ratio = std::pow(1.0 - std::pow(1.0 - MP.confidence, 1.0 / (double)total_iterations), 1.0 / 3.0);
    const long long unseen = (long long)std::llround((double)((long long)N - covered) * ratio);
    if (unseen < MP.min_point_number) break;
    if ((int)models.size() >= MP.max_model_number) break;
  }
  ms.ransac_iterations = (int)total_iterations;
  ms.pearl_iterations = pearl.iterations;
  ms.moves = pearl.moves;
}

}  // namespace ora

// ---------------------------------------------------------------------------------------------------
// C API for the Python test harness (oracle/posefit.py)
// ---------------------------------------------------------------------------------------------------
extern "C" {

typedef ora::Params ora_params;

void ora_params_default(ora_params* p) {
  p->threshold = 4.0; p->spatial_coherence_weight = 0.1; p->neighborhood_ball_radius = 20.0;
  p->scaling_from_millimeters = 0.1; p->min_triangle_area = 0.0; p->min_coverage = 0.5; p->confidence = 1.0;
  p->max_iters = 400; p->min_iters = 10; p->min_iters_before_lo = 20; p->max_lo_trials = 20; p->max_graph_cuts = 10;
  p->max_lsq_iters = 10; p->max_unsuccessful = 100; p->max_neighbors = 5; p->apply_numerical_optimization = 1;
}

unsigned long long ora_rng_u64(unsigned long long seed, unsigned long long stream, unsigned long long a,
                               unsigned long long b, unsigned long long c) {
  return ora::rng_u64(seed, stream, a, b, c);
}
int ora_unique_set(unsigned long long seed, unsigned long long stream, unsigned long long a, unsigned long long b, int n,
                   int k, int* out) {
  return ora::unique_set(seed, stream, a, b, n, k, out) ? 1 : 0;
}
int ora_quartic(const double* c, double* roots) { return ora::solve_quartic_real(c, roots); }
int ora_p3p(const double* pts7, const int* idx, double* models) { return ora::p3p(pts7, idx, models); }
void ora_rodrigues_to_matrix(const double* r, double* R, double* dRdr) { ora::rodrigues_to_matrix(r, R, dRdr); }
void ora_matrix_to_rodrigues(const double* R, double* r) { ora::matrix_to_rodrigues(R, r); }
int ora_solvepnp(int n, const double* X, const double* uv, int use_guess, double* param) {
  return ora::solvepnp_iterative(n, X, uv, use_guess != 0, param) ? 1 : 0;
}

// points -> problem helpers exposed for stage-by-stage comparison with the CUDA path
int ora_neighbors(int N, const double* x2d, const double* x3d, const double* K, const ora_params* P, int* out /* N x max_nbr, -1 padded */) {
  ora::Problem pb;
  ora::make_problem(N, x2d, x3d, K, *P, pb);
  ora::build_neighbors(pb, P->neighborhood_ball_radius, P->scaling_from_millimeters, P->max_neighbors);
  for (int i = 0; i < N; ++i)
    for (int k = 0; k < P->max_neighbors; ++k) out[i * P->max_neighbors + k] = k < (int)pb.nbr[i].size() ? pb.nbr[i][k] : -1;
  return 0;
}

int ora_neighbors_bruteforce(int N, const double* x2d, const double* x3d, const double* K, const ora_params* P, int* out) {
  ora::Problem pb;
  ora::make_problem(N, x2d, x3d, K, *P, pb);
  ora::build_neighbors_bruteforce(pb, P->neighborhood_ball_radius, P->scaling_from_millimeters, P->max_neighbors);
  for (int i = 0; i < N; ++i)
    for (int k = 0; k < P->max_neighbors; ++k) out[i * P->max_neighbors + k] = k < (int)pb.nbr[i].size() ? pb.nbr[i][k] : -1;
  return 0;
}

// score of one model: out = {value, inliers} with the early-out rule against best_inl
int ora_score(int N, const double* x2d, const double* x3d, const double* K, const ora_params* P, const double* model,
              long long best_inl, long long* out, int* inlier_mask) {
  ora::Problem pb;
  ora::make_problem(N, x2d, x3d, K, *P, pb);
  std::vector<int> inl;
  ora::Score s = ora::get_score(pb, model, best_inl, &inl);
  out[0] = (long long)s.value; out[1] = s.inliers; out[2] = pb.used_pixels;
  if (inlier_mask) { for (int i = 0; i < N; ++i) inlier_mask[i] = 0; for (int i : inl) inlier_mask[i] = 1; }
  return 0;
}

// hypotheses of main-loop pass `pass`: returns number of models; fails_out = failed attempts
int ora_generate_models(int N, const double* x2d, const double* x3d, const double* K, const ora_params* P,
                        unsigned long long seed, int pass, double* models, int* fails_out, int* sample_out) {
  ora::Problem pb;
  ora::make_problem(N, x2d, x3d, K, *P, pb);
  return ora::generate_models(pb, *P, seed, pass, models, fails_out, sample_out);
}

// graph-cut labeling of one model with injected (nbr_offsets != NULL) or built neighbour lists
int ora_labeling(int N, const double* x2d, const double* x3d, const double* K, const ora_params* P, const double* model,
                 const int* nbr_offsets, const int* nbr_index, int* labels) {
  ora::Problem pb;
  ora::make_problem(N, x2d, x3d, K, *P, pb);
  if (nbr_offsets) {
    pb.nbr.assign(N, std::vector<int>());
    for (int i = 0; i < N; ++i) pb.nbr[i].assign(nbr_index + nbr_offsets[i], nbr_index + nbr_offsets[i + 1]);
  } else {
    ora::build_neighbors(pb, P->neighborhood_ball_radius, P->scaling_from_millimeters, P->max_neighbors);
  }
  std::vector<int> inl;
  ora::labeling(pb, model, P->spatial_coherence_weight, inl);
  for (int i = 0; i < N; ++i) labels[i] = 0;
  for (int i : inl) labels[i] = 1;
  return (int)inl.size();
}

// the cut graph itself (terminal capacities + n-links), for cross-checking against the reference's BK max-flow
int ora_cut_graph(int N, const double* x2d, const double* x3d, const double* K, const ora_params* P, const double* model,
                  const int* nbr_offsets, const int* nbr_index, double* tr, int* ex, int* ey, double* cxy, double* cyx,
                  double* u0, double* u1, double* e00, int max_edges) {
  ora::Problem pb;
  ora::make_problem(N, x2d, x3d, K, *P, pb);
  pb.nbr.assign(N, std::vector<int>());
  for (int i = 0; i < N; ++i) pb.nbr[i].assign(nbr_index + nbr_offsets[i], nbr_index + nbr_offsets[i + 1]);
  ora::CutGraph g;
  ora::build_cut_graph(pb, model, P->spatial_coherence_weight, g);
  int E = (int)g.ex.size();
  if (E > max_edges) return -1;
  for (int i = 0; i < N; ++i) { tr[i] = g.tr[i]; u0[i] = g.u0[i]; u1[i] = g.u1[i]; }
  for (int e = 0; e < E; ++e) { ex[e] = g.ex[e]; ey[e] = g.ey[e]; cxy[e] = g.cxy[e]; cyx[e] = g.cyx[e]; e00[e] = g.e00[e]; }
  return E;
}

// LO trace of the last ora_find6dposes call: returns the number of rounds, copies up to cap rows of 72 ints
int ora_last_trace(int* out, int cap) {
  int rounds = (int)(ora::g_trace.size() / 72);
  for (int r = 0; r < rounds && r < cap; ++r) std::memcpy(out + 72 * r, ora::g_trace.data() + 72 * r, 72 * sizeof(int));
  return rounds;
}

// find6DPoses, single-instance branch.  stats: {iterations, graph_cuts, lo_runs, passes, found}
int ora_find6dposes(int N, const double* x2d, const double* x3d, const double* K, const ora_params* P,
                    unsigned long long seed, const int* nbr_offsets, const int* nbr_index, double* pose,
                    int* labeling_out, int* stats) {
  ora::Stats st;
  ora::g_trace.clear();
  int r = ora::find6dposes(N, x2d, x3d, K, *P, seed, nbr_offsets, nbr_index, pose, labeling_out, st);
  if (stats) { stats[0] = st.iterations; stats[1] = st.graph_cuts; stats[2] = st.lo_runs; stats[3] = st.passes; stats[4] = st.found; }
  return r;
}

// find6DPoses, multi-instance branch (max_model_number != 1).  poses [max_out][12], scores [max_out], labeling [N];
// stats: {proposals, accepted, ransac_iterations, pearl_iterations, expansion_moves, sped_up}.  Returns the number of
// instances found (at most max_out are written).
int ora_find6dposes_multi(int N, const double* x2d, const double* x3d, const double* K, const ora_params* P,
                          unsigned long long seed, int max_model_number, int max_model_number_for_pearl, double conf,
                          double max_tanimoto, int min_point_number, const int* nbr_offsets, const int* nbr_index,
                          int max_out, double* poses, int* labeling_out, double* scores, int* stats) {
  ora::Problem pb;
  ora::make_problem(N, x2d, x3d, K, *P, pb);
  if (nbr_offsets) {
    pb.nbr.assign(N, std::vector<int>());
    for (int i = 0; i < N; ++i) pb.nbr[i].assign(nbr_index + nbr_offsets[i], nbr_index + nbr_offsets[i + 1]);
  } else {
    ora::build_neighbors(pb, P->neighborhood_ball_radius, P->scaling_from_millimeters, P->max_neighbors);
  }
  ora::MultiParams MP = {max_model_number, max_model_number_for_pearl, conf, max_tanimoto, min_point_number};
  std::vector<ora::Instance> models;
  std::vector<int> lab;
  std::vector<double> sc;
  ora::MultiStats ms;
  ora::g_trace.clear();
  ora::progx_run(pb, *P, MP, seed, models, lab, sc, ms);
  for (int i = 0; i < N; ++i) labeling_out[i] = lab[i];
  for (size_t k = 0; k < models.size() && (int)k < max_out; ++k) {
    std::memcpy(poses + 12 * k, models[k].m, 12 * sizeof(double));
    scores[k] = sc[k];
  }
  if (stats) { stats[0] = ms.proposals; stats[1] = ms.accepted; stats[2] = ms.ransac_iterations; stats[3] = ms.pearl_iterations; stats[4] = ms.moves; stats[5] = ms.sped_up; }
  return (int)models.size();
}

// alpha-expansion on an explicit problem (tests: cross-check against the reference's GCoptimization in oracle/_ref).
// D [n][L] data costs, neighbour listings as CSR (listing j of site i -> setNeighbors(i, j)), Potts weight lambda,
// uniform label cost; labels in/out (initial labeling).  Returns the energy.
double ora_alpha_expansion(int n, int L, const double* D, const int* nbr_offsets, const int* nbr_index, double lambda,
                           double label_cost, int* labels, int max_iterations) {
  ora::AlphaExpansion ax;
  ax.n = n; ax.L = L; ax.lambda = lambda; ax.label_cost = label_cost;
  ax.D.assign(D, D + (size_t)n * L);
  std::vector<std::vector<int>> nbr(n);
  for (int i = 0; i < n; ++i) nbr[i].assign(nbr_index + nbr_offsets[i], nbr_index + nbr_offsets[i + 1]);
  ora::gco_neighbors(nbr, ax.nb);
  ax.label.assign(labels, labels + n);
  double e = ax.expansion(max_iterations);
  for (int i = 0; i < n; ++i) labels[i] = ax.label[i];
  return e;
}

}  // extern "C"
