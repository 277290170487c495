// Device-side f64 geometry for the pose-fitting kernels (sm_100a): counter-based random stream, quartic roots,
// Kneip P3P, Rodrigues, small linear algebra.  Each routine cites the reference code whose behaviour it provides
// (paths relative to /root/reference/external/progressive-x/graph-cut-ransac/src/pygcransac/include).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>

namespace epos {
namespace pose {

typedef unsigned long long u64;

// ---- random stream (DESIGN.md "RANSAC random stream"): replaces std::mt19937(random_device()) of
// uniform_random_generator.h:51-54, which has no reproducible stream --------------------------------------
__device__ __forceinline__ u64 mix64(u64 z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ u64 rng_u64(u64 seed, u64 stream, u64 a, u64 b, u64 c) {
  return mix64(mix64(mix64(mix64(seed ^ (stream * 0xD6E8FEB86659FD93ULL)) + a) + b) + c);
}
__device__ __forceinline__ int rng_index(u64 seed, u64 stream, u64 a, u64 b, u64 c, int n) {
  return (int)__umul64hi(rng_u64(seed, stream, a, b, c), (u64)n);
}
// k distinct indices in [0,n): a draw equal to an accepted one is discarded (uniform_random_generator.h:85-97)
__device__ inline bool unique_set(u64 seed, u64 stream, u64 a, u64 b, int n, int k, int* out) {
  if (k > n) return false;
  u64 c = 0;
  for (int i = 0; i < k;) {
    const int v = rng_index(seed, stream, a, b, c++, n);
    bool dup = false;
    for (int j = 0; j < i; ++j) dup |= (out[j] == v);
    if (!dup) out[i++] = v;
  }
  return true;
}

// ---- 3-vector / 3x3 helpers ----------------------------------------------------------------------------
__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void normalize3(double* a) {
  const double n = sqrt(dot3(a, a));
  a[0] /= n; a[1] /= n; a[2] /= n;
}
__device__ __forceinline__ double det3(const double* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
__device__ inline bool inv3(const double* m, double* o) {
  const double d = det3(m);
  if (d == 0.0) return false;
  const double id = 1.0 / d;
  o[0] = (m[4] * m[8] - m[5] * m[7]) * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = (m[5] * m[6] - m[3] * m[8]) * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = (m[3] * m[7] - m[4] * m[6]) * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return true;
}

// orthogonal polar factor (= U V^T of the SVD) of a 3x3 matrix with det > 0: R <- (R + R^-T)/2
__device__ inline bool polar_rotation(const double* M, double* R) {
  for (int i = 0; i < 9; ++i) R[i] = M[i];
  for (int it = 0; it < 100; ++it) {
    double inv[9], Rn[9];
    if (!inv3(R, inv)) return false;
    double diff = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        Rn[i * 3 + j] = 0.5 * (R[i * 3 + j] + inv[j * 3 + i]);
        const double e = Rn[i * 3 + j] - R[i * 3 + j];
        diff += e * e;
      }
    for (int i = 0; i < 9; ++i) R[i] = Rn[i];
    if (diff < 1e-30) break;
  }
  return true;
}

// Jacobi eigen-solvers in the round-robin parallel ordering (each step of a sweep rotates n/2 disjoint index pairs at
// once: A <- J^T A J, all angles taken from A before the step).  Stand in for the SVDs of OpenCV's
// cvFindExtrinsicCameraParams2 / cvFindHomography (solver_epnp_lm.h:139-146 call path).
__device__ __forceinline__ void jacobi_pair(int m, int step, int k, int* p, int* q) {
  const int a = k == 0 ? m - 1 : (step + k) % (m - 1);
  const int b = k == 0 ? step : (step - k + (m - 1)) % (m - 1);
  *p = a < b ? a : b;
  *q = a < b ? b : a;
}
__device__ __forceinline__ void jacobi_angle(const double* A, int n, int p, int q, double* c, double* s) {
  const double apq = A[p * n + q];
  *c = 1.0; *s = 0.0;
  if (apq != 0.0) {
    const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
    const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    *c = 1.0 / sqrt(t * t + 1.0);
    *s = t * (*c);
  }
}

// serial version for n = 3 (every thread computes it redundantly on private copies); V columns = eigenvectors
__device__ inline void jacobi_eig3(double* A, double* V, double* w) {
  const int n = 3, m = 4;
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        const double v = A[i * n + j] * A[i * n + j];
        if (i == j) diag += v; else off += v;
      }
    if (off <= 1e-300 || off < 1e-28 * diag) break;
    for (int step = 0; step < m - 1; ++step)
      for (int k = 0; k < m / 2; ++k) {                  // at most one real pair per step for n = 3
        int p, q;
        jacobi_pair(m, step, k, &p, &q);
        if (q >= n) continue;
        double c, s;
        jacobi_angle(A, n, p, q, &c, &s);
        for (int r = 0; r < n; ++r) {
          const double akp = A[r * n + p], akq = A[r * n + q];
          A[r * n + p] = c * akp - s * akq;
          A[r * n + q] = s * akp + c * akq;
          const double vkp = V[r * n + p], vkq = V[r * n + q];
          V[r * n + p] = c * vkp - s * vkq;
          V[r * n + q] = s * vkp + c * vkq;
        }
        for (int r = 0; r < n; ++r) {
          const double apk = A[p * n + r], aqk = A[q * n + r];
          A[p * n + r] = c * apk - s * aqk;
          A[q * n + r] = s * apk + c * aqk;
        }
      }
  }
  for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
}

// Warp-cooperative version on a symmetric n x n matrix (n <= 12) held in shared memory.  All 32 lanes must call.
// A is destroyed; V (n x n) receives eigenvectors in columns, w the eigenvalues.  Lane j < n/2 owns pair j of a step.
__device__ inline void jacobi_eig_warp(int n, double* A, double* V, double* w, int lane) {
  for (int i = lane; i < n * n; i += 32) V[i] = (i / n == i % n) ? 1.0 : 0.0;
  __syncwarp();
  const int m = (n + 1) & ~1, half = m / 2;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = lane; i < n * n; i += 32) {
      const double v = A[i] * A[i];
      if (i / n == i % n) diag += v; else off += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      off += __shfl_xor_sync(0xffffffffu, off, o);
      diag += __shfl_xor_sync(0xffffffffu, diag, o);
    }
    if (off <= 1e-300 || off < 1e-28 * diag) break;
    for (int step = 0; step < m - 1; ++step) {
      int p = 0, q = n;                                  // q >= n marks "no rotation" for this lane
      double c = 1.0, s = 0.0;
      if (lane < half) {
        jacobi_pair(m, step, lane, &p, &q);
        if (q < n) jacobi_angle(A, n, p, q, &c, &s);
      }
      __syncwarp();
      const int items = half * n;
      for (int it = lane; it < ((items + 31) & ~31); it += 32) {        // A <- A J, V <- V J
        const int j = it < items ? it / n : 0, k = it < items ? it % n : 0;
        const int pj = __shfl_sync(0xffffffffu, p, j), qj = __shfl_sync(0xffffffffu, q, j);
        const double cj = __shfl_sync(0xffffffffu, c, j), sj = __shfl_sync(0xffffffffu, s, j);
        if (it < items && qj < n) {
          const double akp = A[k * n + pj], akq = A[k * n + qj];
          A[k * n + pj] = cj * akp - sj * akq;
          A[k * n + qj] = sj * akp + cj * akq;
          const double vkp = V[k * n + pj], vkq = V[k * n + qj];
          V[k * n + pj] = cj * vkp - sj * vkq;
          V[k * n + qj] = sj * vkp + cj * vkq;
        }
      }
      __syncwarp();
      for (int it = lane; it < ((items + 31) & ~31); it += 32) {        // A <- J^T A
        const int j = it < items ? it / n : 0, k = it < items ? it % n : 0;
        const int pj = __shfl_sync(0xffffffffu, p, j), qj = __shfl_sync(0xffffffffu, q, j);
        const double cj = __shfl_sync(0xffffffffu, c, j), sj = __shfl_sync(0xffffffffu, s, j);
        if (it < items && qj < n) {
          const double apk = A[pj * n + k], aqk = A[qj * n + k];
          A[pj * n + k] = cj * apk - sj * aqk;
          A[qj * n + k] = sj * apk + cj * aqk;
        }
      }
      __syncwarp();
    }
  }
  __syncwarp();
  if (lane < n) w[lane] = A[lane * n + lane];
  __syncwarp();
}

// 6x6 Gaussian elimination with partial pivoting, fully unrolled so that the matrix lives in registers
// (same operation order as a textbook row-major elimination: the oracle's solve_linear).
__device__ __forceinline__ bool solve_linear6(double* A_, double* b_, double* x) {
  double A[6][6], b[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    b[r] = b_[r];
#pragma unroll
    for (int c = 0; c < 6; ++c) A[r][c] = A_[r * 6 + c];
  }
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    double best = fabs(A[c][c]);
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      const double v = fabs(A[r][c]);
      if (v > best) { best = v; piv = r; }
    }
    if (best == 0.0) ok = false;
#pragma unroll
    for (int r = c + 1; r < 6; ++r)
      if (piv == r) {
#pragma unroll
        for (int k = 0; k < 6; ++k) { const double t = A[c][k]; A[c][k] = A[r][k]; A[r][k] = t; }
        const double t = b[c]; b[c] = b[r]; b[r] = t;
      }
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      const double f = A[r][c] / A[c][c];
#pragma unroll
      for (int k = c; k < 6; ++k) A[r][k] -= f * A[c][k];
      b[r] -= f * b[c];
    }
  }
  if (!ok) return false;
#pragma unroll
  for (int r = 5; r >= 0; --r) {
    double s = b[r];
#pragma unroll
    for (int k = r + 1; k < 6; ++k) s -= A[r][k] * x[k];
    x[r] = s / A[r][r];
  }
  return true;
}

// ---- cv::Rodrigues (solver_epnp_lm.h:136,149; progressivex_python.cpp:289,302) --------------------------------
__device__ inline void rodrigues_to_matrix(const double* r, double* R, double* dRdr /* 3x9 or nullptr */) {
  const double theta = sqrt(dot3(r, r));
  if (theta < DBL_EPSILON) {
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (dRdr) {
      for (int i = 0; i < 27; ++i) dRdr[i] = 0.0;
      dRdr[5] = -1; dRdr[7] = 1; dRdr[9 + 2] = 1; dRdr[9 + 6] = -1; dRdr[18 + 1] = -1; dRdr[18 + 3] = 1;
    }
    return;
  }
  const double c = cos(theta), s = sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
  const double k[3] = {r[0] * it, r[1] * it, r[2] * it};
  const double kkt[9] = {k[0] * k[0], k[0] * k[1], k[0] * k[2], k[0] * k[1], k[1] * k[1], k[1] * k[2],
                         k[0] * k[2], k[1] * k[2], k[2] * k[2]};
  const double kx[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
  for (int i = 0; i < 9; ++i) R[i] = c * ((i % 4 == 0) ? 1.0 : 0.0) + c1 * kkt[i] + s * kx[i];
  if (dRdr) {
    const double dkkt[27] = {2 * k[0], k[1], k[2], k[1], 0, 0, k[2], 0, 0,
                             0, k[0], 0, k[0], 2 * k[1], k[2], 0, k[2], 0,
                             0, 0, k[0], 0, 0, k[1], k[0], k[1], 2 * k[2]};
    const double dkx[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
    for (int i = 0; i < 3; ++i) {
      const double ki = k[i];
      const double a0 = -s * ki, a1 = (s - 2 * c1 * it) * ki, a2 = c1 * it, a3 = (c - s * it) * ki, a4 = s * it;
      for (int j = 0; j < 9; ++j)
        dRdr[i * 9 + j] = a0 * ((j % 4 == 0) ? 1.0 : 0.0) + a1 * kkt[j] + a2 * dkkt[i * 9 + j] + a3 * kx[j] +
                          a4 * dkx[i * 9 + j];
    }
  }
}

__device__ inline void matrix_to_rodrigues(const double* Rin, double* r) {
  double R[9];
  if (!polar_rotation(Rin, R))
    for (int i = 0; i < 9; ++i) R[i] = Rin[i];
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  double theta = acos(c);
  if (s < 1e-5) {
    if (c > 0) { r[0] = r[1] = r[2] = 0.0; return; }
    double t;
    t = (R[0] + 1) * 0.5; rx = sqrt(fmax(t, 0.0));
    t = (R[4] + 1) * 0.5; ry = sqrt(fmax(t, 0.0)) * (R[1] < 0 ? -1.0 : 1.0);
    t = (R[8] + 1) * 0.5; rz = sqrt(fmax(t, 0.0)) * (R[2] < 0 ? -1.0 : 1.0);
    if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
    theta /= sqrt(rx * rx + ry * ry + rz * rz);
    r[0] = rx * theta; r[1] = ry * theta; r[2] = rz * theta;
    return;
  }
  const double vth = theta / (2.0 * s);
  r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
}

// ---- real roots of a quartic (stands in for Eigen::PolynomialSolver<double,4>::realRoots, solver_p3p.h:175-177):
// Ferrari + Newton polish, ascending order -----------------------------------------------------------------------
__device__ inline double cubic_largest_real_root(double A, double B, double C) {
  const double a3 = A / 3.0;
  const double p = B - A * a3, q = 2.0 * a3 * a3 * a3 - a3 * B + C;
  const double disc = q * q / 4.0 + p * p * p / 27.0;
  double t;
  if (disc > 0) {
    const double sq = sqrt(disc);
    t = cbrt(-q / 2.0 + sq) + cbrt(-q / 2.0 - sq);
  } else if (p == 0.0) {
    t = cbrt(-q);
  } else {
    const double m = 2.0 * sqrt(-p / 3.0);
    double arg = 3.0 * q / (p * m);
    arg = arg > 1.0 ? 1.0 : (arg < -1.0 ? -1.0 : arg);
    t = m * cos(acos(arg) / 3.0);
  }
  double x = t - a3;
  for (int i = 0; i < 3; ++i) {
    const double f = ((x + A) * x + B) * x + C, df = (3.0 * x + 2.0 * A) * x + B;
    if (df == 0.0) break;
    const double xn = x - f / df;
    if (!isfinite(xn)) break;
    x = xn;
  }
  return x;
}

__device__ inline int quadratic_real(double b, double c, double* r) {
  const double disc = b * b - 4.0 * c;
  if (!(disc >= 0.0)) return 0;
  const double sq = sqrt(disc);
  const double q = -0.5 * (b + (b >= 0 ? sq : -sq));
  r[0] = q;
  r[1] = (q != 0.0) ? c / q : 0.0;
  return 2;
}

__device__ inline int solve_quartic_real(const double* c, double* roots) {
  if (c[4] == 0.0 || !isfinite(c[4])) return 0;
  const double a = c[3] / c[4], b = c[2] / c[4], cc = c[1] / c[4], d = c[0] / c[4];
  if (!(isfinite(a) && isfinite(b) && isfinite(cc) && isfinite(d))) return 0;
  const double a4 = a / 4.0;
  const double p = b - 6.0 * a4 * a4;
  const double q = cc - 2.0 * b * a4 + 8.0 * a4 * a4 * a4;
  const double r = d - cc * a4 + b * a4 * a4 - 3.0 * a4 * a4 * a4 * a4;
  double y[4];
  int n = 0;
  const double m = cubic_largest_real_root(p, p * p / 4.0 - r, -q * q / 8.0);
  if (m > 0.0 && fabs(q) > 0.0) {
    const double s = sqrt(2.0 * m);
    n += quadratic_real(-s, p / 2.0 + m + q / (2.0 * s), y + n);
    n += quadratic_real(s, p / 2.0 + m - q / (2.0 * s), y + n);
  } else {
    double z[2];
    const int nz = quadratic_real(p, r, z);
    for (int i = 0; i < nz; ++i)
      if (z[i] >= 0.0) { const double s = sqrt(z[i]); y[n++] = s; y[n++] = -s; }
  }
  for (int i = 0; i < n; ++i) {
    double x = y[i] - a4;
    for (int it = 0; it < 3; ++it) {
      const double f = (((c[4] * x + c[3]) * x + c[2]) * x + c[1]) * x + c[0];
      const double df = ((4.0 * c[4] * x + 3.0 * c[3]) * x + 2.0 * c[2]) * x + c[1];
      if (df == 0.0) break;
      const double xn = x - f / df;
      if (!isfinite(xn)) break;
      x = xn;
    }
    roots[i] = x;
  }
  for (int i = 1; i < n; ++i) {                        // insertion sort, ascending
    const double v = roots[i];
    int j = i - 1;
    while (j >= 0 && roots[j] > v) { roots[j + 1] = roots[j]; --j; }
    roots[j + 1] = v;
  }
  return n;
}

// ---- Kneip P3P (solver_p3p.h:239-371, :111-187 plane rotation, :192-237 back-substitution) followed by the
// admissibility filter of perspective_n_point_estimator.h:110-131.  un/vn: normalised image coords of the 3 sample
// points, X: their 3D points.  models: up to 4 row-major 3x4 [R|t]. ------------------------------------------------
__device__ inline int p3p_kneip(const double* un, const double* vn, const double (*X)[3], double* models) {
  double f[3][3], P[3][3];
  for (int i = 0; i < 3; ++i) {
    f[i][0] = un[i]; f[i][1] = vn[i]; f[i][2] = 1.0;
    normalize3(f[i]);
    P[i][0] = X[i][0]; P[i][1] = X[i][1]; P[i][2] = X[i][2];
  }
  double e1[3], e2[3], cr[3];
  for (int k = 0; k < 3; ++k) { e1[k] = P[1][k] - P[0][k]; e2[k] = P[2][k] - P[0][k]; }
  cross3(e1, e2, cr);
  if (dot3(cr, cr) < 1e-6) return 0;
  double T[3][3], f2c[3];
  for (int pass = 0; pass < 2; ++pass) {
    for (int k = 0; k < 3; ++k) T[0][k] = f[0][k];
    cross3(f[0], f[1], T[2]);
    normalize3(T[2]);
    cross3(T[2], T[0], T[1]);
    for (int k = 0; k < 3; ++k) f2c[k] = dot3(T[k], f[2]);
    if (pass == 1 || !(f2c[2] > 0)) break;
    for (int k = 0; k < 3; ++k) {
      double t = f[0][k]; f[0][k] = f[1][k]; f[1][k] = t;
      t = P[0][k]; P[0][k] = P[1][k]; P[1][k] = t;
    }
    for (int k = 0; k < 3; ++k) { e1[k] = P[1][k] - P[0][k]; e2[k] = P[2][k] - P[0][k]; }
  }
  if (fabs(f2c[2]) < DBL_EPSILON) return 0;
  double Nw[3][3], P2w[3];
  const double d12 = sqrt(dot3(e1, e1));
  for (int k = 0; k < 3; ++k) Nw[0][k] = e1[k] / d12;
  cross3(Nw[0], e2, Nw[2]);
  normalize3(Nw[2]);
  cross3(Nw[2], Nw[0], Nw[1]);
  for (int k = 0; k < 3; ++k) P2w[k] = dot3(Nw[k], e2);
  const double f1 = f2c[0] / f2c[2], f2 = f2c[1] / f2c[2], p1 = P2w[0], p2 = P2w[1];
  const double cos_beta = dot3(f[0], f[1]);
  double b = 1.0 / (1.0 - cos_beta * cos_beta) - 1.0;
  b = cos_beta < 0 ? -sqrt(b) : sqrt(b);
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
  const int nr = solve_quartic_real(c, roots);
  int nm = 0;
  for (int i = 0; i < nr; ++i) {
    const double ct = roots[i] > 1.0 ? 1.0 : (roots[i] < -1.0 ? -1.0 : roots[i]);
    const double cot_a = (-f1 * p1 / f2 - ct * p2 + D * b) / (-f1 * ct * p2 / f2 + p1 - D);
    const double st = sqrt(1.0 - ct * ct);
    const double sa = sqrt(1.0 / (cot_a * cot_a + 1.0));
    double ca = sqrt(1.0 - sa * sa);
    if (cot_a < 0) ca = -ca;
    const double g = sa * b + ca;
    const double cnu[3] = {D * ca * g, ct * D * sa * g, st * D * sa * g};
    double C[3];
    for (int k = 0; k < 3; ++k) C[k] = P[0][k] + Nw[0][k] * cnu[0] + Nw[1][k] * cnu[1] + Nw[2][k] * cnu[2];
    const double Q[3][3] = {{-ca, -sa * ct, -sa * st}, {sa, -ca * ct, -ca * st}, {0.0, -st, ct}};
    double QN[3][3], R[9];
    for (int r_ = 0; r_ < 3; ++r_)
      for (int k = 0; k < 3; ++k) QN[r_][k] = Q[r_][0] * Nw[0][k] + Q[r_][1] * Nw[1][k] + Q[r_][2] * Nw[2][k];
    for (int r_ = 0; r_ < 3; ++r_)
      for (int k = 0; k < 3; ++k) R[r_ * 3 + k] = T[0][r_] * QN[0][k] + T[1][r_] * QN[1][k] + T[2][r_] * QN[2][k];
    double t[3];
    for (int r_ = 0; r_ < 3; ++r_) t[r_] = -(R[r_ * 3] * C[0] + R[r_ * 3 + 1] * C[1] + R[r_ * 3 + 2] * C[2]);
    bool fin = true;
    for (int k = 0; k < 9; ++k) fin &= isfinite(R[k]);
    for (int k = 0; k < 3; ++k) fin &= isfinite(t[k]);
    if (!fin) continue;
    if (t[2] < 0.0 || det3(R) < -0.95) continue;
    double* m = models + 12 * nm++;
    for (int r_ = 0; r_ < 3; ++r_) {
      m[r_ * 4 + 0] = R[r_ * 3]; m[r_ * 4 + 1] = R[r_ * 3 + 1]; m[r_ * 4 + 2] = R[r_ * 3 + 2]; m[r_ * 4 + 3] = t[r_];
    }
  }
  return nm;
}

// squared reprojection error (perspective_n_point_estimator.h:133-170), explicit fma chain
__device__ __forceinline__ double sq_residual(double un, double vn, double x, double y, double z, const double* m) {
  const double px = fma(m[0], x, fma(m[1], y, fma(m[2], z, m[3])));
  const double py = fma(m[4], x, fma(m[5], y, fma(m[6], z, m[7])));
  const double pz = fma(m[8], x, fma(m[9], y, fma(m[10], z, m[11])));
  const double du = px / pz - un, dv = py / pz - vn;
  return fma(du, du, dv * dv);
}

}  // namespace pose
}  // namespace epos
