// 2D-3D correspondence extraction on sm_100a: establish_many_to_many (/root/reference/epos_lib/corresp.py:9-101)
// plus the confidence top-K of scripts/infer.py:425-440, for every (image, object slot) segment of a batch.
//
// One CTA per segment.  HBM-bound: the segment's obj_conf column and, for masked pixels only, the F-float frag_conf row
// (one coalesced 256-byte read per warp at F = 64) and 12 bytes of frag_loc per emitted row.
//   phase 1  warp per pixel: obj_conf > tau_a, row max, frag_conf > tau_b * max  -> per-pixel counts -> block scan
//            (row-major pixel, then fragment: the emission order of the reference)
//   phase 2  (only if the segment has more than max_corr rows) radix select of the max_corr largest 64-bit keys
//            (conf bits << 32 | emission index), i.e. np.argsort(conf)[::-1][:K] with ties by descending index,
//            then a shared-memory bitonic sort of the K survivors
//   phase 3  rows: coord_2d = (x + 0.5, y + 0.5) / output_scale (misc.py:14-26), coord_3d = centre[f] +
//            f32(loc * size[f]) (corresp.py:70-78), conf = obj_conf * frag_conf
#include <stdlib.h>
#include "common.cuh"

namespace epos {

constexpr int CT = 512;             // threads per CTA
constexpr int CW = CT / 32;
constexpr int MAXF = 512;           // fragments per object supported
constexpr int SORT_MAX = 4096;      // max_corr supported by the in-kernel sort

struct CorrArgs {
  const float* obj_conf; const float* frag_conf; const float* frag_loc;
  int B, h, w, O, F;
  const int* obj_ids; int J;
  const double* centers; const double* sizes;
  double inv_scale; float min_obj_conf, min_rel;
  int cap, max_corr;
  double* c2d; double* c3d; float* conf; float* conf_obj; float* conf_frag; int* px; int* frag; int* counts; int* totals;
  unsigned int* ws;                // workspace [B*J][4][HW+1]: masked-pixel list, obj_conf, row max, emission offsets
  // lazy localisation head (frag_loc == NULL): pred_frag_loc is never materialised; the three local coordinates of a
  // surviving (pixel, object, fragment) row are computed here from the decoder features and the logit weights
  const uint16_t* feat; int ldf; long long feat_plane;     // split-bf16 decoder features [2][B*h*w][ldf], C = feat_c
  const float* w_loc; const float* b_loc; int feat_c;      // logits/pred_frag_loc weights [O*F*3][feat_c] f32, bias
};

__device__ inline int block_scan_excl(int v, int* sh, int* total) {
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
    int s = lane < CW ? sh[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    if (lane < CW) sh[lane] = s;
  }
  __syncthreads();
  const int base = w > 0 ? sh[w - 1] : 0;
  *total = sh[CW - 1];
  __syncthreads();
  return base + x - v;
}

// Per-segment view of the masked pixels (ordered by pixel index).
struct SegLists {
  const unsigned int* mpix;   // [nm]   pixel index
  const float* vobj;          // [nm]   obj_conf of the pixel
  const float* rmax;          // [nm]   max_f frag_conf
  const unsigned int* off;    // [nm+1] emission offset of the pixel's first row
  int nm;
};

// Visits every candidate row of the segment, ONE THREAD PER MASKED PIXEL (many rows in flight per SM; the F-float row is
// read with 16-byte loads).  fn(p, f, e, vobj, vfrag) with e = emission index (row-major pixel, then fragment).
template <class Fn>
__device__ inline void for_each_candidate(const CorrArgs& a, int b, int obj_id, const SegLists& L, Fn fn) {
  const int HW = a.h * a.w;
  const float* fcb = a.frag_conf + ((size_t)b * HW * a.O + (obj_id - 1)) * a.F;
  const bool vec = (a.F % 4 == 0);
  for (int m = threadIdx.x; m < L.nm; m += CT) {
    const unsigned int e0 = L.off[m];
    if (L.off[m + 1] == e0) continue;
    const int p = (int)L.mpix[m];
    const float vobj = L.vobj[m];
    const float thr = __fmul_rn(L.rmax[m], a.min_rel);              // corresp.py:62-64 (float32 product)
    const float* row = fcb + (size_t)p * a.O * a.F;
    unsigned int e = e0;
    if (vec) {
      for (int f = 0; f < a.F; f += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row + f));
        if (v.x > thr) fn(p, f, e++, vobj, v.x);
        if (v.y > thr) fn(p, f + 1, e++, vobj, v.y);
        if (v.z > thr) fn(p, f + 2, e++, vobj, v.z);
        if (v.w > thr) fn(p, f + 3, e++, vobj, v.w);
      }
    } else {
      for (int f = 0; f < a.F; ++f) {
        const float v = __ldg(row + f);
        if (v > thr) fn(p, f, e++, vobj, v);
      }
    }
  }
}

__device__ __forceinline__ unsigned long long make_key(float vobj, float vfrag, unsigned int e) {
  const float c = __fmul_rn(vobj, vfrag);                        // corresp.py:95 conf = conf_obj * conf_frag (f32)
  return ((unsigned long long)__float_as_uint(c) << 32) | (unsigned long long)e;
}

__device__ inline void write_row(const CorrArgs& a, int b, int obj_id, size_t dst, int p, int f, float vobj, float vfrag) {
  const int x = p % a.w, y = p / a.w;
  a.c2d[2 * dst] = a.inv_scale * ((double)x + 0.5);
  a.c2d[2 * dst + 1] = a.inv_scale * ((double)y + 0.5);
  const int HW = a.h * a.w;
  if (a.frag_loc) {
    const float* loc = a.frag_loc + ((((size_t)b * HW + p) * a.O + (obj_id - 1)) * a.F + f) * 3;
    const double* cen = a.centers + ((size_t)(obj_id - 1) * a.F + f) * 3;
    const double sz = a.sizes[(size_t)(obj_id - 1) * a.F + f];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float l32 = (float)((double)__ldg(loc + k) * sz);      // numpy: f32 array *= f64 array -> computed in f64, cast to f32
      a.c3d[3 * dst + k] = cen[k] + (double)l32;
    }
  }
  a.conf[dst] = __fmul_rn(vobj, vfrag);
  a.conf_obj[dst] = vobj;
  a.conf_frag[dst] = vfrag;
  a.px[dst] = p;
  a.frag[dst] = f;
}

// Lazy localisation head: rows [0, n) of the segment already carry px / frag; one warp per row computes
//   loc[k] = b_loc[c] + sum_i feat[p][i] * w_loc[c][i],  c = ((obj_id-1) F + f) 3 + k      (model.py:448-456, 1x1 conv + bias)
// in fp32 (feat = hi + lo is exact in fp32; lanes own interleaved 4-element groups of the channel axis, partial sums are
// combined by a butterfly), then coord_3d as in write_row.  Replaces the dense 3 O F-column GEMM whose output is read at
// <= max_corr rows per (image, object).
__device__ inline void lazy_loc_rows(const CorrArgs& a, int b, int obj_id, size_t seg_base, int n) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HW = a.h * a.w;
  for (int i = warp; i < n; i += CW) {
    const size_t dst = seg_base + i;
    const int p = a.px[dst], f = a.frag[dst];
    const uint16_t* fh = a.feat + ((size_t)b * HW + p) * a.ldf;
    const uint16_t* fl = fh + a.feat_plane;
    const size_t c0 = ((size_t)(obj_id - 1) * a.F + f) * 3;
    const float* w0 = a.w_loc + c0 * a.feat_c;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int k = lane * 4; k < a.feat_c; k += 128) {
      const uint2 h = __ldg(reinterpret_cast<const uint2*>(fh + k)), l = __ldg(reinterpret_cast<const uint2*>(fl + k));
      const float x0 = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
      const float x1 = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
      const float x2 = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
      const float x3 = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
      const float4 wa = __ldg(reinterpret_cast<const float4*>(w0 + k));
      const float4 wb = __ldg(reinterpret_cast<const float4*>(w0 + a.feat_c + k));
      const float4 wc = __ldg(reinterpret_cast<const float4*>(w0 + 2 * a.feat_c + k));
      s0 = fmaf(x0, wa.x, fmaf(x1, wa.y, fmaf(x2, wa.z, fmaf(x3, wa.w, s0))));
      s1 = fmaf(x0, wb.x, fmaf(x1, wb.y, fmaf(x2, wb.z, fmaf(x3, wb.w, s1))));
      s2 = fmaf(x0, wc.x, fmaf(x1, wc.y, fmaf(x2, wc.z, fmaf(x3, wc.w, s2))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane < 3) {
      const float loc = (lane == 0 ? s0 : (lane == 1 ? s1 : s2)) + (a.b_loc ? __ldg(a.b_loc + c0 + lane) : 0.f);
      const double sz = a.sizes[(size_t)(obj_id - 1) * a.F + f];
      const float l32 = (float)((double)loc * sz);
      a.c3d[3 * dst + lane] = a.centers[((size_t)(obj_id - 1) * a.F + f) * 3 + lane] + (double)l32;
    }
  }
}

__global__ void __launch_bounds__(CT, 1) corresp_kernel(CorrArgs a) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ int scan_sh[CW + 1];
  __shared__ unsigned long long s_prefix, s_maskbits;
  __shared__ int s_need, s_count;
  const int seg = blockIdx.x, b = seg / a.J, j = seg % a.J, tid = threadIdx.x;
  const int obj_id = a.obj_ids[j];
  const int HW = a.h * a.w;
  unsigned int* mpix = a.ws + (size_t)seg * 4 * (HW + 1);
  float* vobj_l = reinterpret_cast<float*>(mpix + (HW + 1));
  float* rmax_l = reinterpret_cast<float*>(mpix + 2 * (HW + 1));
  unsigned int* off = mpix + 3 * (HW + 1);
  const size_t seg_base = (size_t)seg * a.cap;
  if (obj_id < 1 || obj_id > a.O) {
    if (tid == 0) { a.counts[seg] = 0; if (a.totals) a.totals[seg] = 0; }
    return;
  }
  // ---- phase 0: ordered list of masked pixels (obj_conf > tau_a, corresp.py:46-47); contiguous range per thread ----
  const float* oc = a.obj_conf + (size_t)b * HW * (a.O + 1) + obj_id;
  const int per = (HW + CT - 1) / CT;
  const int lo = min(tid * per, HW), hi = min(lo + per, HW);
  int nmask = 0;
  for (int i = lo; i < hi; ++i) nmask += (__ldg(oc + (size_t)i * (a.O + 1)) > a.min_obj_conf);
  int nm;
  {
    int base = block_scan_excl(nmask, scan_sh, &nm);
    for (int i = lo; i < hi; ++i) {
      const float v = __ldg(oc + (size_t)i * (a.O + 1));
      if (v > a.min_obj_conf) { mpix[base] = (unsigned int)i; vobj_l[base] = v; ++base; }
    }
  }
  __syncthreads();
  // ---- phase 1: row max and number of selected fragments per masked pixel, then exclusive scan ----
  {
    const float* fcb = a.frag_conf + ((size_t)b * HW * a.O + (obj_id - 1)) * a.F;
    const bool vec = (a.F % 4 == 0);
    for (int m = tid; m < nm; m += CT) {
      const float* row = fcb + (size_t)mpix[m] * a.O * a.F;
      float mx = -INFINITY;
      if (vec) {
        for (int f = 0; f < a.F; f += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(row + f));
          mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        }
      } else {
        for (int f = 0; f < a.F; ++f) mx = fmaxf(mx, __ldg(row + f));
      }
      const float thr = __fmul_rn(mx, a.min_rel);
      int c = 0;
      if (vec) {
        for (int f = 0; f < a.F; f += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(row + f));
          c += (v.x > thr) + (v.y > thr) + (v.z > thr) + (v.w > thr);
        }
      } else {
        for (int f = 0; f < a.F; ++f) c += (__ldg(row + f) > thr);
      }
      rmax_l[m] = mx;
      off[m] = (unsigned int)c;
    }
  }
  __syncthreads();
  int total;
  {
    const int per2 = (nm + CT - 1) / CT;
    const int l2 = min(tid * per2, nm), h2 = min(l2 + per2, nm);
    int sc = 0;
    for (int i = l2; i < h2; ++i) sc += (int)off[i];
    int base = block_scan_excl(sc, scan_sh, &total);
    for (int i = l2; i < h2; ++i) { const int c = (int)off[i]; off[i] = (unsigned int)base; base += c; }
    if (tid == 0) off[nm] = (unsigned int)total;
    __syncthreads();
  }
  SegLists L;
  L.mpix = mpix; L.vobj = vobj_l; L.rmax = rmax_l; L.off = off; L.nm = nm;
  if (tid == 0 && a.totals) a.totals[seg] = total;
  const bool topk = a.max_corr > 0 && total > a.max_corr;
  if (!topk) {
    const int n = total < a.cap ? total : a.cap;
    if (tid == 0) a.counts[seg] = n;
    for_each_candidate(a, b, obj_id, L, [&](int p, int f, unsigned int e, float vo, float vf) {
      if ((int)e < a.cap) write_row(a, b, obj_id, seg_base + e, p, f, vo, vf);
    });
    if (!a.frag_loc) { __syncthreads(); lazy_loc_rows(a, b, obj_id, seg_base, n); }
    return;
  }
  // ---- phase 2: radix select of the K largest 64-bit keys ----
  const int K = a.max_corr;
  unsigned int* hist = reinterpret_cast<unsigned int*>(smem_raw);            // 2048 bins
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + 2048 * 4);   // SORT_MAX
  unsigned int* payload = reinterpret_cast<unsigned int*>(smem_raw + 2048 * 4 + SORT_MAX * 8);
  if (tid == 0) { s_prefix = 0ULL; s_maskbits = 0ULL; s_need = K; }
  __syncthreads();
  // digits from the top: conf bits 31..21, 20..10, 9..0, then emission index bits 31..21 (always 0 for < 2^21 rows
  // but kept general), 20..10, 9..0
  const int shifts[6] = {53, 42, 32, 21, 10, 0};
  const int widths[6] = {11, 11, 10, 11, 11, 10};
  for (int d = 0; d < 6; ++d) {
    const int sh_ = shifts[d], wd = widths[d];
    if (d == 3 && s_need == s_count) break;     // all rows tied with the K-th confidence are kept: no index digits needed
    for (int i = tid; i < 2048; i += CT) hist[i] = 0u;
    __syncthreads();
    const unsigned long long pre = s_prefix, mk = s_maskbits;
    for_each_candidate(a, b, obj_id, L, [&](int, int, unsigned int e, float vo, float vf) {
      const unsigned long long key = make_key(vo, vf, e);
      if ((key & mk) == pre) atomicAdd(&hist[(unsigned int)(key >> sh_) & ((1u << wd) - 1u)], 1u);
    });
    __syncthreads();
    // suffix scan from the top bin: thread t owns bins 2047-4t .. 2044-4t
    int cb[4];
    int csum = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { cb[q] = (int)hist[2047 - 4 * tid - q]; csum += cb[q]; }
    int tot2;
    int above = block_scan_excl(csum, scan_sh, &tot2);            // rows in bins strictly above this thread's first bin
    const int need = s_need;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (above < need && need <= above + cb[q]) {
        s_prefix = pre | ((unsigned long long)(2047 - 4 * tid - q) << sh_);
        s_need = need - above;
      }
      above += cb[q];
    }
    if (tid == 0) s_maskbits = mk | (((1ULL << wd) - 1ULL) << sh_);
    __syncthreads();
    if (d == 2) {                               // number of rows whose confidence equals the K-th one
      const unsigned int b3 = (unsigned int)(s_prefix >> sh_) & ((1u << wd) - 1u);
      if (tid == 0) s_count = (int)hist[b3];
      __syncthreads();
    }
  }
  // s_prefix is now the K-th largest key; keys are unique (emission index), so exactly K keys are >= it
  const unsigned long long kth = s_prefix;
  if (tid == 0) s_count = 0;
  for (int i = tid; i < SORT_MAX; i += CT) { keys[i] = 0ULL; payload[i] = 0u; }
  __syncthreads();
  for_each_candidate(a, b, obj_id, L, [&](int p, int f, unsigned int e, float vo, float vf) {
    const unsigned long long key = make_key(vo, vf, e);
    if (key >= kth) {
      const int slot = atomicAdd(&s_count, 1);
      if (slot < SORT_MAX) { keys[slot] = key; payload[slot] = ((unsigned int)p << 9) | (unsigned int)f; }
    }
  });
  __syncthreads();
  // bitonic sort, descending, SORT_MAX elements (padding keys are 0 = smallest)
  for (int k = 2; k <= SORT_MAX; k <<= 1)
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int i = tid; i < SORT_MAX; i += CT) {
        const int l = i ^ jj;
        if (l > i) {
          const bool desc = (i & k) == 0;
          const unsigned long long ki = keys[i], kl = keys[l];
          if (desc ? (ki < kl) : (ki > kl)) {
            keys[i] = kl; keys[l] = ki;
            const unsigned int t = payload[i]; payload[i] = payload[l]; payload[l] = t;
          }
        }
      }
      __syncthreads();
    }
  const int n = K < a.cap ? K : a.cap;
  if (tid == 0) a.counts[seg] = n;
  for (int i = tid; i < n; i += CT) {
    const unsigned int pl = payload[i];
    const int p = (int)(pl >> 9), f = (int)(pl & 511u);
    const float vo = __ldg(a.obj_conf + ((size_t)b * HW + p) * (a.O + 1) + obj_id);
    const float vf = __ldg(a.frag_conf + (((size_t)b * HW + p) * a.O + (obj_id - 1)) * a.F + f);
    write_row(a, b, obj_id, seg_base + i, p, f, vo, vf);
  }
  if (!a.frag_loc) { __syncthreads(); lazy_loc_rows(a, b, obj_id, seg_base, n); }
}


// =====================================================================================================
// Grid-wide pipeline (default).  The single-CTA-per-segment kernel above leaves the GPU to a handful of CTAs: the heavy
// segments (tens of thousands of candidate rows, radix select over several passes) run on ONE SM each while the other
// 140 idle, and all J CTAs of an image re-scan obj_conf with an (O+1)-float stride.  Here every stage is spread over the
// whole grid and obj_conf is read once per image:
//   rows     thread per pixel: the pixel's obj_conf row once; for every object slot above tau_a the F-float frag_conf row
//            -> row max and number of selected fragments, dense per (segment, pixel)
//   scan     CTA per segment: exclusive scan of the counts = emission offsets (row-major pixel, then fragment), totals,
//            top-K decision
//   emit     thread per (segment, pixel): rows of the segments that keep everything
//   hist/pick (x6, top-K segments only) radix select of the max_corr largest 64-bit keys (conf bits << 32 | emission
//            index) with per-CTA shared-memory histograms merged into a global one; later digits stop early as before
//   gather   survivors (key >= K-th key) -> per-segment list;  sort   CTA per segment: bitonic sort, rows in descending order
//   loc      (lazy localisation head) warp per emitted row over all segments
// Results are bit-identical to the kernel above (same keys, same order).
// =====================================================================================================
struct SegState {
  unsigned long long prefix, maskbits;
  int need, tie_count, topk, done, sel_count, total, pad0, pad1;
};

struct Corr2 {
  CorrArgs a;
  unsigned int* cnt;        // [S][HW+1] selected fragments per (segment, pixel)
  unsigned int* off;        // [S][HW+1] exclusive scan; off[HW] = total
  float* rmax;              // [S][HW+1]
  unsigned int* hist;       // [S][2048]
  SegState* state;          // [S]
  unsigned long long* sel_keys;   // [S][SORT_MAX]
  unsigned int* sel_payload;      // [S][SORT_MAX]
};

__global__ void __launch_bounds__(128) corr2_rows_kernel(Corr2 c) {
  __shared__ int s_ids[64];
  const CorrArgs& a = c.a;
  const int HW = a.h * a.w, b = blockIdx.y;
  for (int j = threadIdx.x; j < a.J; j += 128) s_ids[j] = a.obj_ids[j];
  __syncthreads();
  const int p = blockIdx.x * 128 + threadIdx.x;
  if (p >= HW) return;
  const float* oc = a.obj_conf + ((size_t)b * HW + p) * (a.O + 1);
  const bool vec = (a.F % 4 == 0);
  for (int j = 0; j < a.J; ++j) {
    const int obj_id = s_ids[j];
    const size_t si = ((size_t)b * a.J + j) * (HW + 1) + p;
    unsigned int n = 0;
    if (obj_id >= 1 && obj_id <= a.O && __ldg(oc + obj_id) > a.min_obj_conf) {
      const float* row = a.frag_conf + (((size_t)b * HW + p) * a.O + (obj_id - 1)) * a.F;
      float mx = -INFINITY;
      if (vec) {
        for (int f = 0; f < a.F; f += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(row + f));
          mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        }
      } else {
        for (int f = 0; f < a.F; ++f) mx = fmaxf(mx, __ldg(row + f));
      }
      const float thr = __fmul_rn(mx, a.min_rel);
      if (vec) {
        for (int f = 0; f < a.F; f += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(row + f));
          n += (v.x > thr) + (v.y > thr) + (v.z > thr) + (v.w > thr);
        }
      } else {
        for (int f = 0; f < a.F; ++f) n += (__ldg(row + f) > thr);
      }
      c.rmax[si] = mx;
    }
    c.cnt[si] = n;
  }
}

__global__ void __launch_bounds__(CT, 1) corr2_scan_kernel(Corr2 c) {
  __shared__ int scan_sh[CW + 1];
  const CorrArgs& a = c.a;
  const int seg = blockIdx.x, tid = threadIdx.x, HW = a.h * a.w;
  const unsigned int* cnt = c.cnt + (size_t)seg * (HW + 1);
  unsigned int* off = c.off + (size_t)seg * (HW + 1);
  const int per = (HW + CT - 1) / CT;
  const int lo = min(tid * per, HW), hi = min(lo + per, HW);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += (int)cnt[i];
  int total;
  int base = block_scan_excl(s, scan_sh, &total);
  for (int i = lo; i < hi; ++i) { off[i] = (unsigned int)base; base += (int)cnt[i]; }
  for (int i = tid; i < 2048; i += CT) c.hist[(size_t)seg * 2048 + i] = 0u;
  if (tid == 0) {
    off[HW] = (unsigned int)total;
    const bool topk = a.max_corr > 0 && total > a.max_corr;
    const int n = topk ? (a.max_corr < a.cap ? a.max_corr : a.cap) : (total < a.cap ? total : a.cap);
    a.counts[seg] = n;
    if (a.totals) a.totals[seg] = total;
    SegState st;
    st.prefix = 0ULL; st.maskbits = 0ULL; st.need = a.max_corr; st.tie_count = 0; st.topk = topk ? 1 : 0; st.done = 0;
    st.sel_count = 0; st.total = total; st.pad0 = st.pad1 = 0;
    c.state[seg] = st;
  }
}

// visits the selected fragments of (segment, pixel): fn(f, e, vfrag)
template <class Fn>
__device__ __forceinline__ void corr2_candidates(const Corr2& c, int b, int obj_id, int seg, int p, Fn fn) {
  const CorrArgs& a = c.a;
  const int HW = a.h * a.w;
  const size_t si = (size_t)seg * (HW + 1) + p;
  if (c.cnt[si] == 0) return;
  unsigned int e = c.off[si];
  const float thr = __fmul_rn(c.rmax[si], a.min_rel);
  const float* row = a.frag_conf + (((size_t)b * HW + p) * a.O + (obj_id - 1)) * a.F;
  if (a.F % 4 == 0) {
    for (int f = 0; f < a.F; f += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(row + f));
      if (v.x > thr) fn(f, e++, v.x);
      if (v.y > thr) fn(f + 1, e++, v.y);
      if (v.z > thr) fn(f + 2, e++, v.z);
      if (v.w > thr) fn(f + 3, e++, v.w);
    }
  } else {
    for (int f = 0; f < a.F; ++f) {
      const float v = __ldg(row + f);
      if (v > thr) fn(f, e++, v);
    }
  }
}

__global__ void __launch_bounds__(128) corr2_emit_kernel(Corr2 c) {
  const CorrArgs& a = c.a;
  const int seg = blockIdx.y, HW = a.h * a.w;
  if (c.state[seg].topk) return;
  const int p = blockIdx.x * 128 + threadIdx.x;
  if (p >= HW) return;
  const int b = seg / a.J, obj_id = a.obj_ids[seg % a.J];
  if (obj_id < 1 || obj_id > a.O) return;
  const size_t seg_base = (size_t)seg * a.cap;
  float vobj = 0.f;
  bool have = false;
  corr2_candidates(c, b, obj_id, seg, p, [&](int f, unsigned int e, float vf) {
    if (!have) { vobj = __ldg(a.obj_conf + ((size_t)b * HW + p) * (a.O + 1) + obj_id); have = true; }
    if ((int)e < a.cap) write_row(a, b, obj_id, seg_base + e, p, f, vobj, vf);
  });
}

__constant__ int c2_shifts[6] = {53, 42, 32, 21, 10, 0};
__constant__ int c2_widths[6] = {11, 11, 10, 11, 11, 10};

__global__ void __launch_bounds__(256) corr2_hist_kernel(Corr2 c, int d) {
  __shared__ unsigned int sh[2048];
  const CorrArgs& a = c.a;
  const int seg = blockIdx.y, HW = a.h * a.w;
  const SegState st = c.state[seg];
  if (!st.topk || st.done) return;
  for (int i = threadIdx.x; i < 2048; i += 256) sh[i] = 0u;
  __syncthreads();
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int b = seg / a.J, obj_id = a.obj_ids[seg % a.J];
  const int sh_ = c2_shifts[d], wd = c2_widths[d];
  if (p < HW && obj_id >= 1 && obj_id <= a.O) {
    float vobj = 0.f;
    bool have = false;
    corr2_candidates(c, b, obj_id, seg, p, [&](int, unsigned int e, float vf) {
      if (!have) { vobj = __ldg(a.obj_conf + ((size_t)b * HW + p) * (a.O + 1) + obj_id); have = true; }
      const unsigned long long key = make_key(vobj, vf, e);
      if ((key & st.maskbits) == st.prefix) atomicAdd(&sh[(unsigned int)(key >> sh_) & ((1u << wd) - 1u)], 1u);
    });
  }
  __syncthreads();
  unsigned int* gh = c.hist + (size_t)seg * 2048;
  for (int i = threadIdx.x; i < 2048; i += 256)
    if (sh[i]) atomicAdd(&gh[i], sh[i]);
}

__global__ void __launch_bounds__(CT, 1) corr2_pick_kernel(Corr2 c, int d) {
  __shared__ int scan_sh[CW + 1];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need;
  const int seg = blockIdx.x, tid = threadIdx.x;
  SegState* stp = c.state + seg;
  const SegState st = *stp;
  if (!st.topk || st.done) return;
  unsigned int* hist = c.hist + (size_t)seg * 2048;
  const int sh_ = c2_shifts[d], wd = c2_widths[d];
  if (tid == 0) { s_prefix = st.prefix; s_need = st.need; }
  int cb[4];
  int csum = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) { cb[q] = (int)hist[2047 - 4 * tid - q]; csum += cb[q]; }
  int tot2;
  int above = block_scan_excl(csum, scan_sh, &tot2);
  const int need = st.need;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (above < need && need <= above + cb[q]) {
      s_prefix = st.prefix | ((unsigned long long)(2047 - 4 * tid - q) << sh_);
      s_need = need - above;
    }
    above += cb[q];
  }
  __syncthreads();
  if (tid == 0) {
    stp->prefix = s_prefix; stp->need = s_need;
    stp->maskbits = st.maskbits | (((1ULL << wd) - 1ULL) << sh_);
    if (d == 2) {
      // rows whose confidence equals the K-th one: if all of them are needed the index digits are not (ties are all kept)
      const int ties = (int)hist[(unsigned int)(s_prefix >> sh_) & ((1u << wd) - 1u)];
      stp->tie_count = ties;
      if (s_need == ties) stp->done = 1;
    }
    if (d == 5) stp->done = 1;
  }
  __syncthreads();
  for (int i = tid; i < 2048; i += CT) hist[i] = 0u;
}

__global__ void __launch_bounds__(256) corr2_gather_kernel(Corr2 c) {
  const CorrArgs& a = c.a;
  const int seg = blockIdx.y, HW = a.h * a.w;
  const SegState st = c.state[seg];
  if (!st.topk) return;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int b = seg / a.J, obj_id = a.obj_ids[seg % a.J];
  if (p >= HW || obj_id < 1 || obj_id > a.O) return;
  const unsigned long long kth = st.prefix;
  float vobj = 0.f;
  bool have = false;
  corr2_candidates(c, b, obj_id, seg, p, [&](int f, unsigned int e, float vf) {
    if (!have) { vobj = __ldg(a.obj_conf + ((size_t)b * HW + p) * (a.O + 1) + obj_id); have = true; }
    const unsigned long long key = make_key(vobj, vf, e);
    if (key >= kth) {
      const int slot = atomicAdd(&c.state[seg].sel_count, 1);
      if (slot < SORT_MAX) {
        c.sel_keys[(size_t)seg * SORT_MAX + slot] = key;
        c.sel_payload[(size_t)seg * SORT_MAX + slot] = ((unsigned int)p << 9) | (unsigned int)f;
      }
    }
  });
}

__global__ void __launch_bounds__(CT, 1) corr2_sort_kernel(Corr2 c) {
  extern __shared__ unsigned char smem_raw[];
  const CorrArgs& a = c.a;
  const int seg = blockIdx.x, tid = threadIdx.x, HW = a.h * a.w;
  const SegState st = c.state[seg];
  if (!st.topk) return;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  unsigned int* payload = reinterpret_cast<unsigned int*>(smem_raw + SORT_MAX * 8);
  const int nsel = st.sel_count < SORT_MAX ? st.sel_count : SORT_MAX;
  for (int i = tid; i < SORT_MAX; i += CT) {
    keys[i] = i < nsel ? c.sel_keys[(size_t)seg * SORT_MAX + i] : 0ULL;
    payload[i] = i < nsel ? c.sel_payload[(size_t)seg * SORT_MAX + i] : 0u;
  }
  __syncthreads();
  for (int k = 2; k <= SORT_MAX; k <<= 1)
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int i = tid; i < SORT_MAX; i += CT) {
        const int l = i ^ jj;
        if (l > i) {
          const bool desc = (i & k) == 0;
          const unsigned long long ki = keys[i], kl = keys[l];
          if (desc ? (ki < kl) : (ki > kl)) {
            keys[i] = kl; keys[l] = ki;
            const unsigned int t = payload[i]; payload[i] = payload[l]; payload[l] = t;
          }
        }
      }
      __syncthreads();
    }
  const int b = seg / a.J, obj_id = a.obj_ids[seg % a.J];
  const int n = a.counts[seg];
  const size_t seg_base = (size_t)seg * a.cap;
  for (int i = tid; i < n; i += CT) {
    const unsigned int pl = payload[i];
    const int p = (int)(pl >> 9), f = (int)(pl & 511u);
    const float vo = __ldg(a.obj_conf + ((size_t)b * HW + p) * (a.O + 1) + obj_id);
    const float vf = __ldg(a.frag_conf + (((size_t)b * HW + p) * a.O + (obj_id - 1)) * a.F + f);
    write_row(a, b, obj_id, seg_base + i, p, f, vo, vf);
  }
}

// lazy localisation head over the emitted rows of every segment: CW warps per CTA, grid (row chunks, segments)
__global__ void __launch_bounds__(CT) corr2_loc_kernel(Corr2 c, int rows_per_cta) {
  const CorrArgs& a = c.a;
  const int seg = blockIdx.y;
  const int n = a.counts[seg];
  const int r0 = blockIdx.x * rows_per_cta;
  if (r0 >= n) return;
  const int b = seg / a.J, obj_id = a.obj_ids[seg % a.J];
  const int r1 = min(n, r0 + rows_per_cta);
  // lazy_loc_rows walks rows [0, n) of a segment with the CTA's warps: give it the sub-range as its own "segment"
  lazy_loc_rows(a, b, obj_id, (size_t)seg * a.cap + r0, r1 - r0);
}

static size_t corr2_extra_bytes(int S) { return (size_t)S * (2048 * 4 + sizeof(SegState) + SORT_MAX * 12) + 1024; }

static int corresp2_launch(const CorrArgs& a, void* ws_aligned, cudaStream_t stream) {
  const int S = a.B * a.J, HW = a.h * a.w;
  Corr2 c;
  c.a = a;
  unsigned char* p = reinterpret_cast<unsigned char*>(ws_aligned);
  c.cnt = reinterpret_cast<unsigned int*>(p); p += (size_t)S * (HW + 1) * 4;
  c.off = reinterpret_cast<unsigned int*>(p); p += (size_t)S * (HW + 1) * 4;
  c.rmax = reinterpret_cast<float*>(p); p += (size_t)S * (HW + 1) * 4;
  p += (size_t)S * (HW + 1) * 4;                                   // (fourth plane of the v1 layout: unused)
  p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 255) & ~(uintptr_t)255);
  c.hist = reinterpret_cast<unsigned int*>(p); p += (size_t)S * 2048 * 4;
  c.sel_keys = reinterpret_cast<unsigned long long*>(p); p += (size_t)S * SORT_MAX * 8;
  c.sel_payload = reinterpret_cast<unsigned int*>(p); p += (size_t)S * SORT_MAX * 4;
  c.state = reinterpret_cast<SegState*>(p);
  EPOS_CHECK_ARG(a.J <= 64 && S <= 65535);
  corr2_rows_kernel<<<dim3(ceil_div(HW, 128), a.B), 128, 0, stream>>>(c);
  EPOS_LAUNCH_CHECK();
  corr2_scan_kernel<<<S, CT, 0, stream>>>(c);
  EPOS_LAUNCH_CHECK();
  corr2_emit_kernel<<<dim3(ceil_div(HW, 128), S), 128, 0, stream>>>(c);
  EPOS_LAUNCH_CHECK();
  if (a.max_corr > 0) {
    for (int d = 0; d < 6; ++d) {
      corr2_hist_kernel<<<dim3(ceil_div(HW, 256), S), 256, 0, stream>>>(c, d);
      EPOS_LAUNCH_CHECK();
      corr2_pick_kernel<<<S, CT, 0, stream>>>(c, d);
      EPOS_LAUNCH_CHECK();
    }
    corr2_gather_kernel<<<dim3(ceil_div(HW, 256), S), 256, 0, stream>>>(c);
    EPOS_LAUNCH_CHECK();
    const size_t smem = (size_t)SORT_MAX * 12;
    static std::atomic<int> attr[EPOS_MAX_DEVICES];
    const int dslot = device_slot();
    if (!attr[dslot].load(std::memory_order_acquire)) {
      EPOS_CUDA(cudaFuncSetAttribute(corr2_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr[dslot].store(1, std::memory_order_release);
    }
    corr2_sort_kernel<<<S, CT, smem, stream>>>(c);
    EPOS_LAUNCH_CHECK();
  }
  if (!a.frag_loc) {
    const int rows_per_cta = 4 * CW;
    corr2_loc_kernel<<<dim3(ceil_div(a.cap, rows_per_cta), S), CT, 0, stream>>>(c, rows_per_cta);
    EPOS_LAUNCH_CHECK();
  }
  return EPOS_OK;
}

}  // namespace epos

using namespace epos;

extern "C" {

size_t epos_corresp_workspace_bytes(int B, int J, int h, int w) {
  if (B <= 0 || J <= 0 || h <= 0 || w <= 0) return 0;
  return (size_t)B * J * ((size_t)h * w + 1) * 16 + 512 + corr2_extra_bytes(B * J);
}

static int corresp_launch(const float* obj_conf, const float* frag_conf, const float* frag_loc, const uint16_t* feat, int ldf,
                          size_t feat_plane, int feat_c, const float* w_loc, const float* b_loc, int B, int h, int w,
                          int num_objs, int num_frags, const int32_t* obj_ids, int J, const double* frag_centers,
                          const double* frag_sizes, double output_scale, float min_obj_conf, float min_frag_rel_conf,
                          int cap, int max_corr, double* coord_2d, double* coord_3d, float* conf, float* conf_obj,
                          float* conf_frag, int32_t* px, int32_t* frag, int32_t* counts, int32_t* totals, void* workspace,
                          size_t workspace_bytes, void* stream) {
  EPOS_CHECK_ARG(obj_conf && frag_conf && obj_ids && frag_centers && frag_sizes);
  EPOS_CHECK_ARG(frag_loc || (feat && w_loc && feat_c > 0 && (feat_c % 4) == 0 && ldf >= feat_c && (ldf % 4) == 0 &&
                              (feat_plane % 4) == 0 && (reinterpret_cast<uintptr_t>(feat) & 7) == 0 &&
                              (reinterpret_cast<uintptr_t>(w_loc) & 15) == 0));
  EPOS_CHECK_ARG(coord_2d && coord_3d && conf && conf_obj && conf_frag && px && frag && counts && workspace);
  EPOS_CHECK_ARG(B > 0 && h > 0 && w > 0 && num_objs > 0 && num_frags > 0 && J > 0 && cap > 0 && output_scale > 0);
  EPOS_CHECK_ARG((size_t)h * w < (1u << 23));
  if (num_frags > MAXF) { set_error("epos_corresp: num_frags=%d > %d unsupported", num_frags, MAXF); return EPOS_ERR_UNSUPPORTED; }
  if (max_corr > SORT_MAX) { set_error("epos_corresp: max_corr=%d > %d unsupported", max_corr, SORT_MAX); return EPOS_ERR_UNSUPPORTED; }
  if (workspace_bytes < epos_corresp_workspace_bytes(B, J, h, w)) { set_error("epos_corresp: workspace too small"); return EPOS_ERR_INVALID_ARG; }
  CorrArgs a;
  a.obj_conf = obj_conf; a.frag_conf = frag_conf; a.frag_loc = frag_loc;
  a.B = B; a.h = h; a.w = w; a.O = num_objs; a.F = num_frags; a.obj_ids = obj_ids; a.J = J;
  a.centers = frag_centers; a.sizes = frag_sizes; a.inv_scale = 1.0 / output_scale;
  a.min_obj_conf = min_obj_conf; a.min_rel = min_frag_rel_conf; a.cap = cap; a.max_corr = max_corr;
  a.c2d = coord_2d; a.c3d = coord_3d; a.conf = conf; a.conf_obj = conf_obj; a.conf_frag = conf_frag; a.px = px; a.frag = frag;
  a.counts = counts; a.totals = totals;
  a.feat = feat; a.ldf = ldf; a.feat_plane = (long long)feat_plane; a.w_loc = w_loc; a.b_loc = b_loc; a.feat_c = feat_c;
  a.ws = reinterpret_cast<unsigned int*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  // grid-wide pipeline by default; EPOS_CORRESP_V1=1 selects the one-CTA-per-segment kernel (developer A/B)
  static int v1 = -1;
  if (v1 < 0) { const char* e = getenv("EPOS_CORRESP_V1"); v1 = e ? atoi(e) : 0; }
  if (!v1) return corresp2_launch(a, a.ws, (cudaStream_t)stream);
  const size_t smem = 2048 * 4 + (size_t)SORT_MAX * 12;
  static std::atomic<int> attr[EPOS_MAX_DEVICES];
  const int dslot = device_slot();
  if (!attr[dslot].load(std::memory_order_acquire)) {
    EPOS_CUDA(cudaFuncSetAttribute(corresp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[dslot].store(1, std::memory_order_release);
  }
  corresp_kernel<<<B * J, CT, smem, (cudaStream_t)stream>>>(a);
  EPOS_LAUNCH_CHECK();
  return EPOS_OK;
}

int epos_corresp(const float* obj_conf, const float* frag_conf, const float* frag_loc, int B, int h, int w, int num_objs,
                 int num_frags, const int32_t* obj_ids, int J, const double* frag_centers, const double* frag_sizes,
                 double output_scale, float min_obj_conf, float min_frag_rel_conf, int cap, int max_corr, double* coord_2d,
                 double* coord_3d, float* conf, float* conf_obj, float* conf_frag, int32_t* px, int32_t* frag,
                 int32_t* counts, int32_t* totals, void* workspace, size_t workspace_bytes, void* stream) {
  EPOS_CHECK_ARG(frag_loc);
  return corresp_launch(obj_conf, frag_conf, frag_loc, nullptr, 0, 0, 0, nullptr, nullptr, B, h, w, num_objs, num_frags,
                        obj_ids, J, frag_centers, frag_sizes, output_scale, min_obj_conf, min_frag_rel_conf, cap, max_corr,
                        coord_2d, coord_3d, conf, conf_obj, conf_frag, px, frag, counts, totals, workspace, workspace_bytes,
                        stream);
}

int epos_corresp_lazy_loc(const float* obj_conf, const float* frag_conf, const uint16_t* feat_split, int ldf,
                          size_t feat_plane_stride, int feat_channels, const float* w_loc, const float* b_loc, int B, int h,
                          int w, int num_objs, int num_frags, const int32_t* obj_ids, int J, const double* frag_centers,
                          const double* frag_sizes, double output_scale, float min_obj_conf, float min_frag_rel_conf,
                          int cap, int max_corr, double* coord_2d, double* coord_3d, float* conf, float* conf_obj,
                          float* conf_frag, int32_t* px, int32_t* frag, int32_t* counts, int32_t* totals, void* workspace,
                          size_t workspace_bytes, void* stream) {
  EPOS_CHECK_ARG(feat_split && w_loc);
  return corresp_launch(obj_conf, frag_conf, nullptr, feat_split, ldf, feat_plane_stride, feat_channels, w_loc, b_loc, B, h, w,
                        num_objs, num_frags, obj_ids, J, frag_centers, frag_sizes, output_scale, min_obj_conf,
                        min_frag_rel_conf, cap, max_corr, coord_2d, coord_3d, conf, conf_obj, conf_frag, px, frag, counts,
                        totals, workspace, workspace_bytes, stream);
}

}  // extern "C"
