// A3 sphere collider, A4 spaced sampler, A5/A6/A7 NeuS up-sampling round (fixed-inv_s alphas ->
// cumprod weights -> pdf/cdf -> searchsorted(right) -> inverse-cdf lerp -> sorted merge).
// ref: src/model_components/scene_colliders.py:60-113, src/model_components/ray_samplers.py:38-68,
//      183-296,316-422,516-551, src/cameras/rays.py:201-217
//
// Bit-exactness: everything that is plain IEEE add/sub/mul/div on the CPU is written with the
// round-to-nearest intrinsics (no FMA contraction) in the reference's association order; cumsum /
// cumprod accumulate sequentially in double like ATen's CPU kernels (verified against torch 2.11);
// the only ulp-level differences left are expf inside the sigmoid and torch.sum's vector-lane order.
#include "common.cuh"
#include <stdlib.h>

namespace mmsb {

__global__ void __launch_bounds__(256) sphere_collide_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                                             float radius, float* __restrict__ nears,
                                                             float* __restrict__ fars, uint8_t* __restrict__ mask,
                                                             float* __restrict__ bg_nears, float* __restrict__ bg_fars,
                                                             int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float ox = o[3 * i], oy = o[3 * i + 1], oz = o[3 * i + 2];
  const float dx = d[3 * i], dy = d[3 * i + 1], dz = d[3 * i + 2];
  // ref: scene_colliders.py:62-63
  const float b = __fadd_rn(__fadd_rn(__fmul_rn(dx, ox), __fmul_rn(dy, oy)), __fmul_rn(dz, oz));
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy)), __fmul_rn(oz, oz)));
  float under = __fsub_rn(__fmul_rn(b, b), __fsub_rn(__fmul_rn(nrm, nrm), __fmul_rn(radius, radius)));
  const bool hit = under > 0.01f;
  under = fmaxf(under, 0.01f);
  const float sq = sqrtf(under);
  const float nr = fmaxf(__fsub_rn(-sq, b), 0.01f);
  const float fr = fmaxf(__fsub_rn(sq, b), 0.01f);
  if (nears) nears[i] = nr;
  if (fars) fars[i] = fr;
  if (mask) mask[i] = hit ? 1 : 0;
  if (bg_nears) bg_nears[i] = hit ? fr : nr;            // scene_colliders.py:112
  if (bg_fars) bg_fars[i] = __fadd_rn(fr, 3.0f);        // scene_colliders.py:113
}

__device__ __forceinline__ float spacing_to_euclid(float x, float nr, float fr, int spacing) {
  // ref: ray_samplers.py:178-181
  if (spacing == MMSB_SPACING_DISPARITY) {
    const float sn = __fdiv_rn(1.f, nr), sf = __fdiv_rn(1.f, fr);
    return __fdiv_rn(1.f, __fadd_rn(__fmul_rn(sf, x), __fmul_rn(sn, __fsub_rn(1.f, x))));
  }
  return __fadd_rn(__fmul_rn(fr, x), __fmul_rn(nr, __fsub_rn(1.f, x)));
}

// one thread per (ray, bin edge)
__global__ void __launch_bounds__(256) spaced_bins_kernel(const float* __restrict__ nears, const float* __restrict__ fars,
                                                          const float* __restrict__ lin, const float* __restrict__ t_rand,
                                                          int rand_per_ray, int ns, int spacing,
                                                          float* __restrict__ sbins, float* __restrict__ ebins, int64_t n) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int nb = ns + 1;
  if (t >= n * nb) return;
  const int64_t r = t / nb;
  const int j = int(t - r * nb);
  float b = __ldg(lin + j);
  if (t_rand) {
    // ref: ray_samplers.py:212-220
    const float tr = rand_per_ray == 1 ? __ldg(t_rand + r) : __ldg(t_rand + r * nb + j);
    const float upper = j < ns ? __fdiv_rn(__fadd_rn(__ldg(lin + j + 1), b), 2.f) : __ldg(lin + ns);
    const float lower = j > 0 ? __fdiv_rn(__fadd_rn(b, __ldg(lin + j - 1)), 2.f) : __ldg(lin);
    b = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), tr));
  }
  sbins[t] = b;
  ebins[t] = spacing_to_euclid(b, __ldg(nears + r), __ldg(fars + r), spacing);
}

// index of the first element of row[0..m) that is > v  (torch.searchsorted side="right")
__device__ __forceinline__ int upper_bound(const float* __restrict__ row, int m, float v) {
  int lo = 0, hi = m;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (!(row[mid] > v)) lo = mid + 1; else hi = mid;   // NaN in v: comparisons false -> lo = m (as ATen)
  }
  return lo;
}

__global__ void __launch_bounds__(256) searchsorted_kernel(const float* __restrict__ cdf, const float* __restrict__ u,
                                                           int64_t* __restrict__ inds, int m, int q, int64_t n) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * q) return;
  const int64_t r = t / q;
  inds[t] = upper_bound(cdf + r * m, m, __ldg(u + t));
}

// inverse-cdf lerp of one stratified sample; ref: ray_samplers.py:394-403
__device__ __forceinline__ float pdf_inverse_one(const float* __restrict__ cdf, const float* __restrict__ bins, int nb,
                                                 float uv, int* ind_out) {
  const int ind = upper_bound(cdf, nb, uv);
  if (ind_out) *ind_out = ind;
  const int below = min(max(ind - 1, 0), nb - 1), above = min(max(ind, 0), nb - 1);
  const float c0 = cdf[below], c1 = cdf[above], b0 = bins[below], b1 = bins[above];
  float t = __fdiv_rn(__fsub_rn(uv, c0), __fsub_rn(c1, c0));
  if (isnan(t)) t = 0.f;                       // nan_to_num(nan=0); +-inf are clipped below
  t = fminf(fmaxf(t, 0.f), 1.f);
  return __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
}

__global__ void __launch_bounds__(256) pdf_inverse_kernel(const float* __restrict__ cdf, const float* __restrict__ bins,
                                                          const float* __restrict__ u, int nb, int q,
                                                          int64_t* __restrict__ inds, float* __restrict__ out, int64_t n) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * q) return;
  const int64_t r = t / q;
  int ind;
  out[t] = pdf_inverse_one(cdf + r * nb, bins + r * nb, nb, __ldg(u + t), &ind);
  if (inds) inds[t] = ind;
}

// One thread per ray (the scans are sequential in the reference and m <= a few hundred).
// cdf_ws: [n, m+1] workspace, holds the cdf on return.
__global__ void __launch_bounds__(128) neus_upsample_kernel(const float* __restrict__ bins, const float* __restrict__ sdf,
                                                            const float* __restrict__ u, const float* __restrict__ nears,
                                                            const float* __restrict__ fars, float inv_s, float hist_pad,
                                                            float eps, int m, int k, float* __restrict__ cdf_ws,
                                                            int64_t* __restrict__ inds_out, float* __restrict__ new_bins,
                                                            float* __restrict__ merged_bins,
                                                            int64_t* __restrict__ merged_index, int64_t n) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float* b = bins + r * (m + 1);
  const float* sd = sdf + r * m;
  float* cdf = cdf_ws + r * (m + 1);
  const float nr = nears[r], fr = fars[r];

  // --- alphas with fixed inv_s + transmittance weights (ray_samplers.py:516-551, rays.py:201-217)
  double T = 1.0;          // ATen's CPU cumprod accumulates in double
  double wsum = 0.0;
  float prev_cos = 0.f;
  float e_prev = spacing_to_euclid(b[0], nr, fr, MMSB_SPACING_UNIFORM);
  for (int j = 0; j < m - 1; ++j) {
    const float e_next = spacing_to_euclid(b[j + 1], nr, fr, MMSB_SPACING_UNIFORM);
    const float delta = __fsub_rn(e_next, e_prev);
    e_prev = e_next;
    const float ps = sd[j], ns = sd[j + 1];
    const float mid = __fmul_rn(__fadd_rn(ps, ns), 0.5f);
    const float cosv = __fdiv_rn(__fsub_rn(ns, ps), __fadd_rn(delta, 1e-5f));
    float c = fminf(prev_cos, cosv);
    prev_cos = cosv;
    c = fminf(fmaxf(c, -1e3f), 0.f);
    const float half = __fmul_rn(__fmul_rn(c, delta), 0.5f);
    const float pc = sigmoidf_(__fmul_rn(__fsub_rn(mid, half), inv_s));
    const float nc = sigmoidf_(__fmul_rn(__fadd_rn(mid, half), inv_s));
    const float alpha = __fdiv_rn(__fadd_rn(__fsub_rn(pc, nc), 1e-5f), __fadd_rn(pc, 1e-5f));
    const float w = __fmul_rn(alpha, float(T));
    T *= double(__fadd_rn(__fsub_rn(1.0f, alpha), 1e-7f));
    const float wp = __fadd_rn(w, hist_pad);   // ray_samplers.py:353
    cdf[j + 1] = wp;
    wsum += double(wp);
  }
  {  // the zero weight appended at ray_samplers.py:498
    const float wp = __fadd_rn(0.f, hist_pad);
    cdf[m] = wp;
    wsum += double(wp);
  }
  // --- pdf / cdf (ray_samplers.py:356-363)
  float ws = float(wsum);
  const float padding = fmaxf(__fsub_rn(eps, ws), 0.f);
  const float padw = __fdiv_rn(padding, float(m));
  ws = __fadd_rn(ws, padding);
  double acc = 0.0;        // ATen's CPU cumsum accumulates in double
  cdf[0] = 0.f;
  for (int j = 1; j <= m; ++j) {
    const float pdf = __fdiv_rn(__fadd_rn(cdf[j], padw), ws);
    acc += double(pdf);
    cdf[j] = fminf(1.f, float(acc));
  }
  // --- stratified inverse-cdf samples (ray_samplers.py:386-403)
  float* nb = new_bins + r * (k + 1);
  for (int q = 0; q <= k; ++q) {
    int ind;
    nb[q] = pdf_inverse_one(cdf, b, m + 1, __ldg(u + r * (k + 1) + q), &ind);
    if (inds_out) inds_out[r * (k + 1) + q] = ind;
  }
  // --- sorted merge of the bin starts, old first on ties (ray_samplers.py:46-53)
  float* mb = merged_bins + r * (m + k + 1);
  int64_t* mi = merged_index ? merged_index + r * (m + k) : nullptr;
  int ia = 0, ib = 0;
  for (int o = 0; o < m + k; ++o) {
    const bool take_a = ib >= k || (ia < m && !(nb[ib] < b[ia]));
    if (take_a) { mb[o] = b[ia]; if (mi) mi[o] = ia; ++ia; }
    else        { mb[o] = nb[ib]; if (mi) mi[o] = m + ib; ++ib; }
  }
  mb[m + k] = fmaxf(b[m], nb[k]);
}

// ---- the same round, one WARP per ray ----------------------------------------------------------------------------
// Lane l owns the elements l, l + 32, ... of the ray; a row's bins / cdf / new bins live in shared memory (coalesced
// global loads and stores).  cumprod and cumsum are warp scans in double with a carry between 32-element chunks: the
// reference's sequential double accumulation re-associated, i.e. equal before the cast back to float except for
// double-rounding ties (probability ~2^-29 per value).  searchsorted: one lane per query, binary search in shared
// memory.  Sorted merge: every element's output slot is its rank (old first on ties, ray_samplers.py:46-53).
__device__ __forceinline__ double warp_scan_prod_f64(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
__device__ __forceinline__ double warp_scan_sum_f64(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// number of elements of the sorted row[0..m) that are < v
__device__ __forceinline__ int lower_bound(const float* __restrict__ row, int m, float v) {
  int lo = 0, hi = m;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (row[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

constexpr int kUpsampleWarps = 8;

__global__ void __launch_bounds__(kUpsampleWarps * 32) neus_upsample_warp_kernel(
    const float* __restrict__ bins, const float* __restrict__ sdf, const float* __restrict__ u,
    const float* __restrict__ nears, const float* __restrict__ fars, float inv_s, float hist_pad, float eps, int m, int k,
    float* __restrict__ cdf_ws, int64_t* __restrict__ inds_out, float* __restrict__ new_bins,
    float* __restrict__ merged_bins, int64_t* __restrict__ merged_index, int64_t n) {
  extern __shared__ float sm_up[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = int64_t(blockIdx.x) * kUpsampleWarps + warp;
  if (r >= n) return;                                   // warp-uniform; no block-level synchronisation below
  const int ne = m + 1;
  float* s_b = sm_up + warp * (2 * ne + k + 1);         // bin edges [ne]
  float* s_c = s_b + ne;                                // weights, then the cdf [ne]
  float* s_nb = s_c + ne;                               // new bins [k + 1]
  const float* b = bins + r * ne;
  const float* sd = sdf + r * m;
  for (int j = lane; j < ne; j += 32) s_b[j] = b[j];
  __syncwarp();
  const float nr = nears[r], fr = fars[r];

  // --- alphas with fixed inv_s + transmittance weights (ray_samplers.py:516-551, rays.py:201-217)
  double carry_t = 1.0, wsum = 0.0;
  float carry_cos = 0.f;
  for (int base = 0; base < m; base += 32) {
    const int j = base + lane;
    const bool valid = j < m - 1;
    float cosv = 0.f, delta = 0.f, mid = 0.f;
    if (valid) {
      delta = __fsub_rn(spacing_to_euclid(s_b[j + 1], nr, fr, MMSB_SPACING_UNIFORM),
                        spacing_to_euclid(s_b[j], nr, fr, MMSB_SPACING_UNIFORM));
      const float ps = sd[j], ns = sd[j + 1];
      mid = __fmul_rn(__fadd_rn(ps, ns), 0.5f);
      cosv = __fdiv_rn(__fsub_rn(ns, ps), __fadd_rn(delta, 1e-5f));
    }
    float prev_cos = __shfl_up_sync(0xffffffffu, cosv, 1);
    if (lane == 0) prev_cos = carry_cos;
    carry_cos = __shfl_sync(0xffffffffu, cosv, 31);
    float alpha = 0.f;
    if (valid) {
      float c = fminf(prev_cos, cosv);
      c = fminf(fmaxf(c, -1e3f), 0.f);
      const float half = __fmul_rn(__fmul_rn(c, delta), 0.5f);
      const float pc = sigmoidf_(__fmul_rn(__fsub_rn(mid, half), inv_s));
      const float nc = sigmoidf_(__fmul_rn(__fadd_rn(mid, half), inv_s));
      alpha = __fdiv_rn(__fadd_rn(__fsub_rn(pc, nc), 1e-5f), __fadd_rn(pc, 1e-5f));
    }
    const double f = valid ? double(__fadd_rn(__fsub_rn(1.0f, alpha), 1e-7f)) : 1.0;
    const double incl = warp_scan_prod_f64(f, lane);
    double excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.0;
    const double t_j = carry_t * excl;                  // T before sample j
    carry_t = carry_t * __shfl_sync(0xffffffffu, incl, 31);
    if (j < m) {
      // j == m - 1: the zero weight appended at ray_samplers.py:498
      const float w = valid ? __fmul_rn(alpha, float(t_j)) : 0.f;
      const float wp = __fadd_rn(w, hist_pad);          // ray_samplers.py:353
      s_c[j + 1] = wp;
      wsum += double(wp);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  // --- pdf / cdf (ray_samplers.py:356-363)
  float ws = float(wsum);
  const float padding = fmaxf(__fsub_rn(eps, ws), 0.f);
  const float padw = __fdiv_rn(padding, float(m));
  ws = __fadd_rn(ws, padding);
  double carry = 0.0;
  if (lane == 0) s_c[0] = 0.f;
  for (int base = 0; base < m; base += 32) {
    const int j = base + lane;
    const double pdf = j < m ? double(__fdiv_rn(__fadd_rn(s_c[j + 1], padw), ws)) : 0.0;
    const double incl = carry + warp_scan_sum_f64(pdf, lane);
    carry = __shfl_sync(0xffffffffu, incl, 31);
    if (j < m) s_c[j + 1] = fminf(1.f, float(incl));
  }
  __syncwarp();
  float* cdf = cdf_ws + r * ne;
  for (int j = lane; j < ne; j += 32) cdf[j] = s_c[j];
  // --- stratified inverse-cdf samples (ray_samplers.py:386-403), one lane per query
  for (int q = lane; q <= k; q += 32) {
    int ind;
    const float v = pdf_inverse_one(s_c, s_b, ne, __ldg(u + r * (k + 1) + q), &ind);
    s_nb[q] = v;
    new_bins[r * (k + 1) + q] = v;
    if (inds_out) inds_out[r * (k + 1) + q] = ind;
  }
  __syncwarp();
  // --- sorted merge of the bin starts by rank, old first on ties (ray_samplers.py:46-53)
  float* mb = merged_bins + r * (m + k + 1);
  int64_t* mi = merged_index ? merged_index + r * (m + k) : nullptr;
  for (int ia = lane; ia < m; ia += 32) {
    const float v = s_b[ia];
    const int pos = ia + lower_bound(s_nb, k, v);       // new starts strictly below v come first
    mb[pos] = v;
    if (mi) mi[pos] = ia;
  }
  for (int ib = lane; ib < k; ib += 32) {
    const float v = s_nb[ib];
    const int pos = ib + upper_bound(s_b, m, v);        // old starts <= v come first
    mb[pos] = v;
    if (mi) mi[pos] = m + ib;
  }
  if (lane == 0) mb[m + k] = fmaxf(s_b[m], s_nb[k]);
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ a, int m, const float* __restrict__ b2,
                                                          int k, const int64_t* __restrict__ index,
                                                          float* __restrict__ out, int64_t n) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int w = m + k;
  if (t >= n * w) return;
  const int64_t r = t / w;
  const int64_t j = index[t];
  out[t] = j < m ? a[r * m + j] : b2[r * k + (j - m)];
}

}  // namespace mmsb

using namespace mmsb;

extern "C" int mmsb_sphere_collide(const float* origins, const float* directions, float radius, float* nears,
                                   float* fars, uint8_t* mask, float* bg_nears, float* bg_fars, int64_t n,
                                   mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && radius > 0.f, "sphere_collide: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(origins && directions, "sphere_collide: NULL pointer");
  sphere_collide_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(origins, directions, radius, nears,
                                                                                   fars, mask, bg_nears, bg_fars, n);
  return check_launch("sphere_collide");
}

extern "C" int mmsb_spaced_bins(const float* nears, const float* fars, const float* lin, const float* t_rand,
                                int32_t rand_per_ray, int32_t num_samples, int32_t spacing, float* spacing_bins,
                                float* euclid_bins, int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && num_samples >= 1, "spaced_bins: bad sizes");
  MMSB_REQUIRE(spacing == MMSB_SPACING_UNIFORM || spacing == MMSB_SPACING_DISPARITY, "spaced_bins: unknown spacing %d",
               spacing);
  MMSB_REQUIRE(!t_rand || rand_per_ray == 1 || rand_per_ray == num_samples + 1,
               "spaced_bins: rand_per_ray must be 1 or num_samples+1");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(nears && fars && lin && spacing_bins && euclid_bins, "spaced_bins: NULL pointer");
  const int64_t total = n * (num_samples + 1);
  spaced_bins_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
      nears, fars, lin, t_rand, rand_per_ray, num_samples, spacing, spacing_bins, euclid_bins, n);
  return check_launch("spaced_bins");
}

extern "C" int mmsb_searchsorted_right(const float* cdf, const float* u, int64_t* inds, int32_t m, int32_t q, int64_t n,
                                       mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && m >= 0 && q >= 1, "searchsorted_right: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(cdf && u && inds, "searchsorted_right: NULL pointer");
  searchsorted_kernel<<<(unsigned)ceil_div(n * q, 256), 256, 0, as_stream(stream)>>>(cdf, u, inds, m, q, n);
  return check_launch("searchsorted_right");
}

extern "C" int mmsb_pdf_inverse(const float* cdf, const float* bins, const float* u, int32_t num_edges, int32_t q,
                                int64_t* inds, float* new_bins, int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && num_edges >= 1 && q >= 1, "pdf_inverse: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(cdf && bins && u && new_bins, "pdf_inverse: NULL pointer");
  pdf_inverse_kernel<<<(unsigned)ceil_div(n * q, 256), 256, 0, as_stream(stream)>>>(cdf, bins, u, num_edges, q, inds,
                                                                                    new_bins, n);
  return check_launch("pdf_inverse");
}

extern "C" int mmsb_neus_upsample(const float* bins, const float* sdf, const float* u, const float* nears,
                                  const float* fars, float inv_s, float histogram_padding, float eps, int32_t m,
                                  int32_t k, float* cdf_ws, int64_t* inds_out, float* new_bins, float* merged_bins,
                                  int64_t* merged_index, int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && m >= 2 && k >= 1, "neus_upsample: bad sizes m=%d k=%d", m, k);
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(bins && sdf && u && nears && fars && cdf_ws && new_bins && merged_bins, "neus_upsample: NULL pointer");
  // one warp per ray while a row's bins, cdf and new bins fit the block's shared memory (m up to ~700 samples)
  const size_t smem = size_t(kUpsampleWarps) * (2 * (m + 1) + k + 1) * sizeof(float);
  static int force_serial = -1;
  if (force_serial < 0) { const char* e = getenv("MMSB_UPSAMPLE_SERIAL"); force_serial = e ? atoi(e) : 0; }
  if (smem <= 48 * 1024 && !force_serial) {
    neus_upsample_warp_kernel<<<(unsigned)ceil_div(n, kUpsampleWarps), kUpsampleWarps * 32, smem, as_stream(stream)>>>(
        bins, sdf, u, nears, fars, inv_s, histogram_padding, eps, m, k, cdf_ws, inds_out, new_bins, merged_bins,
        merged_index, n);
    return check_launch("neus_upsample");
  }
  neus_upsample_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, as_stream(stream)>>>(
      bins, sdf, u, nears, fars, inv_s, histogram_padding, eps, m, k, cdf_ws, inds_out, new_bins, merged_bins,
      merged_index, n);
  return check_launch("neus_upsample");
}

extern "C" int mmsb_merge_rows(const float* a, int32_t m, const float* b, int32_t k, const int64_t* index, float* out,
                               int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && m >= 0 && k >= 0 && m + k >= 1, "merge_rows: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(a && b && index && out, "merge_rows: NULL pointer");
  gather_rows_kernel<<<(unsigned)ceil_div(n * (m + k), 256), 256, 0, as_stream(stream)>>>(a, m, b, k, index, out, n);
  return check_launch("merge_rows");
}
