// A13 tap gradients, A14 NeuS alphas + transmittance weights, A18 density weights, A19 compositing.
// One warp per ray; samples are strided over the lanes and the exclusive transmittance product /
// suffix sums are warp scans with a carry between 32-sample chunks.
// ref: src/model_components/surface_model.py:137-153,98; src/model_components/volume_rendering.py:177-213;
//      src/cameras/rays.py:138-151,201-217; src/model_components/renderers.py:76-243
#include "common.cuh"

namespace mmsb {

constexpr int kWarpsPerBlock = 8;

struct NeusSample {
  float alpha, x;        // x = 1 - alpha + 1e-7
  float pc, nc, prev, next, ic, tc, craw;
};

__device__ __forceinline__ NeusSample neus_alpha(float sdf, float gx, float gy, float gz, float dx, float dy, float dz,
                                                 float delta, float s, float anneal) {
  NeusSample o;
  o.tc = dx * gx + dy * gy + dz * gz;
  o.ic = -(fmaxf(-o.tc * 0.5f + 0.5f, 0.f) * (1.f - anneal) + fmaxf(-o.tc, 0.f) * anneal);
  const float half = o.ic * delta * 0.5f;
  o.next = sdf + half;
  o.prev = sdf - half;
  o.pc = sigmoidf_(o.prev * s);
  o.nc = sigmoidf_(o.next * s);
  o.craw = (o.pc - o.nc + 1e-5f) / (o.pc + 1e-5f);
  o.alpha = fminf(fmaxf(o.craw, 0.f), 1.f);
  o.x = 1.f - o.alpha + 1e-7f;
  return o;
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock) neus_weights_fwd_kernel(
    const float* __restrict__ sdf, const float* __restrict__ grad, const float* __restrict__ dirs,
    const float* __restrict__ deltas, const float* __restrict__ inv_s, const uint8_t* __restrict__ mask, float anneal,
    float* __restrict__ weights, int s, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t r = int64_t(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= n) return;
  if (mask && !mask[r]) {
    for (int j = lane; j < s; j += 32) weights[r * s + j] = 0.f;
    return;
  }
  const float sv = __ldg(inv_s);
  const float dx = dirs[3 * r], dy = dirs[3 * r + 1], dz = dirs[3 * r + 2];
  float carry = 1.f;
  for (int j0 = 0; j0 < s; j0 += 32) {
    const int j = j0 + lane;
    float alpha = 0.f, x = 1.f;
    if (j < s) {
      const int64_t q = r * s + j;
      const NeusSample a = neus_alpha(sdf[q], grad[3 * q], grad[3 * q + 1], grad[3 * q + 2], dx, dy, dz, deltas[q], sv, anneal);
      alpha = a.alpha; x = a.x;
    }
    const float inc = warp_scan_prod(x, lane);
    float excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = 1.f;
    const float T = carry * excl;
    if (j < s) weights[r * s + j] = alpha * T;
    carry *= __shfl_sync(0xffffffffu, inc, 31);
  }
}

// Backward: d alpha_j = g_j T_j - (sum_{i>j} g_i w_i) / x_j.  Chunks are walked from the far end so the
// suffix sum is a carried reverse scan; T_j is recomputed with a forward pre-pass per chunk.
__global__ void __launch_bounds__(32 * kWarpsPerBlock) neus_weights_bwd_kernel(
    const float* __restrict__ sdf, const float* __restrict__ grad, const float* __restrict__ dirs,
    const float* __restrict__ deltas, const float* __restrict__ inv_s, const uint8_t* __restrict__ mask, float anneal,
    const float* __restrict__ d_weights, float* __restrict__ d_sdf, float* __restrict__ d_grad,
    float* __restrict__ d_dirs, float* __restrict__ d_deltas, float* __restrict__ d_inv_s, int s, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t r = int64_t(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= n) return;
  if (mask && !mask[r]) {
    for (int j = lane; j < s; j += 32) {
      const int64_t q = r * s + j;
      d_sdf[q] = 0.f;
      d_grad[3 * q] = d_grad[3 * q + 1] = d_grad[3 * q + 2] = 0.f;
      if (d_deltas) d_deltas[q] = 0.f;
    }
    if (d_dirs && lane < 3) d_dirs[3 * r + lane] = 0.f;
    return;
  }
  const float sv = __ldg(inv_s);
  const float dx = dirs[3 * r], dy = dirs[3 * r + 1], dz = dirs[3 * r + 2];
  const int nchunks = (s + 31) / 32;
  // total transmittance products at chunk starts: pass 1 (forward) computes the carry of every chunk
  // on the fly when walking backwards we need carry[c]; s <= 1024 -> at most 32 chunks, kept in a lane.
  float chunk_carry = 1.f;   // lane c holds carry entering chunk c
  {
    float carry = 1.f;
    for (int c = 0; c < nchunks; ++c) {
      const int j = c * 32 + lane;
      float x = 1.f;
      if (j < s) {
        const int64_t q = r * s + j;
        x = neus_alpha(sdf[q], grad[3 * q], grad[3 * q + 1], grad[3 * q + 2], dx, dy, dz, deltas[q], sv, anneal).x;
      }
      const float inc = warp_scan_prod(x, lane);
      if (lane == c) chunk_carry = carry;
      carry *= __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  float suffix = 0.f;        // sum_{i in later chunks} g_i w_i
  float ddx = 0.f, ddy = 0.f, ddz = 0.f, ds = 0.f;
  for (int c = nchunks - 1; c >= 0; --c) {
    const int j = c * 32 + lane;
    const int64_t q = r * s + j;
    const bool ok = j < s;
    NeusSample a;
    float gxs = 0.f, gys = 0.f, gzs = 0.f, del = 0.f, g = 0.f;
    if (ok) {
      gxs = grad[3 * q]; gys = grad[3 * q + 1]; gzs = grad[3 * q + 2]; del = deltas[q];
      a = neus_alpha(sdf[q], gxs, gys, gzs, dx, dy, dz, del, sv, anneal);
      g = d_weights[q];
    } else {
      a.alpha = 0.f; a.x = 1.f; a.pc = a.nc = 0.5f; a.prev = a.next = a.ic = a.tc = 0.f; a.craw = 0.f;
    }
    const float inc = warp_scan_prod(a.x, lane);
    float excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = 1.f;
    const float T = __shfl_sync(0xffffffffu, chunk_carry, c) * excl;
    const float gw = g * a.alpha * T;
    // exclusive suffix sum within the chunk: total - inclusive prefix
    const float pre = warp_scan_sum(gw, lane);
    const float tot = __shfl_sync(0xffffffffu, pre, 31);
    const float S = suffix + (tot - pre);
    suffix += tot;
    if (!ok) continue;
    float dalpha = g * T - S / a.x;
    if (!(a.craw >= 0.f && a.craw <= 1.f)) dalpha = 0.f;
    const float den = a.pc + 1e-5f;
    const float dpc = dalpha * (1.f / den - (a.pc - a.nc + 1e-5f) / (den * den));
    const float dnc = -dalpha / den;
    const float dps = dpc * a.pc * (1.f - a.pc);   // d / d(prev * s)
    const float dns = dnc * a.nc * (1.f - a.nc);   // d / d(next * s)
    ds += dps * a.prev + dns * a.next;
    const float dprev = dps * sv, dnext = dns * sv;
    d_sdf[q] = dprev + dnext;
    const float dhalf = dnext - dprev;
    const float dic = dhalf * del * 0.5f;
    if (d_deltas) d_deltas[q] = dhalf * a.ic * 0.5f;
    const float u1 = -a.tc * 0.5f + 0.5f, u2 = -a.tc;
    const float dtc = dic * ((u1 > 0.f ? 0.5f * (1.f - anneal) : 0.f) + (u2 > 0.f ? anneal : 0.f));
    d_grad[3 * q] = dtc * dx; d_grad[3 * q + 1] = dtc * dy; d_grad[3 * q + 2] = dtc * dz;
    ddx += dtc * gxs; ddy += dtc * gys; ddz += dtc * gzs;
  }
  ds = warp_sum(ds);
  if (d_dirs) {
    ddx = warp_sum(ddx); ddy = warp_sum(ddy); ddz = warp_sum(ddz);
    if (lane == 0) { d_dirs[3 * r] = ddx; d_dirs[3 * r + 1] = ddy; d_dirs[3 * r + 2] = ddz; }
  }
  if (d_inv_s && lane == 0) atomicAdd(d_inv_s, ds);
}

// ---- density -> alpha -> weights (background field) ----------------------------------------
__global__ void __launch_bounds__(32 * kWarpsPerBlock) density_weights_fwd_kernel(const float* __restrict__ density,
                                                                                  const float* __restrict__ deltas,
                                                                                  float* __restrict__ weights, int s,
                                                                                  int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t r = int64_t(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= n) return;
  float carry = 1.f;
  for (int j0 = 0; j0 < s; j0 += 32) {
    const int j = j0 + lane;
    float alpha = 0.f, x = 1.f;
    if (j < s) {
      alpha = 1.f - expf(-(deltas[r * s + j] * density[r * s + j]));   // rays.py:148-149
      x = 1.f - alpha + 1e-7f;                                          // rays.py:211-213
    }
    const float inc = warp_scan_prod(x, lane);
    float excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = 1.f;
    if (j < s) weights[r * s + j] = alpha * carry * excl;
    carry *= __shfl_sync(0xffffffffu, inc, 31);
  }
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock) density_weights_bwd_kernel(
    const float* __restrict__ density, const float* __restrict__ deltas, const float* __restrict__ d_weights,
    float* __restrict__ d_density, float* __restrict__ d_deltas, int s, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t r = int64_t(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= n) return;
  const int nchunks = (s + 31) / 32;
  float chunk_carry = 1.f;
  {
    float carry = 1.f;
    for (int c = 0; c < nchunks; ++c) {
      const int j = c * 32 + lane;
      float x = 1.f;
      if (j < s) x = expf(-(deltas[r * s + j] * density[r * s + j])) + 1e-7f;
      const float inc = warp_scan_prod(x, lane);
      if (lane == c) chunk_carry = carry;
      carry *= __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  float suffix = 0.f;
  for (int c = nchunks - 1; c >= 0; --c) {
    const int j = c * 32 + lane;
    const bool ok = j < s;
    float e = 1.f, del = 0.f, den = 0.f, g = 0.f;
    if (ok) { del = deltas[r * s + j]; den = density[r * s + j]; e = expf(-(del * den)); g = d_weights[r * s + j]; }
    const float alpha = ok ? 1.f - e : 0.f;
    const float x = ok ? 1.f - alpha + 1e-7f : 1.f;
    const float inc = warp_scan_prod(x, lane);
    float excl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) excl = 1.f;
    const float T = __shfl_sync(0xffffffffu, chunk_carry, c) * excl;
    const float gw = g * alpha * T;
    const float pre = warp_scan_sum(gw, lane);
    const float tot = __shfl_sync(0xffffffffu, pre, 31);
    const float S = suffix + (tot - pre);
    suffix += tot;
    if (!ok) continue;
    const float dalpha = g * T - S / x;
    const float dprod = dalpha * e;          // d alpha / d(delta * density) = exp(-delta density)
    d_density[r * s + j] = dprod * del;
    if (d_deltas) d_deltas[r * s + j] = dprod * den;
  }
}

// ---- compositing -----------------------------------------------------------------------------
constexpr int kMaxChannels = 16;

__global__ void __launch_bounds__(32 * kWarpsPerBlock) composite_fwd_kernel(
    const float* __restrict__ weights, const float* __restrict__ values, const float* __restrict__ background, int c,
    const float* __restrict__ normals, const float* __restrict__ starts, const float* __restrict__ ends,
    float* __restrict__ out_color, float* __restrict__ out_normals, float* __restrict__ out_depth,
    float* __restrict__ out_acc, int s, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t r = int64_t(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= n) return;
  float col[kMaxChannels];
#pragma unroll
  for (int k = 0; k < kMaxChannels; ++k) col[k] = 0.f;
  float acc = 0.f, dep = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
  for (int j = lane; j < s; j += 32) {
    const int64_t q = r * s + j;
    const float w = weights[q];
    acc += w;
    if (values) {
#pragma unroll
      for (int k = 0; k < kMaxChannels; ++k)
        if (k < c) col[k] = fmaf(w, values[q * c + k], col[k]);
    }
    if (out_depth) dep = fmaf(w, (starts[q] + ends[q]) * 0.5f, dep);
    if (out_normals) {
      nx = fmaf(w, normals[3 * q], nx); ny = fmaf(w, normals[3 * q + 1], ny); nz = fmaf(w, normals[3 * q + 2], nz);
    }
  }
  acc = warp_sum(acc);
  if (out_color) {
#pragma unroll
    for (int k = 0; k < kMaxChannels; ++k)
      if (k < c) {
        const float v = warp_sum(col[k]);
        if (lane == 0) out_color[r * c + k] = v + (background ? background[r * c + k] * (1.f - acc) : 0.f);
      }
  }
  if (out_depth) { dep = warp_sum(dep); if (lane == 0) out_depth[r] = dep; }
  if (out_normals) {
    nx = warp_sum(nx); ny = warp_sum(ny); nz = warp_sum(nz);
    if (lane == 0) { out_normals[3 * r] = nx; out_normals[3 * r + 1] = ny; out_normals[3 * r + 2] = nz; }
  }
  if (out_acc && lane == 0) out_acc[r] = acc;
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock) composite_bwd_kernel(
    const float* __restrict__ weights, const float* __restrict__ values, const float* __restrict__ background, int c,
    const float* __restrict__ d_color, const float* __restrict__ normals, const float* __restrict__ d_out_normals,
    const float* __restrict__ starts, const float* __restrict__ ends, const float* __restrict__ d_out_depth,
    const float* __restrict__ d_out_acc, float* __restrict__ d_weights, float* __restrict__ d_values,
    float* __restrict__ d_background, float* __restrict__ d_normals, int s, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t r = int64_t(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= n) return;
  float dc[kMaxChannels], bg[kMaxChannels];
#pragma unroll
  for (int k = 0; k < kMaxChannels; ++k) {
    dc[k] = (k < c && d_color) ? d_color[r * c + k] : 0.f;
    bg[k] = (k < c && background) ? background[r * c + k] : 0.f;
  }
  const float dacc = d_out_acc ? d_out_acc[r] : 0.f;
  const float ddep = d_out_depth ? d_out_depth[r] : 0.f;
  float dnx = 0.f, dny = 0.f, dnz = 0.f;
  if (d_out_normals) { dnx = d_out_normals[3 * r]; dny = d_out_normals[3 * r + 1]; dnz = d_out_normals[3 * r + 2]; }
  float acc = 0.f;
  for (int j = lane; j < s; j += 32) {
    const int64_t q = r * s + j;
    const float w = weights[q];
    acc += w;
    float dw = dacc;
#pragma unroll
    for (int k = 0; k < kMaxChannels; ++k)
      if (k < c) {
        const float v = values ? values[q * c + k] : 0.f;
        dw = fmaf(dc[k], v - bg[k], dw);
        if (d_values) d_values[q * c + k] = dc[k] * w;
      }
    if (d_out_depth) dw = fmaf(ddep, (starts[q] + ends[q]) * 0.5f, dw);
    if (d_out_normals) {
      dw += dnx * normals[3 * q] + dny * normals[3 * q + 1] + dnz * normals[3 * q + 2];
      if (d_normals) { d_normals[3 * q] = dnx * w; d_normals[3 * q + 1] = dny * w; d_normals[3 * q + 2] = dnz * w; }
    }
    d_weights[q] = dw;
  }
  if (d_background) {
    acc = warp_sum(acc);
    if (lane < c) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < kMaxChannels; ++k)
        if (k == lane) v = dc[k];
      d_background[r * c + lane] = v * (1.f - acc);
    }
  }
}

// ---- numerical gradient taps ----------------------------------------------------------------
// ref: surface_model.py:139-151.  four_delta = float(4.0 * delta), delta_sq = float(delta ** 2) are
// evaluated by the caller in double exactly like the reference's Python scalars.
__global__ void __launch_bounds__(256) sdf_taps_fwd_kernel(const float* __restrict__ sc, const float* __restrict__ st,
                                                           float four_delta, float delta_sq, float* __restrict__ gradients,
                                                           float* __restrict__ hessians, float* __restrict__ normals,
                                                           int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s1 = st[i], s2 = st[n + i], s3 = st[2 * n + i], s4 = st[3 * n + i];
  // k1=(1,-1,-1) k2=(-1,-1,1) k3=(-1,1,-1) k4=(1,1,1); ((k1 s1 + k2 s2) + k3 s3) + k4 s4
  const float gx = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(s1, -s2), -s3), s4), four_delta);
  const float gy = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(-s1, -s2), s3), s4), four_delta);
  const float gz = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(-s1, s2), -s3), s4), four_delta);
  gradients[3 * i] = gx; gradients[3 * i + 1] = gy; gradients[3 * i + 2] = gz;
  if (hessians) {
    const float sum = __fadd_rn(__fadd_rn(__fadd_rn(s1, s2), s3), s4);
    const float hxx = __fdiv_rn(__fsub_rn(__fdiv_rn(sum, 2.f), __fmul_rn(2.f, sc[i])), delta_sq);
    const float h = __fdiv_rn(hxx, 3.f);
    hessians[3 * i] = h; hessians[3 * i + 1] = h; hessians[3 * i + 2] = h;
  }
  if (normals) {
    const float nrm = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-12f);   // F.normalize eps
    normals[3 * i] = gx / nrm; normals[3 * i + 1] = gy / nrm; normals[3 * i + 2] = gz / nrm;
  }
}

__global__ void __launch_bounds__(256) sdf_taps_bwd_kernel(const float* __restrict__ st, float four_delta, float delta_sq,
                                                           const float* __restrict__ dg, const float* __restrict__ dh,
                                                           const float* __restrict__ dn, float* __restrict__ d_sc,
                                                           float* __restrict__ d_st, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float ax = dg ? dg[3 * i] : 0.f, ay = dg ? dg[3 * i + 1] : 0.f, az = dg ? dg[3 * i + 2] : 0.f;
  if (dn) {
    const float s1 = st[i], s2 = st[n + i], s3 = st[2 * n + i], s4 = st[3 * n + i];
    const float gx = (s1 - s2 - s3 + s4) / four_delta, gy = (-s1 - s2 + s3 + s4) / four_delta,
                gz = (-s1 + s2 - s3 + s4) / four_delta;
    const float nr = sqrtf(gx * gx + gy * gy + gz * gz);
    if (nr > 1e-12f) {
      const float nx = gx / nr, ny = gy / nr, nz = gz / nr;
      const float dot = nx * dn[3 * i] + ny * dn[3 * i + 1] + nz * dn[3 * i + 2];
      ax += (dn[3 * i] - nx * dot) / nr; ay += (dn[3 * i + 1] - ny * dot) / nr; az += (dn[3 * i + 2] - nz * dot) / nr;
    } else {
      ax += dn[3 * i] / 1e-12f; ay += dn[3 * i + 1] / 1e-12f; az += dn[3 * i + 2] / 1e-12f;
    }
  }
  ax /= four_delta; ay /= four_delta; az /= four_delta;
  float hs = 0.f;   // d / d hessian_xx
  if (dh) hs = (dh[3 * i] + dh[3 * i + 1] + dh[3 * i + 2]) / 3.f / delta_sq;
  const float ht = hs * 0.5f;
  d_st[i] = ax - ay - az + ht;
  d_st[n + i] = -ax - ay + az + ht;
  d_st[2 * n + i] = -ax + ay - az + ht;
  d_st[3 * n + i] = ax + ay + az + ht;
  if (d_sc) d_sc[i] = -2.f * hs;
}

}  // namespace mmsb

using namespace mmsb;

#define WARP_GRID(n) (unsigned)ceil_div((n), kWarpsPerBlock), 32 * kWarpsPerBlock, 0, as_stream(stream)

extern "C" int mmsb_neus_weights_fwd(const float* sdf, const float* grad, const float* dirs, const float* deltas,
                                     const float* inv_s, const uint8_t* mask, float anneal, float* weights, int32_t s,
                                     int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && s >= 1 && s <= 1024, "neus_weights_fwd: bad sizes s=%d", s);
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(sdf && grad && dirs && deltas && inv_s && weights, "neus_weights_fwd: NULL pointer");
  neus_weights_fwd_kernel<<<WARP_GRID(n)>>>(sdf, grad, dirs, deltas, inv_s, mask, anneal, weights, s, n);
  return check_launch("neus_weights_fwd");
}

extern "C" int mmsb_neus_weights_bwd(const float* sdf, const float* grad, const float* dirs, const float* deltas,
                                     const float* inv_s, const uint8_t* mask, float anneal, const float* d_weights,
                                     float* d_sdf, float* d_grad, float* d_dirs, float* d_deltas, float* d_inv_s,
                                     int32_t s, int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && s >= 1 && s <= 1024, "neus_weights_bwd: bad sizes s=%d", s);
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(sdf && grad && dirs && deltas && inv_s && d_weights && d_sdf && d_grad, "neus_weights_bwd: NULL pointer");
  neus_weights_bwd_kernel<<<WARP_GRID(n)>>>(sdf, grad, dirs, deltas, inv_s, mask, anneal, d_weights, d_sdf, d_grad,
                                            d_dirs, d_deltas, d_inv_s, s, n);
  return check_launch("neus_weights_bwd");
}

extern "C" int mmsb_density_weights_fwd(const float* density, const float* deltas, float* weights, int32_t s, int64_t n,
                                        mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && s >= 1 && s <= 1024, "density_weights_fwd: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(density && deltas && weights, "density_weights_fwd: NULL pointer");
  density_weights_fwd_kernel<<<WARP_GRID(n)>>>(density, deltas, weights, s, n);
  return check_launch("density_weights_fwd");
}

extern "C" int mmsb_density_weights_bwd(const float* density, const float* deltas, const float* d_weights,
                                        float* d_density, float* d_deltas, int32_t s, int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && s >= 1 && s <= 1024, "density_weights_bwd: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(density && deltas && d_weights && d_density, "density_weights_bwd: NULL pointer");
  density_weights_bwd_kernel<<<WARP_GRID(n)>>>(density, deltas, d_weights, d_density, d_deltas, s, n);
  return check_launch("density_weights_bwd");
}

extern "C" int mmsb_composite_fwd(const float* weights, const float* values, const float* background, int32_t c,
                                  const float* normals, const float* starts, const float* ends, float* out_color,
                                  float* out_normals, float* out_depth, float* out_acc, int32_t s, int64_t n,
                                  mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && s >= 1 && c >= 0 && c <= kMaxChannels, "composite_fwd: bad sizes s=%d c=%d", s, c);
  MMSB_REQUIRE(!out_color || values, "composite_fwd: out_color needs values");
  MMSB_REQUIRE(!out_normals || normals, "composite_fwd: out_normals needs normals");
  MMSB_REQUIRE(!out_depth || (starts && ends), "composite_fwd: out_depth needs starts/ends");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(weights != nullptr, "composite_fwd: NULL pointer");
  composite_fwd_kernel<<<WARP_GRID(n)>>>(weights, values, background, c, normals, starts, ends, out_color, out_normals,
                                         out_depth, out_acc, s, n);
  return check_launch("composite_fwd");
}

extern "C" int mmsb_composite_bwd(const float* weights, const float* values, const float* background, int32_t c,
                                  const float* d_color, const float* normals, const float* d_out_normals,
                                  const float* starts, const float* ends, const float* d_out_depth,
                                  const float* d_out_acc, float* d_weights, float* d_values, float* d_background,
                                  float* d_normals, int32_t s, int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && s >= 1 && c >= 0 && c <= kMaxChannels, "composite_bwd: bad sizes s=%d c=%d", s, c);
  MMSB_REQUIRE(!d_out_normals || normals, "composite_bwd: d_out_normals needs normals");
  MMSB_REQUIRE(!d_out_depth || (starts && ends), "composite_bwd: d_out_depth needs starts/ends");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(weights && d_weights, "composite_bwd: NULL pointer");
  composite_bwd_kernel<<<WARP_GRID(n)>>>(weights, values, background, c, d_color, normals, d_out_normals, starts, ends,
                                         d_out_depth, d_out_acc, d_weights, d_values, d_background, d_normals, s, n);
  return check_launch("composite_bwd");
}

extern "C" int mmsb_sdf_taps_fwd(const float* sdf_c, const float* sdf_t, float four_delta, float delta_sq,
                                 float* gradients, float* hessians, float* normals, int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && four_delta > 0.f && delta_sq > 0.f, "sdf_taps_fwd: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(sdf_t && gradients && (!hessians || sdf_c), "sdf_taps_fwd: NULL pointer");
  sdf_taps_fwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(sdf_c, sdf_t, four_delta, delta_sq,
                                                                                 gradients, hessians, normals, n);
  return check_launch("sdf_taps_fwd");
}

extern "C" int mmsb_sdf_taps_bwd(const float* sdf_t, float four_delta, float delta_sq, const float* d_gradients,
                                 const float* d_hessians, const float* d_normals, float* d_sdf_c, float* d_sdf_t,
                                 int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && four_delta > 0.f && delta_sq > 0.f, "sdf_taps_bwd: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(sdf_t && d_sdf_t, "sdf_taps_bwd: NULL pointer");
  sdf_taps_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(sdf_t, four_delta, delta_sq, d_gradients,
                                                                                 d_hessians, d_normals, d_sdf_c, d_sdf_t, n);
  return check_launch("sdf_taps_bwd");
}
