// A9/A10 — multi-resolution hash-grid encoding, forward / backward (scatter-add + input gradient).
// Follows the arithmetic of the reference's torch path (src/field_components/encodings.py:244-304)
// operation by operation (IEEE mul/add without FMA contraction, same association order) so the
// features are reproducible bit for bit against a CPU evaluation of the same formula; the FeatureGrid
// rescale and coarse-to-fine mask (src/field_components/feature_structures.py:78-88) are fused.
//
// Mapping: one thread per (point, level); blockIdx.y = level so that a block's gathers stay inside
// one level's 2^log2 x F slice of the table (4 MiB at the shipped config: L2-resident, 126 MB L2).
#include "common.cuh"

namespace mmsb {

struct HashGridParams {
  int L, log2T, interp;
  float radius, inv_2r_dummy;
  float res[MMSB_MAX_LEVELS];
};

__device__ __forceinline__ uint32_t hash3(int cx, int cy, int cz, uint32_t tmask) {
  // ref: encodings.py:256-259 — int64 multiply / xor / python-style mod 2^k == uint32 wrap + mask
  return (uint32_t(cx) ^ (uint32_t(cy) * 2654435761u) ^ (uint32_t(cz) * 805459861u)) & tmask;
}

template <int F>
struct Feat {
  float v[F];
};

template <int F>
__device__ __forceinline__ Feat<F> load_feat(const float* __restrict__ table, uint32_t row) {
  Feat<F> f;
  if constexpr (F == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(table) + row);
    f.v[0] = t.x; f.v[1] = t.y;
  } else if constexpr (F == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(table) + row);
    f.v[0] = t.x; f.v[1] = t.y; f.v[2] = t.z; f.v[3] = t.w;
  } else if constexpr (F == 8) {
    float4 t0 = __ldg(reinterpret_cast<const float4*>(table) + 2 * size_t(row));
    float4 t1 = __ldg(reinterpret_cast<const float4*>(table) + 2 * size_t(row) + 1);
    f.v[0] = t0.x; f.v[1] = t0.y; f.v[2] = t0.z; f.v[3] = t0.w;
    f.v[4] = t1.x; f.v[5] = t1.y; f.v[6] = t1.z; f.v[7] = t1.w;
  } else {
#pragma unroll
    for (int i = 0; i < F; ++i) f.v[i] = __ldg(table + size_t(row) * F + i);
  }
  return f;
}

// a*o + b*(1-o) exactly as torch evaluates `f_hi * o + f_lo * (1 - o)` (three roundings + one).
__device__ __forceinline__ float lerp_ref(float hi, float lo, float o, float om) {
  return __fadd_rn(__fmul_rn(hi, o), __fmul_rn(lo, om));
}

struct Corner {
  uint32_t h[8];
  float o[3];   // interpolation weights of the "ceil" corner per axis (offset or smoothstep(offset))
  float d[3];   // d(weight)/d(offset) (1 for linear)
  float res;
};

__device__ __forceinline__ Corner corners(const HashGridParams& p, int level, float x0, float x1, float x2) {
  Corner c;
  const float res = p.res[level];
  c.res = res;
  float xs[3] = {x0, x1, x2};
  int cc[3], cf[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float v = xs[a];
    if (p.radius > 0.f) v = __fdiv_rn(__fadd_rn(v, p.radius), __fmul_rn(2.0f, p.radius));
    const float scaled = __fmul_rn(v, res);
    const float fl = floorf(scaled), ce = ceilf(scaled);
    cc[a] = int(ce);
    cf[a] = int(fl);
    float off = __fsub_rn(scaled, fl);
    if (p.interp == MMSB_INTERP_SMOOTHSTEP) {
      c.d[a] = 6.f * off * (1.f - off);
      off = off * off * (3.f - 2.f * off);
    } else {
      c.d[a] = 1.f;
    }
    c.o[a] = off;
  }
  const uint32_t tmask = (1u << p.log2T) - 1u;
  const uint32_t base = uint32_t(level) << p.log2T;
  // reference order hashed_0..7 (encodings.py:274-281): c=ceil, f=floor per (x,y,z)
  c.h[0] = base + hash3(cc[0], cc[1], cc[2], tmask);
  c.h[1] = base + hash3(cc[0], cf[1], cc[2], tmask);
  c.h[2] = base + hash3(cf[0], cf[1], cc[2], tmask);
  c.h[3] = base + hash3(cf[0], cc[1], cc[2], tmask);
  c.h[4] = base + hash3(cc[0], cc[1], cf[2], tmask);
  c.h[5] = base + hash3(cc[0], cf[1], cf[2], tmask);
  c.h[6] = base + hash3(cf[0], cf[1], cf[2], tmask);
  c.h[7] = base + hash3(cf[0], cc[1], cf[2], tmask);
  return c;
}

template <int F>
__global__ void __launch_bounds__(256) hashgrid_fwd_kernel(HashGridParams p, const float* __restrict__ x,
                                                           int64_t ldx, const float* __restrict__ table,
                                                           const float* __restrict__ mask,
                                                           float* __restrict__ out, int64_t ld_out,
                                                           int64_t* __restrict__ idx_out, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int level = blockIdx.y;
  if (i >= n) return;
  const float* xi = x + i * ldx;
  const Corner c = corners(p, level, __ldg(xi), __ldg(xi + 1), __ldg(xi + 2));
  Feat<F> f[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = load_feat<F>(table, c.h[k]);
  if (idx_out) {
    int64_t* io = idx_out + (i * p.L + level) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) io[k] = int64_t(c.h[k]);
  }
  const float ox = c.o[0], oy = c.o[1], oz = c.o[2];
  const float mx = __fsub_rn(1.f, ox), my = __fsub_rn(1.f, oy), mz = __fsub_rn(1.f, oz);
  float r[F];
#pragma unroll
  for (int j = 0; j < F; ++j) {
    // ref: encodings.py:292-302
    const float f03 = lerp_ref(f[0].v[j], f[3].v[j], ox, mx);
    const float f12 = lerp_ref(f[1].v[j], f[2].v[j], ox, mx);
    const float f56 = lerp_ref(f[5].v[j], f[6].v[j], ox, mx);
    const float f47 = lerp_ref(f[4].v[j], f[7].v[j], ox, mx);
    const float f0312 = lerp_ref(f03, f12, oy, my);
    const float f4756 = lerp_ref(f47, f56, oy, my);
    float v = lerp_ref(f0312, f4756, oz, mz);
    if (mask) v = __fmul_rn(v, __ldg(mask + level * F + j));
    r[j] = v;
  }
  float* o = out + i * ld_out + level * F;
  if constexpr (F == 2) {
    if ((reinterpret_cast<uintptr_t>(o) & 7) == 0) {
      *reinterpret_cast<float2*>(o) = make_float2(r[0], r[1]);
      return;
    }
  }
#pragma unroll
  for (int j = 0; j < F; ++j) o[j] = r[j];
}

template <int F>
__device__ __forceinline__ void atomic_add_feat(float* __restrict__ dtable, uint32_t row, const float* g, float w) {
  if constexpr (F == 2) {
    atomicAdd(reinterpret_cast<float2*>(dtable) + row, make_float2(g[0] * w, g[1] * w));
  } else if constexpr (F == 4) {
    atomicAdd(reinterpret_cast<float4*>(dtable) + row, make_float4(g[0] * w, g[1] * w, g[2] * w, g[3] * w));
  } else {
#pragma unroll
    for (int j = 0; j < F; ++j) atomicAdd(dtable + size_t(row) * F + j, g[j] * w);
  }
}

template <int F>
__global__ void __launch_bounds__(256) hashgrid_bwd_kernel(HashGridParams p, const float* __restrict__ x,
                                                           int64_t ldx, const float* __restrict__ table,
                                                           const float* __restrict__ mask,
                                                           const float* __restrict__ dout, int64_t ld_dout,
                                                           float* __restrict__ dtable, float* __restrict__ dx,
                                                           int64_t lddx, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int level = blockIdx.y;
  if (i >= n) return;
  const float* xi = x + i * ldx;
  const Corner c = corners(p, level, __ldg(xi), __ldg(xi + 1), __ldg(xi + 2));
  float g[F];
  bool any = false;
#pragma unroll
  for (int j = 0; j < F; ++j) {
    float v = __ldg(dout + i * ld_dout + level * F + j);
    if (mask) v *= __ldg(mask + level * F + j);
    g[j] = v;
    any |= (v != 0.f);
  }
  if (!any) return;  // masked (coarse-to-fine) levels contribute exact zeros in the reference too
  const float ox = c.o[0], oy = c.o[1], oz = c.o[2];
  const float mx = 1.f - ox, my = 1.f - oy, mz = 1.f - oz;
  if (dtable) {
    // weight of corner k = product of its per-axis factors (o for a ceil coordinate, 1-o for floor)
    atomic_add_feat<F>(dtable, c.h[0], g, ox * oy * oz);
    atomic_add_feat<F>(dtable, c.h[1], g, ox * my * oz);
    atomic_add_feat<F>(dtable, c.h[2], g, mx * my * oz);
    atomic_add_feat<F>(dtable, c.h[3], g, mx * oy * oz);
    atomic_add_feat<F>(dtable, c.h[4], g, ox * oy * mz);
    atomic_add_feat<F>(dtable, c.h[5], g, ox * my * mz);
    atomic_add_feat<F>(dtable, c.h[6], g, mx * my * mz);
    atomic_add_feat<F>(dtable, c.h[7], g, mx * oy * mz);
  }
  if (dx) {
    Feat<F> f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = load_feat<F>(table, c.h[k]);
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      const float f03 = f[0].v[j] * ox + f[3].v[j] * mx, d03 = f[0].v[j] - f[3].v[j];
      const float f12 = f[1].v[j] * ox + f[2].v[j] * mx, d12 = f[1].v[j] - f[2].v[j];
      const float f56 = f[5].v[j] * ox + f[6].v[j] * mx, d56 = f[5].v[j] - f[6].v[j];
      const float f47 = f[4].v[j] * ox + f[7].v[j] * mx, d47 = f[4].v[j] - f[7].v[j];
      const float f0312 = f03 * oy + f12 * my, f4756 = f47 * oy + f56 * my;
      gx += g[j] * ((d03 * oy + d12 * my) * oz + (d47 * oy + d56 * my) * mz);
      gy += g[j] * ((f03 - f12) * oz + (f47 - f56) * mz);
      gz += g[j] * (f0312 - f4756);
    }
    float s = c.res;
    if (p.radius > 0.f) s /= (2.f * p.radius);
    float* d = dx + i * lddx;
    atomicAdd(d + 0, gx * c.d[0] * s);
    atomicAdd(d + 1, gy * c.d[1] * s);
    atomicAdd(d + 2, gz * c.d[2] * s);
  }
}


// ---- F = 2 fast path: one thread per (point, group of 4 consecutive levels) ------------------------------------------
// The 8 floats a thread produces are one 32-byte sector of the point's feature row: full-sector stores (the
// one-level-per-thread mapping writes 8 of every 32 bytes, which the L2 turns into a read-modify-write of the whole
// sector from DRAM: 4x write and ~2 GB of fill traffic per 2.6 M look-ups in ncu).  blockIdx.y = level group, so the
// gathers of the concurrently running blocks stay inside 4 levels (16 MiB) of the table; 32 independent gathers per
// thread are in flight.
constexpr int LV = 4;

__global__ void __launch_bounds__(256, 2) hashgrid_fwd4_kernel(HashGridParams p, const float* __restrict__ x, int64_t ldx,
                                                            const float* __restrict__ table,
                                                            const float* __restrict__ mask, float* __restrict__ out,
                                                            int64_t ld_out, int64_t* __restrict__ idx_out, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int level0 = blockIdx.y * LV;
  if (i >= n) return;
  const float* xi = x + i * ldx;
  const float x0 = __ldg(xi), x1 = __ldg(xi + 1), x2 = __ldg(xi + 2);
  Corner c[LV];
  Feat<2> f[LV][8];
#pragma unroll
  for (int l = 0; l < LV; ++l) {
    c[l] = corners(p, level0 + l, x0, x1, x2);
    // The two corners of an x-edge (floor, ceil) hash to rows that differ in bit 0 only whenever floor(x) is even
    // (hash = x ^ (y, z terms), encodings.py:256-259): the rows then share one aligned 16-byte pair and ONE load serves
    // both — a quarter fewer L1 requests on average for the kernel's binding resource (the L1 tag stage).  Same values.
    constexpr int FL[4] = {3, 2, 7, 6}, CE[4] = {0, 1, 4, 5};      // (floor-x, ceil-x) corner of the 4 x-edges
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t hf = c[l].h[FL[e]], hc = c[l].h[CE[e]];
      if (hc == (hf ^ 1u) && !(hf & 1u)) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(table) + (hf >> 1));
        f[l][FL[e]].v[0] = t.x; f[l][FL[e]].v[1] = t.y;
        f[l][CE[e]].v[0] = t.z; f[l][CE[e]].v[1] = t.w;
      } else {
        f[l][FL[e]] = load_feat<2>(table, hf);
        f[l][CE[e]] = load_feat<2>(table, hc);
      }
    }
  }
  float r[2 * LV];
#pragma unroll
  for (int l = 0; l < LV; ++l) {
    const int level = level0 + l;
    if (idx_out) {
      int64_t* io = idx_out + (i * p.L + level) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) io[k] = int64_t(c[l].h[k]);
    }
    const float ox = c[l].o[0], oy = c[l].o[1], oz = c[l].o[2];
    const float mx = __fsub_rn(1.f, ox), my = __fsub_rn(1.f, oy), mz = __fsub_rn(1.f, oz);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      // ref: encodings.py:292-302
      const float f03 = lerp_ref(f[l][0].v[j], f[l][3].v[j], ox, mx);
      const float f12 = lerp_ref(f[l][1].v[j], f[l][2].v[j], ox, mx);
      const float f56 = lerp_ref(f[l][5].v[j], f[l][6].v[j], ox, mx);
      const float f47 = lerp_ref(f[l][4].v[j], f[l][7].v[j], ox, mx);
      const float f0312 = lerp_ref(f03, f12, oy, my);
      const float f4756 = lerp_ref(f47, f56, oy, my);
      float v = lerp_ref(f0312, f4756, oz, mz);
      if (mask) v = __fmul_rn(v, __ldg(mask + level * 2 + j));
      r[2 * l + j] = v;
    }
  }
  float* o = out + i * ld_out + level0 * 2;
  if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
    reinterpret_cast<float4*>(o)[0] = make_float4(r[0], r[1], r[2], r[3]);
    reinterpret_cast<float4*>(o)[1] = make_float4(r[4], r[5], r[6], r[7]);
  } else {
#pragma unroll
    for (int j = 0; j < 2 * LV; ++j) o[j] = r[j];
  }
}

__global__ void __launch_bounds__(256) hashgrid_bwd4_kernel(HashGridParams p, const float* __restrict__ x, int64_t ldx,
                                                            const float* __restrict__ table,
                                                            const float* __restrict__ mask,
                                                            const float* __restrict__ dout, int64_t ld_dout,
                                                            float* __restrict__ dtable, float* __restrict__ dx,
                                                            int64_t lddx, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int level0 = blockIdx.y * LV;
  if (i >= n) return;
  const float* xi = x + i * ldx;
  const float x0 = __ldg(xi), x1 = __ldg(xi + 1), x2 = __ldg(xi + 2);
  float g[2 * LV];
  const float* gi = dout + i * ld_dout + level0 * 2;
  if ((reinterpret_cast<uintptr_t>(gi) & 15) == 0) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(gi)), b = __ldg(reinterpret_cast<const float4*>(gi) + 1);
    g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 2 * LV; ++j) g[j] = __ldg(gi + j);
  }
  float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
  for (int l = 0; l < LV; ++l) {
    const int level = level0 + l;
    float gl[2] = {g[2 * l], g[2 * l + 1]};
    if (mask) {
      gl[0] *= __ldg(mask + level * 2);
      gl[1] *= __ldg(mask + level * 2 + 1);
    }
    if (gl[0] == 0.f && gl[1] == 0.f) continue;   // masked (coarse-to-fine) levels contribute exact zeros
    const Corner c = corners(p, level, x0, x1, x2);
    const float ox = c.o[0], oy = c.o[1], oz = c.o[2];
    const float mx = 1.f - ox, my = 1.f - oy, mz = 1.f - oz;
    // x-edges (floor-x corner, ceil-x corner): rows that differ in bit 0 only (floor(x) even, see the forward) are one
    // aligned 16-byte pair: ONE vector reduction / ONE load serves both corners
    constexpr int FL[4] = {3, 2, 7, 6}, CE[4] = {0, 1, 4, 5};
    if (dtable) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t hf = c.h[FL[e]], hc = c.h[CE[e]];
        // (weights evaluated in the association order of the unpaired code: ox * oy * oz = (ox * oy) * oz, ...)
        const float wf = e < 2 ? (e == 0 ? mx * oy * oz : mx * my * oz) : (e == 2 ? mx * oy * mz : mx * my * mz);
        const float wc = e < 2 ? (e == 0 ? ox * oy * oz : ox * my * oz) : (e == 2 ? ox * oy * mz : ox * my * mz);
        if (hc == (hf ^ 1u) && !(hf & 1u)) {
          atomicAdd(reinterpret_cast<float4*>(dtable) + (hf >> 1), make_float4(gl[0] * wf, gl[1] * wf, gl[0] * wc, gl[1] * wc));
        } else {
          atomic_add_feat<2>(dtable, hf, gl, wf);
          atomic_add_feat<2>(dtable, hc, gl, wc);
        }
      }
    }
    if (dx) {
      Feat<2> f[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t hf = c.h[FL[e]], hc = c.h[CE[e]];
        if (hc == (hf ^ 1u) && !(hf & 1u)) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(table) + (hf >> 1));
          f[FL[e]].v[0] = t.x; f[FL[e]].v[1] = t.y;
          f[CE[e]].v[0] = t.z; f[CE[e]].v[1] = t.w;
        } else {
          f[FL[e]] = load_feat<2>(table, hf);
          f[CE[e]] = load_feat<2>(table, hc);
        }
      }
      float lx = 0.f, ly = 0.f, lz = 0.f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float f03 = f[0].v[j] * ox + f[3].v[j] * mx, d03 = f[0].v[j] - f[3].v[j];
        const float f12 = f[1].v[j] * ox + f[2].v[j] * mx, d12 = f[1].v[j] - f[2].v[j];
        const float f56 = f[5].v[j] * ox + f[6].v[j] * mx, d56 = f[5].v[j] - f[6].v[j];
        const float f47 = f[4].v[j] * ox + f[7].v[j] * mx, d47 = f[4].v[j] - f[7].v[j];
        const float f0312 = f03 * oy + f12 * my, f4756 = f47 * oy + f56 * my;
        lx += gl[j] * ((d03 * oy + d12 * my) * oz + (d47 * oy + d56 * my) * mz);
        ly += gl[j] * ((f03 - f12) * oz + (f47 - f56) * mz);
        lz += gl[j] * (f0312 - f4756);
      }
      float sc = c.res;
      if (p.radius > 0.f) sc /= (2.f * p.radius);
      gx += lx * c.d[0] * sc; gy += ly * c.d[1] * sc; gz += lz * c.d[2] * sc;
    }
  }
  if (dx) {
    float* d = dx + i * lddx;
    atomicAdd(d + 0, gx);
    atomicAdd(d + 1, gy);
    atomicAdd(d + 2, gz);
  }
}

static int make_params(const MmsbHashGridDesc* d, HashGridParams& p) {
  MMSB_REQUIRE(d != nullptr, "hashgrid: desc is NULL");
  MMSB_REQUIRE(d->num_levels >= 1 && d->num_levels <= MMSB_MAX_LEVELS, "hashgrid: num_levels %d not in [1,%d]",
               d->num_levels, MMSB_MAX_LEVELS);
  MMSB_REQUIRE(d->log2_hashmap_size >= 1 && d->log2_hashmap_size <= 26 &&
                   (int64_t(d->num_levels) << d->log2_hashmap_size) < (int64_t(1) << 32),
               "hashgrid: log2_hashmap_size %d unsupported", d->log2_hashmap_size);
  MMSB_REQUIRE(d->interpolation == MMSB_INTERP_LINEAR || d->interpolation == MMSB_INTERP_SMOOTHSTEP,
               "hashgrid: interpolation %d unknown", d->interpolation);
  p.L = d->num_levels;
  p.log2T = d->log2_hashmap_size;
  p.interp = d->interpolation;
  p.radius = d->radius;
  for (int l = 0; l < MMSB_MAX_LEVELS; ++l) p.res[l] = l < d->num_levels ? d->resolution[l] : 0.f;
  return MMSB_OK;
}

}  // namespace mmsb

using namespace mmsb;

extern "C" int mmsb_hashgrid_fwd(const MmsbHashGridDesc* desc, const float* x, int64_t ldx, const float* table,
                                 const float* mask, float* out, int64_t ld_out, int64_t* idx_out, int64_t n,
                                 mmsb_stream_t stream) {
  HashGridParams p;
  if (int e = make_params(desc, p)) return e;
  MMSB_REQUIRE(n >= 0 && ldx >= 3 && ld_out >= int64_t(p.L) * desc->features_per_level,
               "hashgrid_fwd: bad sizes n=%lld ldx=%lld ld_out=%lld", (long long)n, (long long)ldx, (long long)ld_out);
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(x && table && out, "hashgrid_fwd: NULL pointer");
  dim3 grid((unsigned)ceil_div(n, 256), p.L), block(256);
  cudaStream_t s = as_stream(stream);
  if (desc->features_per_level == 2 && p.L % LV == 0) {
    grid.y = p.L / LV;
    hashgrid_fwd4_kernel<<<grid, block, 0, s>>>(p, x, ldx, table, mask, out, ld_out, idx_out, n);
    return check_launch("hashgrid_fwd");
  }
  switch (desc->features_per_level) {
    case 1: hashgrid_fwd_kernel<1><<<grid, block, 0, s>>>(p, x, ldx, table, mask, out, ld_out, idx_out, n); break;
    case 2: hashgrid_fwd_kernel<2><<<grid, block, 0, s>>>(p, x, ldx, table, mask, out, ld_out, idx_out, n); break;
    case 4: hashgrid_fwd_kernel<4><<<grid, block, 0, s>>>(p, x, ldx, table, mask, out, ld_out, idx_out, n); break;
    case 8: hashgrid_fwd_kernel<8><<<grid, block, 0, s>>>(p, x, ldx, table, mask, out, ld_out, idx_out, n); break;
    default:
      set_error("hashgrid_fwd: features_per_level %d not in {1,2,4,8}", desc->features_per_level);
      return MMSB_E_UNSUPPORTED;
  }
  return check_launch("hashgrid_fwd");
}

extern "C" int mmsb_hashgrid_bwd(const MmsbHashGridDesc* desc, const float* x, int64_t ldx, const float* table,
                                 const float* mask, const float* dout, int64_t ld_dout, float* dtable, float* dx,
                                 int64_t lddx, int64_t n, mmsb_stream_t stream) {
  HashGridParams p;
  if (int e = make_params(desc, p)) return e;
  MMSB_REQUIRE(n >= 0 && ldx >= 3 && ld_dout >= int64_t(p.L) * desc->features_per_level && (!dx || lddx >= 3),
               "hashgrid_bwd: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(x && table && dout, "hashgrid_bwd: NULL pointer");
  cudaStream_t s = as_stream(stream);
  if (dx) {
    cudaError_t e = cudaMemset2DAsync(dx, lddx * sizeof(float), 0, 3 * sizeof(float), n, s);
    if (e != cudaSuccess) {
      set_error("hashgrid_bwd: memset failed: %s", cudaGetErrorString(e));
      return MMSB_E_CUDA;
    }
  }
  dim3 grid((unsigned)ceil_div(n, 256), p.L), block(256);
  if (desc->features_per_level == 2 && p.L % LV == 0) {
    grid.y = p.L / LV;
    hashgrid_bwd4_kernel<<<grid, block, 0, s>>>(p, x, ldx, table, mask, dout, ld_dout, dtable, dx, lddx, n);
    return check_launch("hashgrid_bwd");
  }
  switch (desc->features_per_level) {
    case 1: hashgrid_bwd_kernel<1><<<grid, block, 0, s>>>(p, x, ldx, table, mask, dout, ld_dout, dtable, dx, lddx, n); break;
    case 2: hashgrid_bwd_kernel<2><<<grid, block, 0, s>>>(p, x, ldx, table, mask, dout, ld_dout, dtable, dx, lddx, n); break;
    case 4: hashgrid_bwd_kernel<4><<<grid, block, 0, s>>>(p, x, ldx, table, mask, dout, ld_dout, dtable, dx, lddx, n); break;
    case 8: hashgrid_bwd_kernel<8><<<grid, block, 0, s>>>(p, x, ldx, table, mask, dout, ld_dout, dtable, dx, lddx, n); break;
    default:
      set_error("hashgrid_bwd: features_per_level %d not in {1,2,4,8}", desc->features_per_level);
      return MMSB_E_UNSUPPORTED;
  }
  return check_launch("hashgrid_bwd");
}
