// A12 — the SDF network's forward (x -> h0 = act(W0 x + b0) -> h1 = act(W1 h0 + b1) -> sdf = w2[0] . h1 + b2[0]) as ONE
// persistent kernel: the hidden activations never leave the chip on their way to the next layer.
// ref: src/fields/surface_field.py:99-116 (SDFField.forward), src/field_components/mlp.py:152-171 (layer loop),
//      src/model_components/surface_model.py:143-146 (the tap / sampler evaluations keep only output 0).
//
// One CTA pair (tcgen05.mma.cta_group::2, M = 256 rows across two SMs) per 256-row tile, 22 warps per CTA:
//   loader (1 thread)      TMA: this CTA's 128 rows of x (three raw fp32 boxes of 32 columns) and its HALF (128 of the 256
//                          output columns) of every weight k-block of W0 (2 k-blocks) and W1 (4), through a 3-stage ring;
//   4 converter warps      two threads per row: fp32 row -> per-row power-of-two scale (from the row's 2-norm) -> fp16 hi / lo
//                          K-major SWIZZLE_128B operand tiles, IN PLACE over the landed boxes;
//   MMA issuer (1 thread)  layer 0 into TMEM columns [0, 256), layer 1 into [256, 512): three kind::f16 MMAs per product
//                          (lo*hi + hi*lo + hi*hi: 22 significant bits, the fp32-accurate mode) or one (fast mode);
//   16 epilogue warps      layer 0: tcgen05.ld -> bias / activation -> fp16 hi / lo of h0 (scaled by a per-row power of two
//                          bounded through |z| <= |x| max|W0 row| + max b0) written straight into a K-major operand chunk
//                          in shared memory (lane = row: no transpose), 64 columns at a time through a 2-slot ring, so the
//                          layer-1 MMAs of a chunk run while the next chunk is produced; layer 1: bias / activation /
//                          dot product with the head's weights -> one float per row.
// h0 / h1 are written to global memory only when the caller needs them (training: the backward kernels read them; a
// `group` > 1 keeps only every group-th row of h1 = the centre rows whose geometry features the caller evaluates), as
// 32-row blocks through cp.async.bulk.tensor stores (layer 1's blocks are staged in the h0 operand ring, idle by then).
// Waiting roles poll their mbarriers with a nanosleep back-off: the epilogue warps are issue-bound.
#include "tc_common.cuh"

namespace mmsb {
namespace tc {

constexpr int SF_EPI_WARPS = 16;                  // 4 per TMEM lane quadrant: each takes a quarter of the columns
constexpr int SF_PROD_WARPS = 4;
constexpr int SF_THREADS = (SF_EPI_WARPS + SF_PROD_WARPS + 2) * 32;
constexpr int SF_STG_LD = 8;                      // floats per row of a warp's transpose buffer (32 rows x 8 columns)
constexpr int SF_STG_BYTES = SF_EPI_WARPS * 32 * SF_STG_LD * 4;
constexpr int SF_HB = 128 * 128;                  // one CTA's half (128 rows of N) of one part (hi or lo) of a weight k-block
constexpr int SF_BSTAGE = 2 * SF_HB;              // hi + lo
constexpr int SF_BSTAGES = 3;
constexpr int SF_XREG = 3 * PART;                 // kb0 hi | kb0 lo | kb1 (hi at chunks 0-1, lo at chunks 2-3 of the row)
constexpr int SF_HSLOT = 2 * PART;                // one 64-column chunk of h0: hi | lo
constexpr int SF_HSLOTS = 2;
constexpr int SF_OFF_B = 0;
constexpr int SF_OFF_X = SF_OFF_B + SF_BSTAGES * SF_BSTAGE;
constexpr int SF_OFF_H = SF_OFF_X + SF_XREG;
constexpr int SF_OFF_STG = SF_OFF_H + SF_HSLOTS * SF_HSLOT;
constexpr int SF_OFF_XN = SF_OFF_STG + SF_STG_BYTES;         // float xnorm[2][128]
constexpr int SF_OFF_BAR = SF_OFF_XN + 2 * TM * 4;           // 17 mbarriers + constants
constexpr int SF_SMEM = SF_OFF_BAR + 256 + 1024 /*alignment*/;
static_assert(SF_SMEM <= 232448, "shared memory of the fused SDF kernel");

struct SdfFusedArgs {
  int64_t M; int K0;
  const float* w0; const float* b0; const float* b1; const float* head_w; const float* head_b;
  const float* amax_w0; const float* amax_w1;     // trailers of the packed fp16 buffers (max |w|)
  float act_param;
  float* h0; int64_t ldh0;
  float* h1; int64_t ldh1; int h1_group;
  int tma_store;
  float* sdf;
};

// Softplus(beta) = max(z, 0) + ln 2 / beta * lg2(1 + ex2(-beta log2(e) |z|)) with the two constants hoisted (c1, c2): six
// instructions per element, two of them on the MUFU pipe; ReLU: one
template <int ACT>
__device__ __forceinline__ float sf_act(float z, float c1, float c2) {
  if (ACT == MMSB_ACT_SOFTPLUS) return fmaf(fast_lg2(1.f + fast_ex2(fabsf(z) * c1)), c2, fmaxf(z, 0.f));
  return fmaxf(z, 0.f);
}

// per-column vectors (biases, head weights): 3 KB that every epilogue warp re-reads every tile; keep them in L1 while the
// activation stores stream through it
__device__ __forceinline__ float4 ldg_keep(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// 8 post-activation values of this lane's row (columns col0..col0+7) -> the warp's transpose buffer -> 32-byte row segments
// of C (16 rows per store instruction); group > 1: only rows that are multiples of `group` are kept, at row index / group
__device__ __forceinline__ void sf_store8(float* C, int64_t ld, int64_t M, int group, int64_t grow0, int col0, const float* h,
                                          float* stg, int lane) {
  __syncwarp();
  const int sw = (lane >> 2) & 1;
  sts128(smem_u32(stg + lane * SF_STG_LD + 4 * (0 ^ sw)), make_float4(h[0], h[1], h[2], h[3]));
  sts128(smem_u32(stg + lane * SF_STG_LD + 4 * (1 ^ sw)), make_float4(h[4], h[5], h[6], h[7]));
  __syncwarp();
  const int jj = lane & 1;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = (lane >> 1) + 16 * i;
    const float4 v = lds128(smem_u32(stg + r * SF_STG_LD + 4 * (jj ^ ((r >> 2) & 1))));
    int64_t row = grow0 + r;
    if (row >= M) continue;
    if (group > 1) {
      if (row % group) continue;
      row /= group;
    }
    stg_stream(C + row * ld + col0 + 4 * jj, v);
  }
}

// The same 32 x 8 block through the TMA: the lanes write their rows into the warp's (unswizzled) buffer and one lane
// issues a tensor-map store (rows past the end of the tensor are clipped by the copy).  No LDS / STG instruction and no
// L1 traffic; the buffer is reused once the previous copy has READ it.
template <int PENDING>
__device__ __forceinline__ void sf_store8_tma(const CUtensorMap* map, int grow0, int col0, const float* h, float* stg, int lane) {
  if (PENDING >= 0) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING >= 0 ? PENDING : 0) : "memory");
    __syncwarp();
  }
  const uint32_t dst = smem_u32(stg + lane * SF_STG_LD);
  sts128(dst, make_float4(h[0], h[1], h[2], h[3]));
  sts128(dst + 16, make_float4(h[4], h[5], h[6], h[7]));
  fence_async_smem();
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(col0), "r"(grow0),
                 "r"(smem_u32(stg))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

// 32 rows x 16 columns (64-byte rows, one 2 KB box): half as many rows for the copy engine as two 8-column blocks
__device__ __forceinline__ void sf_store16_tma(const CUtensorMap* map, int grow0, int col0, const float* h, float* stg, int lane) {
  const uint32_t dst = smem_u32(stg + lane * 16);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int jj = (j + (lane >> 1)) & 3;            // rotate the chunk order: the lanes of a quarter-warp hit distinct banks
    sts128(dst + 16 * jj, make_float4(h[4 * jj], h[4 * jj + 1], h[4 * jj + 2], h[4 * jj + 3]));
  }
  fence_async_smem();
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(col0), "r"(grow0),
                 "r"(smem_u32(stg))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

template <int ACT, int NPROD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SF_THREADS, 1)
    sdf_fused_fwd_kernel(const __grid_constant__ SdfFusedArgs g, const __grid_constant__ CUtensorMap tmap_x,
                         const __grid_constant__ CUtensorMap tmap_w0, const __grid_constant__ CUtensorMap tmap_w1,
                         const __grid_constant__ CUtensorMap tmap_h0, const __grid_constant__ CUtensorMap tmap_h1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sb = smem_u32(smem);
  const uint32_t BR = sb + SF_OFF_B, XR = sb + SF_OFF_X, HR = sb + SF_OFF_H;
  float* stg_all = reinterpret_cast<float*>(smem + SF_OFF_STG);
  float* xnorm = reinterpret_cast<float*>(smem + SF_OFF_XN);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SF_OFF_BAR);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t b_full = bar0, b_empty = bar0 + 24, x_raw = bar0 + 48, x_full = bar0 + 56, x_empty = bar0 + 64;
  const uint32_t a0_full = bar0 + 72, a0_empty = bar0 + 80, h_full = bar0 + 88, h_empty = bar0 + 104;
  const uint32_t a1_full = bar0 + 120, a1_empty = bar0 + 128;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float* consts = reinterpret_cast<float*>(bars + 20);        // [0] max row 2-norm of W0, [1] max b0 (as ordered uints >= 0)

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int64_t ptiles = (g.M + 2 * TM - 1) / (2 * TM);
  const int ks1 = (g.K0 - TK16 + 15) / 16;                    // k-steps of the second k-block of layer 0 (64 < K0 <= 80: 1)

  if (t == 0) {
    for (int s = 0; s < SF_BSTAGES; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, 1);
    }
    mbar_init(x_raw, 1);
    mbar_init(x_full, 2 * SF_PROD_WARPS);
    mbar_init(x_empty, 1);
    mbar_init(a0_full, 1);
    mbar_init(a0_empty, 2 * SF_EPI_WARPS);
    for (int s = 0; s < SF_HSLOTS; ++s) {
      mbar_init(h_full + 8 * s, 2 * SF_EPI_WARPS);
      mbar_init(h_empty + 8 * s, 1);
    }
    mbar_init(a1_full, 1);
    mbar_init(a1_empty, 2 * SF_EPI_WARPS);
    consts[0] = 0.f;
    consts[1] = 0.f;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp < SF_EPI_WARPS) {
    // bound of the layer-0 pre-activations: |z[r, c]| <= |x_r|_2 * max_c |W0[c, :]|_2 + max_c b0[c]
    float wn = 0.f, bm = 0.f;
    for (int c = t; c < NT; c += SF_EPI_WARPS * 32) {
      float s = 0.f;
      for (int k = 0; k < g.K0; ++k) { const float w = __ldg(g.w0 + int64_t(c) * g.K0 + k); s = fmaf(w, w, s); }
      wn = fmaxf(wn, sqrtf(s));
      bm = fmaxf(bm, g.b0 ? __ldg(g.b0 + c) : 0.f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wn = fmaxf(wn, __shfl_xor_sync(0xffffffffu, wn, o));
      bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
    }
    if (lane == 0) {
      atomicMax(reinterpret_cast<unsigned int*>(consts), __float_as_uint(wn));
      atomicMax(reinterpret_cast<unsigned int*>(consts + 1), __float_as_uint(bm));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float wn_max = consts[0], b_max = consts[1];

  if (warp < SF_EPI_WARPS) {
    // ================= epilogues (this CTA's 128 rows; warp = row quadrant q x column quarter `part`) =================
    const int q = warp & 3, part = warp >> 2;
    const int row = q * 32 + lane;
    float* stg = stg_all + warp * 32 * SF_STG_LD;
    const int ew0 = f16_scale_exp(__ldg(g.amax_w0)), ew1 = f16_scale_exp(__ldg(g.amax_w1));
    const float p = g.act_param;
    const float c1 = -p * 1.4426950408889634f, c2 = 0.6931471805599453f / p;
    const float head_b = g.head_b ? __ldg(g.head_b) : 0.f;
    const uint32_t lane_base = uint32_t(q * 32) << 16;
    uint32_t ti = 0;
    for (int64_t pt = pair; pt < ptiles; pt += npairs, ++ti) {
      const int64_t m0 = pt * 2 * TM + rank * TM;
      mbar_wait_backoff<64>(a0_full, ti & 1);
      tc_fence_after();
      // per-row scales: x was scaled by 2^ex (converters), h0 is scaled by 2^eh with one bit of headroom under the bound
      const float xn = xnorm[(ti & 1) * TM + row];
      const int ex = f16_scale_exp(xn);
      float hb = fmaxf(fmaf(xn, wn_max, b_max), 0.f);
      if (ACT == MMSB_ACT_SOFTPLUS) hb += 0.6931472f / p;
      const int eh = f16_scale_exp(hb) - 1;
      const float dsc0 = pow2f(-ex) * pow2f(-ew0), sh = pow2f(eh), dsc1 = pow2f(-eh) * pow2f(-ew1);
      // ---- layer 0: 4 chunks of 64 columns, this warp's 16 columns of each ----
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const uint32_t slot = c & 1, use = 2 * ti + (c >> 1);
        const int col0 = 64 * c + 16 * part;
        uint32_t v[16];
        tmem_ld16(tmem + lane_base + uint32_t(col0), v);
        tmem_ld_wait();
        if (c == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(a0_empty);
        }
        float h[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b4 = g.b0 ? ldg_keep(g.b0 + col0 + 4 * j4) : make_float4(0.f, 0.f, 0.f, 0.f);
          h[4 * j4 + 0] = sf_act<ACT>(fmaf(__uint_as_float(v[4 * j4 + 0]), dsc0, b4.x), c1, c2);
          h[4 * j4 + 1] = sf_act<ACT>(fmaf(__uint_as_float(v[4 * j4 + 1]), dsc0, b4.y), c1, c2);
          h[4 * j4 + 2] = sf_act<ACT>(fmaf(__uint_as_float(v[4 * j4 + 2]), dsc0, b4.z), c1, c2);
          h[4 * j4 + 3] = sf_act<ACT>(fmaf(__uint_as_float(v[4 * j4 + 3]), dsc0, b4.w), c1, c2);
        }
        mbar_wait_backoff<64>(h_empty + 8 * slot, (use & 1) ^ 1);
        const uint32_t hi_base = HR + slot * SF_HSLOT + uint32_t(row) * 128u;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          uint32_t hh[4], ll[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) split_f16x2_bounded(h[8 * i + 2 * k] * sh, h[8 * i + 2 * k + 1] * sh, hh[k], ll[k]);
          const uint32_t phys = uint32_t((2 * part + i) ^ (row & 7)) << 4;
          sts128u(hi_base + phys, hh[0], hh[1], hh[2], hh[3]);
          if (NPROD == 3) sts128u(hi_base + PART + phys, ll[0], ll[1], ll[2], ll[3]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(h_full + 8 * slot);
        if (g.h0 != nullptr) {
          if (g.tma_store) {
            sf_store8_tma<0>(&tmap_h0, int(m0) + q * 32, col0, &h[0], stg, lane);
            sf_store8_tma<0>(&tmap_h0, int(m0) + q * 32, col0 + 8, &h[8], stg, lane);
          } else {
            sf_store8(g.h0, g.ldh0, g.M, 1, m0 + q * 32, col0, &h[0], stg, lane);
            sf_store8(g.h0, g.ldh0, g.M, 1, m0 + q * 32, col0 + 8, &h[8], stg, lane);
          }
        }
      }
      // ---- layer 1: this warp's 64-column quarter, 16 columns at a time ----
      mbar_wait_backoff<64>(a1_full, ti & 1);
      tc_fence_after();
      float hacc = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col0 = 64 * part + 16 * cc;
        uint32_t v[16];
        tmem_ld16(tmem + uint32_t(NT) + lane_base + uint32_t(col0), v);
        tmem_ld_wait();
        if (cc == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(a1_empty);
        }
        float h[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b4 = g.b1 ? ldg_keep(g.b1 + col0 + 4 * j4) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 w4 = ldg_keep(g.head_w + col0 + 4 * j4);
          h[4 * j4 + 0] = sf_act<ACT>(fmaf(__uint_as_float(v[4 * j4 + 0]), dsc1, b4.x), c1, c2);
          h[4 * j4 + 1] = sf_act<ACT>(fmaf(__uint_as_float(v[4 * j4 + 1]), dsc1, b4.y), c1, c2);
          h[4 * j4 + 2] = sf_act<ACT>(fmaf(__uint_as_float(v[4 * j4 + 2]), dsc1, b4.z), c1, c2);
          h[4 * j4 + 3] = sf_act<ACT>(fmaf(__uint_as_float(v[4 * j4 + 3]), dsc1, b4.w), c1, c2);
          hacc = fmaf(h[4 * j4 + 0], w4.x, fmaf(h[4 * j4 + 1], w4.y, fmaf(h[4 * j4 + 2], w4.z, fmaf(h[4 * j4 + 3], w4.w, hacc))));
        }
        if (g.h1 != nullptr) {
          if (g.tma_store && g.h1_group == 1) {
            // staged in the h0 operand ring, which is idle until the next tile's layer 0 (every layer-1 MMA has completed):
            // 4 KB per warp = two chunks in flight, so a chunk only waits for the copy of the one before the previous
            float* s1 = reinterpret_cast<float*>(smem + SF_OFF_H) + warp * 1024 + (cc & 1) * 512;
            if (cc >= 2) {
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              __syncwarp();
            }
            sf_store16_tma(&tmap_h1, int(m0) + q * 32, col0, h, s1, lane);
          } else {
            if (g.tma_store) {      // the LSU path below reads the buffer back: the last TMA store must have read it
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
            }
            sf_store8(g.h1, g.ldh1, g.M, g.h1_group, m0 + q * 32, col0, &h[0], stg, lane);
            sf_store8(g.h1, g.ldh1, g.M, g.h1_group, m0 + q * 32, col0 + 8, &h[8], stg, lane);
          }
        }
      }
      // the four warps of a row quadrant combine their quarters of the head's dot product (fixed order: deterministic)
      if (g.tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      if (part != 0) stg[lane] = hacc;
      asm volatile("bar.sync 1, %0;" ::"n"(SF_EPI_WARPS * 32) : "memory");
      if (part == 0 && m0 + row < g.M) {
        const float* o = stg_all + q * 32 * SF_STG_LD + lane;
        g.sdf[m0 + row] = ((hacc + o[4 * 32 * SF_STG_LD]) + (o[8 * 32 * SF_STG_LD] + o[12 * 32 * SF_STG_LD])) + head_b;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(SF_EPI_WARPS * 32) : "memory");
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // outstanding activation stores
  } else if (warp < SF_EPI_WARPS + SF_PROD_WARPS) {
    // ================= converters: two threads per row of this CTA's x tile (lane = half * 16 + row % 16) =================
    // half 0 owns box 0 (k 0..31 -> fp16 chunks 0..3) and k 64..71, half 1 owns box 1 (k 32..63 -> chunks 4..7) and k 72..79
    const int cw = warp - SF_EPI_WARPS, half = lane >> 4;
    uint32_t ti = 0;
    for (int64_t pt = pair; pt < ptiles; pt += npairs, ++ti) {
      mbar_wait_backoff<256>(x_raw, ti & 1);
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const int r = cw * 32 + pass * 16 + (lane & 15);
        const uint32_t rowb = uint32_t(r) * 128u;
        const int sw = r & 7;
        float4 a[10];
#pragma unroll
        for (int c = 0; c < 8; ++c) a[c] = lds128(XR + half * PART + rowb + (uint32_t(c ^ sw) << 4));
#pragma unroll
        for (int c = 0; c < 2; ++c) a[8 + c] = lds128(XR + 2 * PART + rowb + (uint32_t((2 * half + c) ^ sw) << 4));
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 10; ++c) ss = fmaf(a[c].x, a[c].x, fmaf(a[c].y, a[c].y, fmaf(a[c].z, a[c].z, fmaf(a[c].w, a[c].w, ss))));
        ss += __shfl_xor_sync(0xffffffffu, ss, 16);
        const float nrm = sqrtf(ss);
        const float sx = pow2f(f16_scale_exp(nrm));
        __syncwarp();                  // every raw chunk of the 16 rows is in registers: the tiles can be overwritten
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          uint32_t hh[4], ll[4];
          split_f16x2_bounded(a[2 * j].x * sx, a[2 * j].y * sx, hh[0], ll[0]);
          split_f16x2_bounded(a[2 * j].z * sx, a[2 * j].w * sx, hh[1], ll[1]);
          split_f16x2_bounded(a[2 * j + 1].x * sx, a[2 * j + 1].y * sx, hh[2], ll[2]);
          split_f16x2_bounded(a[2 * j + 1].z * sx, a[2 * j + 1].w * sx, hh[3], ll[3]);
          if (j < 4) {
            // k-block 0: logical fp16 chunk 4 half + j of the hi tile (box 0's bytes) and of the lo tile (box 1's bytes)
            const uint32_t phys = uint32_t((4 * half + j) ^ sw) << 4;
            sts128u(XR + rowb + phys, hh[0], hh[1], hh[2], hh[3]);
            if (NPROD == 3) sts128u(XR + PART + rowb + phys, ll[0], ll[1], ll[2], ll[3]);
          } else {
            // k-block 1 (k 64..79): hi in logical chunks 0, 1 and lo in chunks 2, 3 of the same 128-byte row
            sts128u(XR + 2 * PART + rowb + (uint32_t(half ^ sw) << 4), hh[0], hh[1], hh[2], hh[3]);
            if (NPROD == 3) sts128u(XR + 2 * PART + rowb + (uint32_t((half + 2) ^ sw) << 4), ll[0], ll[1], ll[2], ll[3]);
          }
        }
        if (half == 0) xnorm[(ti & 1) * TM + r] = nrm;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(x_full);
    }
  } else if (warp == SF_EPI_WARPS + SF_PROD_WARPS) {
    // ================= loader =================
    if (lane == 0) {
      uint32_t bi = 0, xi = 0;
      auto load_x = [&](int64_t pt) {
        const int m0 = int(pt * 2 * TM + rank * TM);
        mbar_wait_backoff<128>(x_empty, (xi & 1) ^ 1);
        mbar_arrive_expect_tx(x_raw, uint32_t(3 * PART));
        tma_load_2d(XR, &tmap_x, 0, m0, x_raw);
        tma_load_2d(XR + PART, &tmap_x, TK, m0, x_raw);
        tma_load_2d(XR + 2 * PART, &tmap_x, 2 * TK, m0, x_raw);
        ++xi;
      };
      auto load_b = [&](const CUtensorMap* map, int kb) {
        const uint32_t s = bi % SF_BSTAGES;
        mbar_wait_backoff<128>(b_empty + 8 * s, ((bi / SF_BSTAGES) & 1) ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(b_full + 8 * s, uint32_t((NPROD == 3 ? 4 : 2) * SF_HB));
        const int row_hi = kb * 2 * NT + int(rank) * 128;
        tma_load_2d_pair(BR + s * SF_BSTAGE, map, 0, row_hi, b_full + 8 * s);
        if (NPROD == 3) tma_load_2d_pair(BR + s * SF_BSTAGE + SF_HB, map, 0, row_hi + NT, b_full + 8 * s);
        ++bi;
      };
      if (pair < ptiles) {
        load_x(pair);
        load_b(&tmap_w0, 0);
        if (ks1 > 0) load_b(&tmap_w0, 1);
      }
      for (int64_t pt = pair; pt < ptiles; pt += npairs) {
        for (int kb = 0; kb < 4; ++kb) load_b(&tmap_w1, kb);
        if (pt + npairs < ptiles) {
          load_x(pt + npairs);
          load_b(&tmap_w0, 0);
          if (ks1 > 0) load_b(&tmap_w0, 1);
        }
      }
    }
  } else if (warp == SF_EPI_WARPS + SF_PROD_WARPS + 1 && rank == 0) {
    // ================= MMA issuer (leader CTA) =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(NT, 2 * TM);
      uint32_t bi = 0;
      auto mma_block = [&](uint64_t dah, uint64_t dal, int ksteps, uint32_t d, bool first) {
        const uint32_t s = bi % SF_BSTAGES;
        mbar_wait_backoff<32>(b_full + 8 * s, (bi / SF_BSTAGES) & 1);
        tc_fence_after();
        const uint64_t dbh = make_desc(BR + s * SF_BSTAGE, 16, 1024), dbl = make_desc(BR + s * SF_BSTAGE + SF_HB, 16, 1024);
        for (int j = 0; j < ksteps; ++j) {
          const uint64_t adv = uint64_t(j * 2);
          const uint32_t acc = (first && j == 0) ? 0u : 1u;
          if (NPROD == 3) {
            umma_f16_pair(d, dal + adv, dbh + adv, idesc, acc);
            umma_f16_pair(d, dah + adv, dbl + adv, idesc, 1u);
            umma_f16_pair(d, dah + adv, dbh + adv, idesc, 1u);
          } else {
            umma_f16_pair(d, dah + adv, dbh + adv, idesc, acc);
          }
        }
        umma_commit_pair(b_empty + 8 * s);
        ++bi;
      };
      auto layer0 = [&](uint32_t ti) {
        mbar_wait_backoff<32>(x_full, ti & 1);
        mbar_wait_backoff<32>(a0_empty, (ti & 1) ^ 1);
        tc_fence_after();
        mma_block(make_desc(XR, 16, 1024), make_desc(XR + PART, 16, 1024), TK16 / 16, tmem, true);
        if (ks1 > 0) mma_block(make_desc(XR + 2 * PART, 16, 1024), make_desc(XR + 2 * PART, 16, 1024) + 2, ks1, tmem, false);
        umma_commit_pair(x_empty);
        umma_commit_pair(a0_full);
      };
      if (pair < ptiles) layer0(0);
      uint32_t ti = 0;
      for (int64_t pt = pair; pt < ptiles; pt += npairs, ++ti) {
        mbar_wait_backoff<32>(a1_empty, (ti & 1) ^ 1);
        tc_fence_after();
        for (int c = 0; c < 4; ++c) {
          const uint32_t slot = c & 1, use = 2 * ti + (c >> 1);
          mbar_wait_backoff<32>(h_full + 8 * slot, use & 1);
          tc_fence_after();
          const uint32_t hb = HR + slot * SF_HSLOT;
          mma_block(make_desc(hb, 16, 1024), make_desc(hb + PART, 16, 1024), TK16 / 16, tmem + NT, c == 0);
          umma_commit_pair(h_empty + 8 * slot);
        }
        umma_commit_pair(a1_full);
        if (pt + npairs < ptiles) layer0(ti + 1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// Tensor map of the x rows: [M rows, K floats] fp32, row stride ld floats, box 32 floats x 128 rows, SWIZZLE_128B;
// columns / rows outside the tensor are zero-filled by the copy.
static bool make_x_map(const float* x, int64_t ld, int64_t rows, int cols, CUtensorMap* map) {
  TensorMapEncodeFn encode = tensor_map_encoder();
  if (encode == nullptr) return false;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (ld & 3) != 0 || rows >= (int64_t(1) << 31)) return false;
  const cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t strides[1] = {cuuint64_t(ld) * sizeof(float)};
  const cuuint32_t box[2] = {cuuint32_t(TK), cuuint32_t(TM)};
  const cuuint32_t estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Tensor map of an activation output: [M rows, 256 floats], box 8 floats x 32 rows, no swizzle (the epilogue warps' store blocks)
static bool make_out_map(float* y, int64_t ld, int64_t rows, int box_cols, CUtensorMap* map) {
  TensorMapEncodeFn encode = tensor_map_encoder();
  if (encode == nullptr || rows >= (int64_t(1) << 31)) return false;
  const cuuint64_t dims[2] = {cuuint64_t(NT), cuuint64_t(rows)};
  const cuuint64_t strides[1] = {cuuint64_t(ld) * sizeof(float)};
  const cuuint32_t box[2] = {cuuint32_t(box_cols), 32};
  const cuuint32_t estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int ACT, int NPROD>
static int launch_sdf_fused(const SdfFusedArgs& g_in, const CUtensorMap& mx, const CUtensorMap& m0, const CUtensorMap& m1,
                            cudaStream_t s) {
  SdfFusedArgs g = g_in;
  CUtensorMap mh0, mh1;
  memset(&mh0, 0, sizeof(mh0));
  memset(&mh1, 0, sizeof(mh1));
  static int use_tma_store = -1;
  if (use_tma_store < 0) { const char* e = getenv("MMSB_SDF_TMA_STORE"); use_tma_store = e ? atoi(e) : 1; }
  g.tma_store = use_tma_store && (g.h0 == nullptr || make_out_map(g.h0, g.ldh0, g.M, SF_STG_LD, &mh0)) &&
                (g.h1 == nullptr || g.h1_group > 1 || make_out_map(g.h1, g.ldh1, g.M, 16, &mh1));
  static PerDeviceFlag flags;
  bool& configured = flags();
  auto kern = sdf_fused_fwd_kernel<ACT, NPROD>;
  if (!configured) {
    int rc = set_smem(kern, SF_SMEM, "sdf_net_fwd_fused");
    if (rc) return rc;
    configured = true;
  }
  const int64_t ptiles = ceil_div(g.M, 2 * TM);
  const int64_t pairs = ptiles < kNumSMs / 2 ? ptiles : kNumSMs / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(SF_THREADS);
  cfg.dynamicSmemBytes = SF_SMEM;
  cfg.stream = s;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, g, mx, m0, m1, mh0, mh1);
  if (e != cudaSuccess) {
    set_error("sdf_net_fwd_fused: cluster launch failed: %s", cudaGetErrorString(e));
    return MMSB_E_CUDA;
  }
  return check_launch("sdf_net_fwd_fused");
}

}  // namespace tc
}  // namespace mmsb

using namespace mmsb;

extern "C" int mmsb_sdf_net_fwd_fused(const float* x, int64_t ldx, int64_t n, int32_t in_dim, int32_t hidden,
                                      const float* w0, const float* packed_w0, const float* b0, const float* packed_w1,
                                      const float* b1, const float* head_w, const float* head_b, int32_t act,
                                      float act_param, int32_t products, float* h0, int64_t ldh0, float* h1, int64_t ldh1,
                                      int32_t h1_group, float* sdf, mmsb_stream_t stream) {
  MMSB_REQUIRE(x && w0 && packed_w0 && packed_w1 && head_w && sdf, "sdf_net_fwd_fused: null pointer");
  MMSB_REQUIRE(n >= 0 && ldx >= in_dim, "sdf_net_fwd_fused: bad shape");
  MMSB_REQUIRE(hidden == tc::NT && in_dim > tc::TK16 && in_dim <= tc::TK16 + 16,
               "sdf_net_fwd_fused: built for hidden width 256 and 64 < in_dim <= 80 (got hidden %d, in_dim %d)", hidden, in_dim);
  MMSB_REQUIRE(act == MMSB_ACT_SOFTPLUS || act == MMSB_ACT_RELU, "sdf_net_fwd_fused: activation must be ReLU or Softplus, got %d", act);
  MMSB_REQUIRE(act != MMSB_ACT_SOFTPLUS || act_param > 0.f, "sdf_net_fwd_fused: Softplus beta must be positive");
  MMSB_REQUIRE(products == 1 || products == 3, "sdf_net_fwd_fused: products must be 3 (fp16 split, fp32-accurate) or 1 (single fp16 pass), got %d", products);
  MMSB_REQUIRE(h1_group >= 1, "sdf_net_fwd_fused: h1_group must be >= 1");
  MMSB_REQUIRE((!h0 || (ldh0 >= hidden && (ldh0 & 3) == 0 && (reinterpret_cast<uintptr_t>(h0) & 15) == 0)) &&
                   (!h1 || (ldh1 >= hidden && (ldh1 & 3) == 0 && (reinterpret_cast<uintptr_t>(h1) & 15) == 0)),
               "sdf_net_fwd_fused: h0 / h1 need 16-byte aligned rows");
  MMSB_REQUIRE((!b0 || (reinterpret_cast<uintptr_t>(b0) & 15) == 0) && (!b1 || (reinterpret_cast<uintptr_t>(b1) & 15) == 0) &&
                   (reinterpret_cast<uintptr_t>(head_w) & 15) == 0,
               "sdf_net_fwd_fused: b0, b1 and head_w must be 16-byte aligned");
  if (n == 0) return MMSB_OK;
  CUtensorMap mx, m0, m1;
  memset(&mx, 0, sizeof(mx));
  memset(&m0, 0, sizeof(m0));
  memset(&m1, 0, sizeof(m1));
  if (!tc::make_x_map(x, ldx, n, in_dim, &mx)) {
    set_error("sdf_net_fwd_fused: x needs 16-byte aligned rows (base %% 16 = %d, ldx %lld) and fewer than 2^31 rows",
              int(reinterpret_cast<uintptr_t>(x) & 15), (long long)ldx);
    return MMSB_E_INVALID_ARGUMENT;
  }
  const int nkb0 = (in_dim + tc::TK16 - 1) / tc::TK16, nkb1 = hidden / tc::TK16;
  if (!tc::make_packed_map(packed_w0, int64_t(tc::NT) * nkb0 * 2, &m0) ||
      !tc::make_packed_map(packed_w1, int64_t(tc::NT) * nkb1 * 2, &m1)) {
    set_error("sdf_net_fwd_fused: could not encode the tensor maps of the packed weights");
    return MMSB_E_CUDA;
  }
  tc::SdfFusedArgs g{};
  g.M = n; g.K0 = in_dim;
  g.w0 = w0; g.b0 = b0; g.b1 = b1; g.head_w = head_w; g.head_b = head_b;
  g.amax_w0 = packed_w0 + tc::packed_floats_f16(hidden, in_dim);
  g.amax_w1 = packed_w1 + tc::packed_floats_f16(hidden, hidden);
  g.act_param = act_param;
  g.h0 = h0; g.ldh0 = ldh0; g.h1 = h1; g.ldh1 = ldh1; g.h1_group = h1_group; g.sdf = sdf;
  cudaStream_t s = as_stream(stream);
  if (act == MMSB_ACT_SOFTPLUS)
    return products == 3 ? tc::launch_sdf_fused<MMSB_ACT_SOFTPLUS, 3>(g, mx, m0, m1, s)
                         : tc::launch_sdf_fused<MMSB_ACT_SOFTPLUS, 1>(g, mx, m0, m1, s);
  return products == 3 ? tc::launch_sdf_fused<MMSB_ACT_RELU, 3>(g, mx, m0, m1, s)
                       : tc::launch_sdf_fused<MMSB_ACT_RELU, 1>(g, mx, m0, m1, s);
}
