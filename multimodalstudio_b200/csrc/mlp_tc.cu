// A11 — dense layers of the field MLPs on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// TMEM).  ref: src/field_components/mlp.py:152-171 (y = act(W x + b) per layer).
//
// fp32 in / fp32 out.  Precision 3 ("3xTF32"): every fp32 operand x is split into hi = tf32(x) and
// lo = tf32(x - hi) and the product is accumulated as lo*hi + hi*lo + hi*hi (three kind::tf32 MMAs, fp32
// accumulation in TMEM), which keeps the 1e-5 parity band of the fp32 reference on the tensor pipe.
// Precision 1: single-pass TF32 (the reference's GPU runs use fp16 autocast; 1e-2 band).
//
// Three persistent, warp-specialised kernels (one CTA per SM, or one CTA pair per two SMs):
//   tc_rows_kernel  (forward, dgrad)   D[rows x N] = A[rows x K] * Bp[N x K]^T, 18 warps
//       A  = activations, fp32 in global memory.  TMA path (16-byte aligned rows): a tensor-map copy lands the raw
//            fp32 k-block as the hi tile (the tensor core reads fp32 as TF32 by truncation), 8 converter warps derive
//            lo = rna(x - trunc(x)) from shared memory.  Otherwise 8 producer warps load with a register prefetch,
//            split into hi / lo and store K-major SWIZZLE_128B tiles.
//       Bp = weights pre-split and pre-swizzled by pack_weight_kernel (once per step and layer): one bulk async copy
//            (cp.async.bulk + mbarrier complete_tx) lands a k-block of B in its stage;
//       1 thread issues tcgen05.mma into one of two 128x256 fp32 accumulators in TMEM (512 columns), so the 8 epilogue
//       warps (tcgen05.ld -> shared-memory transpose -> bias / activation or activation derivative -> 16-byte row
//       stores; branch-free interior path, bounds-checked edge path) drain tile i while the MMAs of tile i+1 run.
//   tc_rows_pair_kernel  the same products for 256-wide accumulators on CTA pairs (tcgen05.mma.cta_group::2, M = 256
//       across two SMs): each CTA stages its 128 rows of A and HALF of the weight k-block, 3-stage ring of 64 KB,
//       14 warps per CTA (8 epilogue, 4 converter, loader, MMA issuer) = up to 128 registers per thread.
//   tc_wgrad_kernel (weight gradient)  dW[out x in] += dz[rows x out]^T * x[rows x in], split over rows
//       both operands are MN-major in global memory and are staged without a transpose into MN-major
//       SWIZZLE_128B_BASE32B tiles by 16 staging warps; the bias gradient (column sums of dz) is accumulated on the way;
//       partial tiles are combined with coalesced fp32 reductions (red.global.add); PAIR variant: the two 128-row
//       m-tiles of a 256 x 256 gradient on a CTA pair, each CTA staging half of the x columns.
#include "tc_common.cuh"

namespace mmsb {
namespace tc {
// ---- packed weights ---------------------------------------------------------------------------------------
// Operand B of a rows product: logical [N x K], element (n, k) = w[n*ldw + k] (forward: B = W) or
// w[k*ldw + n] (dgrad: B = W^T).  Packed as n-tiles of up to 256 rows (N padded to a multiple of 16), each a
// sequence of k-blocks of 32 (K padded), each k-block = [hi tile][lo tile], a tile = rows x 128 B, SWIZZLE_128B.
__global__ void pack_weight_kernel(const float* __restrict__ w, int64_t ldw, int n, int k, int transpose, int nparts,
                                   float* __restrict__ packed) {
  const int n_pad = pad16(n), nkb = (k + TK - 1) / TK;
  const int64_t total = int64_t(n_pad) * nkb * TK;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int kk = int(i % (nkb * TK)), nn = int(i / (nkb * TK));
    float v = 0.f;
    if (nn < n && kk < k) v = transpose ? __ldg(w + int64_t(kk) * ldw + nn) : __ldg(w + int64_t(nn) * ldw + kk);
    const int nt = nn / NT, r = nn % NT, wdt = tile_width(n_pad, nt);
    const int kb = kk / TK, kin = kk % TK;
    const int64_t base = int64_t(nt) * NT * TK * nkb * nparts + int64_t(kb) * nparts * wdt * TK;
    const int64_t off = int64_t(r) * TK + (((kin >> 2) ^ (r & 7)) << 2) + (kin & 3);
    const float h = tf32_rna(v);
    packed[base + off] = h;
    if (nparts == 2) packed[base + int64_t(wdt) * TK + off] = tf32_rna(v - h);
  }
}

// precision 2: the same tiling with fp16 hi / lo tiles of 64 k per 128-byte row (k-blocks of TK16), scaled by the power of
// two derived from *amax (max |w|, computed by amax_kernel just before); the packed buffer ends with one float holding
// that amax (the consuming kernels read their B scale from there).
__global__ void pack_weight_f16_kernel(const float* __restrict__ w, int64_t ldw, int n, int k, int transpose,
                                       const float* __restrict__ amax, uint32_t* __restrict__ packed) {
  const int n_pad = pad16(n), nkb = (k + TK16 - 1) / TK16;
  const float sc = pow2f(f16_scale_exp(__ldg(amax)));
  const int64_t total = int64_t(n_pad) * nkb * 32;       // pairs of consecutive k
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int kk = 2 * int(i % (nkb * 32)), nn = int(i / (nkb * 32));
    float v0 = 0.f, v1 = 0.f;
    if (nn < n) {
      if (kk < k) v0 = transpose ? __ldg(w + int64_t(kk) * ldw + nn) : __ldg(w + int64_t(nn) * ldw + kk);
      if (kk + 1 < k) v1 = transpose ? __ldg(w + int64_t(kk + 1) * ldw + nn) : __ldg(w + int64_t(nn) * ldw + kk + 1);
    }
    const int nt = nn / NT, r = nn % NT, wdt = tile_width(n_pad, nt);
    const int kb = kk / TK16, kin = kk % TK16;
    const int64_t base = int64_t(nt) * NT * nkb * 64 + int64_t(kb) * 2 * wdt * 32;
    const int64_t off = int64_t(r) * 32 + (((kin >> 3) ^ (r & 7)) << 2) + ((kin & 7) >> 1);
    uint32_t hi, lo;
    split_f16x2(v0 * sc, v1 * sc, hi, lo);
    packed[base + off] = hi;
    packed[base + int64_t(wdt) * 32 + off] = lo;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) reinterpret_cast<float*>(packed)[packed_floats_f16(n, k)] = __ldg(amax);
}

// out[0] = max(out[0], max |x[r, c]|) over an [n, cols] row-strided matrix (out >= 0: compared as unsigned bits)
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, int64_t ld, int64_t n, int cols,
                                                   float* __restrict__ out) {
  float m = 0.f;
  const int64_t total = n * cols;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols;
    m = fmaxf(m, fabsf(__ldg(x + r * ld + (i - r * cols))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
}
// contiguous variant: 16-byte loads
__global__ void __launch_bounds__(256) amax4_kernel(const float4* __restrict__ x, int64_t n4, float* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(x + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
}

// ---- epilogue of one 32-row x 32-column chunk: registers -> transpose buffer -> coalesced global ------------
enum { EPI_FWD = 0, EPI_DGRAD = 1, EPI_ATOMIC = 2 };

struct EpiArgs {
  float* C; int64_t ldc; int64_t M; int N;   // logical extents of C
  int64_t cs;                                // column stride of C for the reducing epilogue (0 = 1; != 1: transposed store)
  int direct;                                // 2: fragment-layout epilogue (no shared-memory transpose), 0: transposed
  const float* bias; int act; float act_param;
  const float* yprev; int64_t ld_yprev; int act_prev; float act_prev_param;
  // forward with a fused one-output head: head_out[row] += sum_col act(z)[row, col] * head_w[col] (+ head_b once);
  // C may be NULL (the activations themselves are then not stored)
  const float* head_w; const float* head_b; float* head_out;
  // dgrad with a rank-1 term: (acc + r1_d[row] * r1_w[col]) * act'(yprev)
  const float* r1_d; const float* r1_w;
  int accumulate;                            // dgrad: C += result instead of C = result
  // precision 2: acc_scale = 1 / (s_a s_b) is derived per tile from the operands' amax; amax_out (optional) receives
  // max |stored value| (the scale source of the kernel that consumes C as its streamed operand)
  const float* amax_a; const float* amax_b; float* amax_out;
};
// per-warp epilogue state of one tile: fused-head partial sums, accumulator scale (precision 2), running max |stored|
struct EpiState {
  float hacc[4];
  float sc;
  float amax;
};

template <int ACT>
__device__ __forceinline__ float4 act_fwd4(float4 x, float p) {
  x.x = act_fwd(x.x, ACT, p); x.y = act_fwd(x.y, ACT, p); x.z = act_fwd(x.z, ACT, p); x.w = act_fwd(x.w, ACT, p);
  return x;
}
template <int ACT>
__device__ __forceinline__ float4 act_bwd4(float4 y, float p) {
  y.x = act_bwd_from_y(y.x, ACT, p); y.y = act_bwd_from_y(y.y, ACT, p); y.z = act_bwd_from_y(y.z, ACT, p);
  y.w = act_bwd_from_y(y.w, ACT, p);
  return y;
}

// A chunk is 32 rows x CH columns of the accumulator.  After the transpose through shared memory lane l owns the 4
// consecutive columns col0 + 4 (l % 4) of the rows l / 4 + 8 i (i < 4): bias / activation / derivative work is per
// float4 and a warp's global accesses cover 64 contiguous bytes of 8 rows.
struct YPrev {
  float4 v[4];
  float d[4];    // rank-1 row factors (EpiArgs::r1_d) of the 4 rows
};

template <int EPI>
__device__ __forceinline__ void load_yprev(const EpiArgs& e, YPrev& y, int lane, int64_t row0, int col0, bool vec_y) {
  if (EPI != EPI_DGRAD) return;
  if (e.r1_d) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t row = row0 + (lane >> 2) + 8 * i;
      y.d[i] = row < e.M ? __ldg(e.r1_d + row) : 0.f;
    }
  }
  if (e.yprev == nullptr || e.act_prev == MMSB_ACT_NONE) return;
  const int col = col0 + 4 * (lane & 3);
  if (vec_y && row0 + 32 <= e.M && col0 + CH <= e.N) {
    // interior chunk (warp-uniform): four independent 16-byte loads, no bounds checks
    const float* p = e.yprev + (row0 + (lane >> 2)) * e.ld_yprev + col;
    const int64_t step = 8 * e.ld_yprev;
#pragma unroll
    for (int i = 0; i < 4; ++i) y.v[i] = __ldg(reinterpret_cast<const float4*>(p + i * step));
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = row0 + (lane >> 2) + 8 * i;
    y.v[i] = load4(e.yprev, e.ld_yprev, row, e.M, col, e.N, vec_y);
  }
}

// 4 values of a per-column vector at col..col+3 (zero past n)
__device__ __forceinline__ float4 load_cols4(const float* __restrict__ p, int col, int n) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p == nullptr) return v;
  v.x = __ldg(p + col);
  if (col + 1 < n) v.y = __ldg(p + col + 1);
  if (col + 2 < n) v.z = __ldg(p + col + 2);
  if (col + 3 < n) v.w = __ldg(p + col + 3);
  return v;
}

// 4 values of a per-column vector at col..col+3, all inside the vector (interior chunks)
__device__ __forceinline__ float4 load_cols4_in(const float* __restrict__ p, int col) {
  if (p == nullptr) return make_float4(0.f, 0.f, 0.f, 0.f);
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) return __ldg(reinterpret_cast<const float4*>(p + col));
  return make_float4(__ldg(p + col), __ldg(p + col + 1), __ldg(p + col + 2), __ldg(p + col + 3));
}

// Interior chunk (all 32 rows and CH columns inside C, 16-byte aligned rows; the test is warp-uniform): the four rows of
// a lane are processed as four independent, branch-free dependency chains (shared-memory read -> MUFU -> store), so
// their latencies overlap; the edge chunks take the bounds-checked path below.
// per-column vectors of a lane's 4 columns (bias, fused-head weights | rank-1 weights), fetched at the top of the chunk
// iteration so that their latency is covered by the TMEM wait and the transpose
struct ColVecs {
  float4 b4, h4, r4;
};
template <int EPI>
__device__ __forceinline__ void load_colvecs(const EpiArgs& e, ColVecs& cv, int lane, int col0) {
  const int col = col0 + 4 * (lane & 3);
  cv.b4 = cv.h4 = cv.r4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (EPI == EPI_FWD) {
    cv.b4 = load_cols4_in(e.bias, col);
    cv.h4 = load_cols4_in(e.head_w, col);
  }
  if (EPI == EPI_DGRAD && e.r1_d) cv.r4 = load_cols4_in(e.r1_w, col);
}

template <int EPI, int ACT, bool RICH, typename Next>
__device__ __forceinline__ void epilogue_rows_in(const EpiArgs& e, const float* stg, const YPrev& yp, const ColVecs& cv,
                                                 int lane, int64_t row0, int col0, EpiState& st, Next issue_next) {
  float (&hacc)[4] = st.hacc;
  const int q4 = lane & 3, r0 = lane >> 2;
  const int col = col0 + 4 * q4;
  const float4 b4 = cv.b4, h4 = cv.h4, r4 = cv.r4;
  // rows in flight per lane (the dgrad kernels also hold two derivative operands: fewer registers left, unless the
  // kernel runs with the 128-register budget of the 14-warp pair CTAs, RICH)
  constexpr int G = (EPI == EPI_DGRAD && !RICH) ? MMSB_TC_EPI_ILP / 2 : MMSB_TC_EPI_ILP;
  // transpose buffer: row r holds its four 16-byte chunks at positions c ^ ((r >> 1) & 3): the row-per-lane stores of
  // epilogue_stage and these 4-lanes-per-row loads are both free of bank conflicts (rows r0 + 8 i share the swizzle)
  const uint32_t src = smem_u32(stg + r0 * STG_LD + 4 * (q4 ^ ((r0 >> 1) & 3)));
  float* dst = e.C ? e.C + (row0 + r0) * e.ldc + col : nullptr;
  const int64_t step = 8 * e.ldc;
#pragma unroll
  for (int i0 = 0; i0 < 4; i0 += G) {
    float4 x[G];
#pragma unroll
    for (int j = 0; j < G; ++j) x[j] = lds128(src + uint32_t((i0 + j) * 8 * STG_LD * 4));
    // the transposing stores have been consumed by now: the accumulator registers can take the next TMEM read without
    // the read-after-write stall an issue right behind the stores would see
    if (i0 == 0) issue_next();
    if (e.amax_a != nullptr) {           // precision 2: undo the operands' power-of-two scales (exact)
#pragma unroll
      for (int j = 0; j < G; ++j) { x[j].x *= st.sc; x[j].y *= st.sc; x[j].z *= st.sc; x[j].w *= st.sc; }
    }
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int i = i0 + j;
      if (EPI == EPI_FWD) {
        x[j].x += b4.x; x[j].y += b4.y; x[j].z += b4.z; x[j].w += b4.w;
        x[j] = act_fwd4<ACT>(x[j], e.act_param);
        if (e.head_w) hacc[i] = fmaf(x[j].x, h4.x, fmaf(x[j].y, h4.y, fmaf(x[j].z, h4.z, fmaf(x[j].w, h4.w, hacc[i]))));
      } else {
        if (e.r1_d) {
          const float d = yp.d[i];
          x[j].x = fmaf(d, r4.x, x[j].x); x[j].y = fmaf(d, r4.y, x[j].y); x[j].z = fmaf(d, r4.z, x[j].z);
          x[j].w = fmaf(d, r4.w, x[j].w);
        }
        if (ACT != MMSB_ACT_NONE) {
          const float4 y = act_bwd4<ACT>(yp.v[i], e.act_prev_param);
          x[j].x *= y.x; x[j].y *= y.y; x[j].z *= y.z; x[j].w *= y.w;
        }
      }
    }
    if (dst != nullptr) {
      if (EPI == EPI_DGRAD && e.accumulate) {
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const float4 c = *reinterpret_cast<const float4*>(dst + (i0 + j) * step);
          x[j].x += c.x; x[j].y += c.y; x[j].z += c.z; x[j].w += c.w;
        }
      }
      if (e.amax_out != nullptr) {
#pragma unroll
        for (int j = 0; j < G; ++j)
          st.amax = fmaxf(fmaxf(st.amax, fmaxf(fabsf(x[j].x), fabsf(x[j].y))), fmaxf(fabsf(x[j].z), fabsf(x[j].w)));
      }
#pragma unroll
      for (int j = 0; j < G; ++j) *reinterpret_cast<float4*>(dst + (i0 + j) * step) = x[j];
    }
  }
}

// Edge chunks (last rows / columns of C, unaligned rows) and the reducing epilogue: bounds-checked, activation chosen at
// run time (one copy per kernel: the interior path above is the one that has to be fast).
template <int EPI>
__device__ __forceinline__ void epilogue_rows_edge(const EpiArgs& e, const float* stg, const YPrev& yp, int lane, int64_t row0,
                                                int col0, bool vec_ok, EpiState& st, int ACT) {
  float (&hacc)[4] = st.hacc;
  const int q4 = lane & 3;
  const int col = col0 + 4 * q4;
  if (col >= e.N) return;
  const bool full = col + 3 < e.N;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), h4 = b4, r4 = b4;
  if (EPI == EPI_FWD) {
    b4 = load_cols4(e.bias, col, e.N);
    h4 = load_cols4(e.head_w, col, e.N);
  }
  if (EPI == EPI_DGRAD) r4 = load_cols4(e.r1_w, col, e.N);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 2) + 8 * i;
    const int64_t row = row0 + r;
    if (row >= e.M) continue;
    float4 x = lds128(smem_u32(stg + r * STG_LD + 4 * (q4 ^ ((r >> 1) & 3))));
    if (e.amax_a != nullptr) { x.x *= st.sc; x.y *= st.sc; x.z *= st.sc; x.w *= st.sc; }
    float* dst = e.C + row * e.ldc + col;
    if (EPI == EPI_FWD) {
      x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
      x.x = act_fwd(x.x, ACT, e.act_param); x.y = act_fwd(x.y, ACT, e.act_param); x.z = act_fwd(x.z, ACT, e.act_param);
      x.w = act_fwd(x.w, ACT, e.act_param);
      if (e.head_w) {
        hacc[i] = fmaf(x.x, h4.x, fmaf(x.y, h4.y, fmaf(x.z, h4.z, fmaf(x.w, h4.w, hacc[i]))));
        if (e.C == nullptr) continue;
      }
    } else if (EPI == EPI_DGRAD) {
      if (e.r1_d) {
        const float d = yp.d[i];
        x.x = fmaf(d, r4.x, x.x); x.y = fmaf(d, r4.y, x.y); x.z = fmaf(d, r4.z, x.z); x.w = fmaf(d, r4.w, x.w);
      }
      if (ACT != MMSB_ACT_NONE) {
        float4 y = yp.v[i];
        y.x = act_bwd_from_y(y.x, ACT, e.act_prev_param); y.y = act_bwd_from_y(y.y, ACT, e.act_prev_param);
        y.z = act_bwd_from_y(y.z, ACT, e.act_prev_param); y.w = act_bwd_from_y(y.w, ACT, e.act_prev_param);
        x.x *= y.x; x.y *= y.y; x.z *= y.z; x.w *= y.w;
      }
    }
    if (EPI != EPI_ATOMIC && e.amax_out != nullptr) {
      st.amax = fmaxf(st.amax, fabsf(x.x));
      if (col + 1 < e.N) st.amax = fmaxf(st.amax, fabsf(x.y));
      if (col + 2 < e.N) st.amax = fmaxf(st.amax, fabsf(x.z));
      if (col + 3 < e.N) st.amax = fmaxf(st.amax, fabsf(x.w));
    }
    if (EPI == EPI_ATOMIC) {
      const int64_t cs = e.cs ? e.cs : 1;
      dst = e.C + row * e.ldc + col * cs;
      atomicAdd(dst, x.x);
      if (col + 1 < e.N) atomicAdd(dst + cs, x.y);
      if (col + 2 < e.N) atomicAdd(dst + 2 * cs, x.z);
      if (col + 3 < e.N) atomicAdd(dst + 3 * cs, x.w);
    } else if (vec_ok && full) {
      if (EPI == EPI_DGRAD && e.accumulate) {
        const float4 c = *reinterpret_cast<const float4*>(dst);
        x.x += c.x; x.y += c.y; x.z += c.z; x.w += c.w;
      }
      *reinterpret_cast<float4*>(dst) = x;
    } else {
      const bool acc = EPI == EPI_DGRAD && e.accumulate;
      dst[0] = acc ? dst[0] + x.x : x.x;
      if (col + 1 < e.N) dst[1] = acc ? dst[1] + x.y : x.y;
      if (col + 2 < e.N) dst[2] = acc ? dst[2] + x.z : x.z;
      if (col + 3 < e.N) dst[3] = acc ? dst[3] + x.w : x.w;
    }
  }
}

// One chunk, part 1: lane owns row (row0 + lane) of the accumulator in registers -> transpose buffer.  The registers are
// free again afterwards (the TMEM read of the warp's next chunk is issued into them while this chunk is processed).
__device__ __forceinline__ void epilogue_stage(const uint32_t (&v)[CH], float* stg, int lane) {
#pragma unroll
  for (int j = 0; j < CH / 4; ++j)
    sts128(smem_u32(stg + lane * STG_LD + 4 * (j ^ ((lane >> 1) & 3))),
           make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                       __uint_as_float(v[4 * j + 3])));
}
// Part 2: transpose buffer -> bias / activation / derivative -> global rows.
// interior chunk (warp-uniform test); Sigmoid (output layers only, a few columns wide) always takes the edge path
template <int EPI>
__device__ __forceinline__ bool chunk_is_interior(const EpiArgs& e, int64_t row0, int col0, bool vec_ok, int act) {
  return EPI != EPI_ATOMIC && vec_ok && row0 + 32 <= e.M && col0 + CH <= e.N && act != MMSB_ACT_SIGMOID;
}
template <int EPI>
__device__ __forceinline__ int chunk_act(const EpiArgs& e) {
  return EPI == EPI_FWD ? e.act : (EPI == EPI_DGRAD && e.yprev ? e.act_prev : MMSB_ACT_NONE);
}
template <int EPI, bool RICH, typename Next>
__device__ __forceinline__ void epilogue_chunk_rows(const EpiArgs& e, float* stg, const YPrev& yp, const ColVecs& cv,
                                                    bool interior, int act, int lane, int64_t row0, int col0, bool vec_ok,
                                                    EpiState& st, Next issue_next) {
  __syncwarp();
  if (interior) {
    switch (act) {
      case MMSB_ACT_RELU: epilogue_rows_in<EPI, MMSB_ACT_RELU, RICH>(e, stg, yp, cv, lane, row0, col0, st, issue_next); break;
      case MMSB_ACT_SOFTPLUS: epilogue_rows_in<EPI, MMSB_ACT_SOFTPLUS, RICH>(e, stg, yp, cv, lane, row0, col0, st, issue_next); break;
      default: epilogue_rows_in<EPI, MMSB_ACT_NONE, RICH>(e, stg, yp, cv, lane, row0, col0, st, issue_next); break;
    }
  } else {
    issue_next();
    epilogue_rows_edge<EPI>(e, stg, yp, lane, row0, col0, vec_ok, st, act);
  }
  __syncwarp();
}

// Chunk epilogue straight from the fragment layout (EpiArgs::direct == 2): 32 rows x CH columns, no shared memory.
// va / vb: the two tmem_ld_frag results (rows 0..15 / 16..31 of the warp's TMEM lanes); yv: derivative operand (dgrad).
struct YFrag {
  float2 v[2][4];   // [8-column block][row (lane >> 2) + 8 i]
  float d[4];
};

template <int EPI>
__device__ __forceinline__ void load_yfrag(const EpiArgs& e, YFrag& y, int lane, int64_t row0, int col0) {
  if (EPI != EPI_DGRAD) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = row0 + (lane >> 2) + 8 * i;
    if (e.r1_d) y.d[i] = row < e.M ? __ldg(e.r1_d + row) : 0.f;
    if (e.yprev != nullptr && e.act_prev != MMSB_ACT_NONE) {
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        const int col = col0 + 8 * blk + 2 * (lane & 3);
        float2 v = make_float2(0.f, 0.f);
        if (row < e.M && col < e.N) {
          const float* p = e.yprev + row * e.ld_yprev + col;
          if (col + 1 < e.N && (reinterpret_cast<uintptr_t>(p) & 7) == 0) {
            v = __ldg(reinterpret_cast<const float2*>(p));
          } else {
            v.x = __ldg(p);
            if (col + 1 < e.N) v.y = __ldg(p + 1);
          }
        }
        y.v[blk][i] = v;
      }
    }
  }
}

template <int EPI, int ACT>
__device__ __forceinline__ void epilogue_chunk_frag(const EpiArgs& e, const uint32_t (&va)[8], const uint32_t (&vb)[8],
                                                    const YFrag& y, int lane, int64_t row0, int col0, EpiState& st) {
  float (&hacc)[4] = st.hacc;
  const int q4 = lane & 3;
#pragma unroll
  for (int blk = 0; blk < 2; ++blk) {
    const int col = col0 + 8 * blk + 2 * q4;
    if (col >= e.N) continue;
    const bool two = col + 1 < e.N;
    float b0 = 0.f, b1 = 0.f, h0 = 0.f, h1 = 0.f, r0 = 0.f, r1 = 0.f;
    if (EPI == EPI_FWD) {
      if (e.bias) { b0 = __ldg(e.bias + col); if (two) b1 = __ldg(e.bias + col + 1); }
      if (e.head_w) { h0 = __ldg(e.head_w + col); if (two) h1 = __ldg(e.head_w + col + 1); }
    } else if (e.r1_w) {
      r0 = __ldg(e.r1_w + col);
      if (two) r1 = __ldg(e.r1_w + col + 1);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {                      // rows (lane >> 2) + 8 i, like the transposed layout
      const uint32_t* v = i < 2 ? va : vb;
      const int o = 4 * blk + 2 * (i & 1);
      const int64_t row = row0 + (lane >> 2) + 8 * i;
      float x0 = __uint_as_float(v[o]), x1 = __uint_as_float(v[o + 1]);
      if (e.amax_a != nullptr) { x0 *= st.sc; x1 *= st.sc; }
      if (EPI == EPI_FWD) {
        x0 = act_fwd(x0 + b0, ACT, e.act_param);
        x1 = act_fwd(x1 + b1, ACT, e.act_param);
        if (e.head_w) hacc[i] = fmaf(x0, h0, fmaf(x1, h1, hacc[i]));
      } else {
        if (e.r1_d) { x0 = fmaf(y.d[i], r0, x0); x1 = fmaf(y.d[i], r1, x1); }
        if (ACT != MMSB_ACT_NONE) {
          x0 *= act_bwd_from_y(y.v[blk][i].x, ACT, e.act_prev_param);
          x1 *= act_bwd_from_y(y.v[blk][i].y, ACT, e.act_prev_param);
        }
      }
      if (e.C != nullptr && row < e.M) {
        float* dst = e.C + row * e.ldc + col;
        if (e.amax_out != nullptr) st.amax = fmaxf(st.amax, fmaxf(fabsf(x0), two ? fabsf(x1) : 0.f));
        if (EPI == EPI_DGRAD && e.accumulate) {
          x0 += dst[0];
          if (two) x1 += dst[1];
        }
        if (two && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
          *reinterpret_cast<float2*>(dst) = make_float2(x0, x1);
        } else {
          dst[0] = x0;
          if (two) dst[1] = x1;
        }
      }
    }
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_chunk_frag_dispatch(const EpiArgs& e, const uint32_t (&va)[8], const uint32_t (&vb)[8],
                                                             const YFrag& y, int lane, int64_t row0, int col0,
                                                             EpiState& st) {
  const int act = EPI == EPI_FWD ? e.act : (e.yprev ? e.act_prev : MMSB_ACT_NONE);
  switch (act) {
    case MMSB_ACT_RELU: epilogue_chunk_frag<EPI, MMSB_ACT_RELU>(e, va, vb, y, lane, row0, col0, st); break;
    case MMSB_ACT_SOFTPLUS: epilogue_chunk_frag<EPI, MMSB_ACT_SOFTPLUS>(e, va, vb, y, lane, row0, col0, st); break;
    case MMSB_ACT_SIGMOID: epilogue_chunk_frag<EPI, MMSB_ACT_SIGMOID>(e, va, vb, y, lane, row0, col0, st); break;
    default: epilogue_chunk_frag<EPI, MMSB_ACT_NONE>(e, va, vb, y, lane, row0, col0, st); break;
  }
}

// All chunks of one accumulator that belong to this warp (quadrant q = warp % 4, chunks c = warp / 4, + EPI_WARPS / 4, ...).
// `release` is called by every lane right after the warp's last TMEM read (frees the accumulator for the MMA warp).
template <int EPI, bool RICH, typename Release>
__device__ __forceinline__ void epilogue_tile(const EpiArgs& e, uint32_t tmem_acc, int w, float* stg, int warp, int lane,
                                              int64_t row_base, int col_base, bool vec_ok, bool vec_y, YPrev& y_cur,
                                              Release release) {
  constexpr int STEP = EPI_WARPS / 4;
  const int q = warp & 3, first = warp >> 2;
  const int nch = (w + CH - 1) / CH;
  const int64_t row0 = row_base + q * 32;
  EpiState st;
  st.hacc[0] = st.hacc[1] = st.hacc[2] = st.hacc[3] = 0.f;
  st.sc = 1.f;
  st.amax = 0.f;
  if (e.amax_a != nullptr)
    st.sc = pow2f(-f16_scale_exp(__ldg(e.amax_a))) * pow2f(-f16_scale_exp(__ldg(e.amax_b)));
  float (&hacc)[4] = st.hacc;
  auto amax_flush = [&]() {
    if (EPI == EPI_ATOMIC || e.amax_out == nullptr) return;
    float m = st.amax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(e.amax_out), __float_as_uint(m));
  };
  // fused head: the four lanes that share a row combine their partial dot products; one reduction per row and warp
  // (two warps per row -> two commutative additions onto the caller's zeros: order-independent)
  auto head_flush = [&]() {
    if (EPI != EPI_FWD || e.head_w == nullptr) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = hacc[i];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      const int64_t row = row0 + (lane >> 2) + 8 * i;
      if ((lane & 3) == 0 && row < e.M) {
        if (first == 0 && col_base == 0 && e.head_b) v += __ldg(e.head_b);
        atomicAdd(e.head_out + row, v);
      }
    }
  };
  if (first >= nch) {
    release();
    return;
  }
#if MMSB_TC_FRAG_EPILOGUE
  if (EPI != EPI_ATOMIC && e.direct == 2) {
    // fragment-layout epilogue, software-pipelined: the TMEM reads (and the derivative operand) of the warp's next
    // chunk are in flight while the current one is processed
    const uint32_t tbase = tmem_acc + (uint32_t(q * 32) << 16);
    if constexpr (EPI == EPI_FWD) {
      YFrag ynone;
      for (int c = first; c < nch; c += STEP) {
        uint32_t va[8], vb[8];
        tmem_ld_frag(tbase + uint32_t(c * CH), va);
        tmem_ld_frag(tbase + uint32_t(c * CH) + (16u << 16), vb);
        tmem_ld_wait();
        if (c + STEP >= nch) release();
        epilogue_chunk_frag_dispatch<EPI>(e, va, vb, ynone, lane, row0, col_base + c * CH, st);
      }
    } else {
      // dgrad: the derivative operand of the next chunk is what is kept in flight (registers do not allow both)
      uint32_t va[8], vb[8];
      YFrag ycur, ynext;
      load_yfrag<EPI>(e, ycur, lane, row0, col_base + first * CH);
      for (int c = first; c < nch; c += STEP) {
        const int cn = c + STEP;
        tmem_ld_frag(tbase + uint32_t(c * CH), va);
        tmem_ld_frag(tbase + uint32_t(c * CH) + (16u << 16), vb);
        if (cn < nch) load_yfrag<EPI>(e, ynext, lane, row0, col_base + cn * CH);
        tmem_ld_wait();
        if (cn >= nch) release();
        epilogue_chunk_frag_dispatch<EPI>(e, va, vb, ycur, lane, row0, col_base + c * CH, st);
        if (cn < nch) ycur = ynext;
      }
    }
    if (EPI == EPI_FWD && e.head_w != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = hacc[i];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        const int64_t row = row0 + (lane >> 2) + 8 * i;
        if ((lane & 3) == 0 && row < e.M) {
          if (first == 0 && col_base == 0 && e.head_b) v += __ldg(e.head_b);
          atomicAdd(e.head_out + row, v);
        }
      }
    }
    amax_flush();
    return;
  }
#endif
  // software pipeline over the warp's chunks with ONE register array: chunk c is moved to the transpose buffer, then the
  // TMEM read of chunk c + STEP (and, dgrad, its derivative operand) is issued and stays in flight while chunk c is
  // processed from shared memory.  One call site, loop not unrolled: the code of the activation variants exists once.
  // (dgrad: the derivative operand of the next chunk is what is kept in flight, the registers do not allow both)
  constexpr bool LD_AHEAD = EPI != EPI_DGRAD || RICH;
  uint32_t v[CH];
  if (LD_AHEAD) tmem_ld16(tmem_acc + uint32_t(first * CH) + (uint32_t(q * 32) << 16), v);
  const int act = chunk_act<EPI>(e);
#pragma unroll 1
  for (int c = first; c < nch; c += STEP) {
    const int cn = c + STEP;
    if (!LD_AHEAD) tmem_ld16(tmem_acc + uint32_t(c * CH) + (uint32_t(q * 32) << 16), v);
    const bool interior = chunk_is_interior<EPI>(e, row0, col_base + c * CH, vec_ok, act);
    ColVecs cv;
    if (interior && (EPI != EPI_DGRAD || RICH)) load_colvecs<EPI>(e, cv, lane, col_base + c * CH);
    tmem_ld_wait();
    epilogue_stage(v, stg, lane);
    if (cn >= nch) release();
    if (interior && EPI == EPI_DGRAD && !RICH) load_colvecs<EPI>(e, cv, lane, col_base + c * CH);   // (no registers to spare earlier)
    YPrev y_next;
    if (cn < nch) load_yprev<EPI>(e, y_next, lane, row0, col_base + cn * CH, vec_y);
    epilogue_chunk_rows<EPI, RICH>(e, stg, y_cur, cv, interior, act, lane, row0, col_base + c * CH, vec_ok, st, [&]() {
      if (LD_AHEAD && cn < nch) tmem_ld16(tmem_acc + uint32_t(cn * CH) + (uint32_t(q * 32) << 16), v);
    });
    if (EPI == EPI_DGRAD && cn < nch) y_cur = y_next;
  }
  head_flush();
  amax_flush();
}

// Converter step of one landed A k-block (TMA path): this thread's four 16-byte chunks.  All shared-memory reads are
// issued before the first dependent instruction (the accesses are volatile asm: the compiler keeps their order), so a
// k-block costs one shared-memory latency instead of four.
// generated operand of the products below a fused head: hd[row] * hw[col] * act'(y), 4 columns; the activation is
// dispatched once per 16 bytes, not per element
template <int ACT>
__device__ __forceinline__ float4 gen_dz4(const float4& y, float d, const float4& hw4, float p) {
  const float4 a = act_bwd4<ACT>(y, p);
  return make_float4(d * hw4.x * a.x, d * hw4.y * a.y, d * hw4.z * a.z, d * hw4.w * a.w);
}
__device__ __forceinline__ float4 gen_dz4_rt(const float4& y, float d, const float4& hw4, int act, float p) {
  switch (act) {
    case MMSB_ACT_SOFTPLUS: return gen_dz4<MMSB_ACT_SOFTPLUS>(y, d, hw4, p);
    case MMSB_ACT_RELU: return gen_dz4<MMSB_ACT_RELU>(y, d, hw4, p);
    case MMSB_ACT_SIGMOID: return gen_dz4<MMSB_ACT_SIGMOID>(y, d, hw4, p);
    default: return gen_dz4<MMSB_ACT_NONE>(y, d, hw4, p);
  }
}
template <int ACT>
__device__ __forceinline__ void generate_rows(float4 (&x)[4], uint32_t hi, const uint32_t (&off)[4], const float (&hd)[4],
                                              const float4& hw4, float hact_param) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 d = act_bwd4<ACT>(x[i], hact_param);
    x[i].x = hd[i] * hw4.x * d.x;
    x[i].y = hd[i] * hw4.y * d.y;
    x[i].z = hd[i] * hw4.z * d.z;
    x[i].w = hd[i] * hw4.w * d.w;
    sts128(hi + off[i], x[i]);
  }
}
template <int NPARTS>
__device__ __forceinline__ void convert_block(uint32_t hi, uint32_t lo, const uint32_t (&off)[4], bool gen, const float (&hd)[4],
                                              const float4& hw4, int hact, float hact_param) {
  float4 x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = lds128(hi + off[i]);
  if (gen) {      // activation chosen once per k-block (warp-uniform), not per element
    switch (hact) {
      case MMSB_ACT_SOFTPLUS: generate_rows<MMSB_ACT_SOFTPLUS>(x, hi, off, hd, hw4, hact_param); break;
      case MMSB_ACT_RELU: generate_rows<MMSB_ACT_RELU>(x, hi, off, hd, hw4, hact_param); break;
      case MMSB_ACT_SIGMOID: generate_rows<MMSB_ACT_SIGMOID>(x, hi, off, hd, hw4, hact_param); break;
      default: generate_rows<MMSB_ACT_NONE>(x, hi, off, hd, hw4, hact_param); break;
    }
  }
  if (NPARTS == 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 l;
      l.x = tf32_rna(x[i].x - __uint_as_float(__float_as_uint(x[i].x) & 0xFFFFE000u));
      l.y = tf32_rna(x[i].y - __uint_as_float(__float_as_uint(x[i].y) & 0xFFFFE000u));
      l.z = tf32_rna(x[i].z - __uint_as_float(__float_as_uint(x[i].z) & 0xFFFFE000u));
      l.w = tf32_rna(x[i].w - __uint_as_float(__float_as_uint(x[i].w) & 0xFFFFE000u));
      sts128(lo + off[i], l);
    }
  }
}

// ---- forward / dgrad ----------------------------------------------------------------------------------------
struct RowsArgs {
  const float* A; int64_t lda; int64_t M; int K;
  const float* Bp; int N;
  EpiArgs epi;
  int n_tiles; int nkb; int64_t total_tiles;
  int dbg;   // dev only (MMSB_TC_DEBUG): 1 = no A loads, 2 = no output stores, 4 = no B copies, 8 = no MMAs
  // generated operand (dgrad below a fused head): A[r, k] = hd[r] * hw[k] * act'(A_mem[r, k]), A_mem = the stored
  // activations of the layer
  const float* hd; const float* hw; int hact; float hact_param;
  int stages, stage_bytes;   // ring geometry of this launch (set by launch_rows)
  int prefetch;              // TMA path: L2 prefetch of the A boxes one tile ahead (MMSB_TC_PREFETCH=1; measured: no gain, off)
};

// TMA_A: the A k-blocks are landed by tensor-map copies straight into the stage's "hi" tile (K-major SWIZZLE_128B); the
// tensor core reads fp32 operands as TF32 by TRUNCATION (measured, scripts/dev_tf32_rounding.py), so the raw tile IS the
// hi operand, hi = trunc(x), and the 8 converter warps only derive lo = rna(x - trunc(x)) from shared memory.  No global
// load sits in a register (or on the hand-shake path) any more.  Needs 16-byte aligned rows of A and a stored operand;
// otherwise (generated operand, odd leading dimension) the register-staged producers below are used.
template <int NPARTS, int EPI, bool TMA_A>
__global__ void __launch_bounds__(THREADS, 1) tc_rows_kernel(const __grid_constant__ RowsArgs g, const __grid_constant__ CUtensorMap tmap_a) {
  // stages: as many as fit into the ring (2 x 96 KB) for this launch's widest B tile — 2 for a 256-wide accumulator,
  // 3..4 for narrow layers (their MMAs are short, so the stage refill latency is what a deeper ring hides)
  constexpr int RING = num_stages(NPARTS) * stage_bytes(NPARTS);
  const int STAGE = g.stage_bytes;
  const int S = g.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* stg_all = reinterpret_cast<float*>(smem + RING);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING + STG_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + MAX_STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 2 * MAX_STAGES), bar_tempty = smem_u32(bars + 2 * MAX_STAGES + 2);
  const uint32_t bar_raw = smem_u32(bars + 2 * MAX_STAGES + 4);      // TMA_A: raw A tile landed

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  if (t == 0) {
    for (int s = 0; s < S; ++s) {
      // register path: the four warps of one producer group + the loader's expect_tx;
      // TMA path: the 8 converter warps + the loader (3xTF32) or the loader alone (TF32: nothing to convert)
      mbar_init(bar_full + 8 * s, TMA_A ? ((NPARTS == 2 || g.hd != nullptr) ? PROD_WARPS + 1 : 1) : 4 + 1);
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_raw + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_pad = pad16(g.N);
  const int last_ksteps = (g.K - (g.nkb - 1) * TK + 7) / 8;

  if (warp < EPI_WARPS) {
    // ================= epilogue: TMEM -> registers -> global =================
    float* stg = stg_all + warp * 32 * STG_LD;
    const bool vec_ok = ((g.epi.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.epi.C) & 15) == 0);
    const bool vec_y = ((g.epi.ld_yprev & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.epi.yprev) & 15) == 0);
    uint32_t ti = 0;
    for (int64_t tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++ti) {
      const int64_t mt = tile / g.n_tiles;
      const int nt = int(tile - mt * g.n_tiles);
      const int w = tile_width(n_pad, nt);
      const uint32_t acc = ti & 1;
      if (EPI == EPI_DGRAD && g.epi.yprev != nullptr && g.epi.act_prev != MMSB_ACT_NONE) {
        // the tile's derivative operand (128 rows x w floats) is pulled into the L2 while the tile's MMAs still run
        const int lines_per_row = (w * 4 + 127) / 128;
        for (int l = t; l < TM * lines_per_row; l += EPI_WARPS * 32) {
          const int64_t row = mt * TM + l / lines_per_row;
          const int col = nt * NT + (l % lines_per_row) * 32;
          if (row < g.epi.M && col < g.epi.N)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(g.epi.yprev + row * g.epi.ld_yprev + col));
        }
      }
      // the derivative operand of this warp's first chunk is requested before the wait on the accumulator
      YPrev y_cur;
      load_yprev<EPI>(g.epi, y_cur, lane, mt * TM + (warp & 3) * 32, nt * NT + (warp >> 2) * CH, vec_y);
      mbar_wait(bar_tfull + 8 * acc, (ti >> 1) & 1);
      tc_fence_after();
      auto release = [&]() {
        // one arrival per warp: 256 single-thread arrivals on one mbarrier serialise in the shared-memory atomics
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      };
      if (g.dbg & 2) release();   // dev: no epilogue work
      else epilogue_tile<EPI, false>(g.epi, tmem + acc * NT, w, stg, warp, lane, mt * TM, nt * NT, vec_ok, vec_y, y_cur, release);
    }
  } else if (warp < EPI_WARPS + PROD_WARPS) {
    if constexpr (TMA_A) {
      // ================= converters: lo = rna(x - trunc(x)) of the landed raw tile =================
      // (generated operand, g.hd: the landed tile holds the stored activations y; it is overwritten in place with
      //  a = hd[row] * hw[k] * act'(y), which the tensor core again reads truncated, and lo is derived from a)
      if (NPARTS == 2 || g.hd != nullptr) {
        const int p = t - EPI_WARPS * 32;
        const int64_t my_tiles = blockIdx.x < g.total_tiles ? (g.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int64_t iters = my_tiles * g.nkb;
        const int c = p & 7, r_base = p >> 3;          // generated operand: 16-byte chunk c of rows r_base + 32 i
        for (int64_t it = 0; it < iters; ++it) {
          const uint32_t s = uint32_t(it % S);
          float hd[4];
          float4 hw4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g.hd) {
            const int64_t tl = it / g.nkb;
            const int kb = int(it - tl * g.nkb);
            const int64_t m0 = ((blockIdx.x + tl * gridDim.x) / g.n_tiles) * TM;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int64_t row = m0 + r_base + 32 * i;
              hd[i] = row < g.M ? __ldg(g.hd + row) : 0.f;
            }
            if (kb * TK + c * 4 < g.K) hw4 = load_cols4(g.hw, kb * TK + c * 4, g.K);
          }
          mbar_wait(bar_raw + 8 * s, uint32_t(it / S) & 1);
          const uint32_t hi = smem_u32(smem + s * STAGE), lo = hi + PART;
          static_assert(PART / (PROD_THREADS * 16) == 4, "four chunks per converter thread");
          uint32_t off[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            off[i] = uint32_t(p + i * PROD_THREADS) * 16u;     // plain conversion is element-wise: any mapping works
            if (g.hd) {
              const int r = r_base + 32 * i;
              off[i] = uint32_t(r) * 128u + (uint32_t(c ^ (r & 7)) << 4);
            }
          }
          convert_block<NPARTS>(hi, lo, off, g.hd != nullptr, hd, hw4, g.hact, g.hact_param);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_full + 8 * s);
        }
      }
    } else {
    // ================= producers: A (global fp32) -> hi/lo -> swizzled shared memory =================
    // Two groups of four warps alternate over the k-blocks, each group owning one of the S = 2 stages (its own
    // mbarrier phases in order): a group's loads for its next k-block are in flight while the other group stores
    // and the MMAs of its previous block run.  (A register ring inside one warp does not pipeline: the loads of all
    // ring slots share scoreboard entries, so the wait for the oldest slot also waits for the newest.  More groups
    // than stages would alias the mbarrier phase parities.)
    constexpr int GROUPS = PROD_WARPS / 4;
    static_assert(GROUPS == 2, "two producer groups");      // the host picks an even stage count for this path
    const int p = t - EPI_WARPS * 32;
    const int grp = p >> 7, q = p & 127;
    const int c = q & 7, r_base = q >> 3;          // 16-byte chunk c of rows r_base + 16 i
    const bool vec = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
    const int64_t my_tiles = blockIdx.x < g.total_tiles ? (g.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t iters = my_tiles * g.nkb;
    float4 v[8];
    float hd[8];
    float4 hw4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_block = [&](int64_t it) {
      const int64_t tl = it / g.nkb;
      const int kb = int(it - tl * g.nkb);
      const int64_t m0 = ((blockIdx.x + tl * gridDim.x) / g.n_tiles) * TM;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        v[i] = (g.dbg & 1) ? make_float4(1.f, 2.f, 3.f, 4.f)
                           : load4(g.A, g.lda, m0 + r_base + 16 * i, g.M, kb * TK + c * 4, g.K, vec);
      if (g.hd) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t row = m0 + r_base + 16 * i;
          hd[i] = row < g.M ? __ldg(g.hd + row) : 0.f;
        }
        hw4 = kb * TK + c * 4 < g.K ? load_cols4(g.hw, kb * TK + c * 4, g.K) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto generate = [&](float4 y, float d) { return gen_dz4_rt(y, d, hw4, g.hact, g.hact_param); };
    int64_t it = grp;
    if (it < iters) load_block(it);
    while (it < iters) {
      const uint32_t s = uint32_t(it % S);
      mbar_wait(bar_empty + 8 * s, (uint32_t(it / S) & 1) ^ 1);
      uint8_t* a_hi = smem + s * STAGE;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = r_base + 16 * i;
        const uint32_t off = uint32_t(r) * 128u + (uint32_t(c ^ (r & 7)) << 4);
        if (g.dbg & 32) {   // dev: no hi/lo conversion
          sts128(smem_u32(a_hi) + off, v[i]);
          if (NPARTS == 2) sts128(smem_u32(a_hi + PART) + off, v[i]);
        } else {
          split_store4<NPARTS>(a_hi, a_hi + PART, off, g.hd ? generate(v[i], hd[i]) : v[i]);
        }
      }
      if (!(g.dbg & 16)) fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);     // one arrival per warp (128 single arrivals would serialise)
      it += GROUPS;
      if (it < iters) load_block(it);
    }
    }
  } else if (warp == EPI_WARPS + PROD_WARPS) {
    // ================= B loader: one bulk async copy per k-block =================
    if (lane == 0) {
      const bool pf = g.prefetch != 0;
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        const int nt = int(tile % g.n_tiles);
        const int w = tile_width(n_pad, nt);
        const uint32_t bytes = uint32_t(NPARTS * w * 128);
        const float* src = g.Bp + int64_t(nt) * NT * TK * g.nkb * NPARTS;
        const int m0 = int((tile / g.n_tiles) * TM);
        for (int kb = 0; kb < g.nkb; ++kb, ++it) {
          const uint32_t s = it % S;
          mbar_wait(bar_empty + 8 * s, ((it / S) & 1) ^ 1);
          if constexpr (TMA_A) {
            // raw A tile: 128 rows x 128 B (rows / columns past the tensor are zero-filled and still counted)
            const bool conv = NPARTS == 2 || g.hd != nullptr;
            const uint32_t a_bar = conv ? bar_raw + 8 * s : bar_full + 8 * s;
            if (conv) mbar_arrive_expect_tx(a_bar, uint32_t(PART));
            else mbar_arrive_expect_tx(a_bar, uint32_t(PART) + ((g.dbg & 4) ? 0u : bytes));
            tma_load_2d(smem_u32(smem + s * STAGE), &tmap_a, kb * TK, m0, a_bar);
            if (pf && tile + gridDim.x < g.total_tiles)
              tma_prefetch_2d(&tmap_a, kb * TK, int(((tile + gridDim.x) / g.n_tiles) * TM));
            if (!conv) {
              if (!(g.dbg & 4))
                bulk_g2s(smem_u32(smem + s * STAGE + NPARTS * PART), src + int64_t(kb) * NPARTS * w * TK, bytes, a_bar);
              continue;
            }
          }
          if (g.dbg & 4) {
            mbar_arrive(bar_full + 8 * s);
          } else {
            mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
            bulk_g2s(smem_u32(smem + s * STAGE + NPARTS * PART), src + int64_t(kb) * NPARTS * w * TK, bytes, bar_full + 8 * s);
          }
        }
      }
    }
  } else {
    // ================= MMA issuer =================
    if (lane == 0) {
      uint32_t it = 0, ti = 0;
      for (int64_t tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++ti) {
        const int nt = int(tile % g.n_tiles);
        const int w = tile_width(n_pad, nt);
        const uint32_t idesc = make_idesc_tf32(w, false);
        const uint32_t acc = ti & 1;
        mbar_wait_spin(bar_tempty + 8 * acc, ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem + acc * NT;
        for (int kb = 0; kb < g.nkb; ++kb, ++it) {
          const uint32_t s = it % S;
          mbar_wait_spin(bar_full + 8 * s, (it / S) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + s * STAGE);
          const uint32_t b_hi = a_hi + NPARTS * PART;
          const uint64_t dah = make_desc(a_hi, 16, 1024), dal = make_desc(a_hi + PART, 16, 1024);
          const uint64_t dbh = make_desc(b_hi, 16, 1024), dbl = make_desc(b_hi + w * 128, 16, 1024);
          const int ksteps = (g.dbg & 8) ? 0 : (kb == g.nkb - 1 ? last_ksteps : TK / 8);
          for (int j = 0; j < ksteps; ++j) {
            const uint64_t adv = uint64_t(j * 2);   // 8 tf32 = 32 bytes along K inside the swizzle row
            if (NPARTS == 2) {
              umma_tf32(d, dal + adv, dbh + adv, idesc, (kb | j) ? 1u : 0u);
              umma_tf32(d, dah + adv, dbl + adv, idesc, 1u);
              umma_tf32(d, dah + adv, dbh + adv, idesc, 1u);
            } else {
              umma_tf32(d, dah + adv, dbh + adv, idesc, (kb | j) ? 1u : 0u);
            }
          }
          umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_tfull + 8 * acc);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}


// ---- forward / dgrad on CTA pairs (cta_group::2) ------------------------------------------------------------------
// Two CTAs of a cluster (two SMs of one TPC) share one 256-row x 256-column tile: each CTA lands and converts the A
// k-blocks of its own 128 rows, loads ITS HALF of the weight k-block (128 of the 256 rows of W: 32 KB instead of 64 KB),
// and the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), which reads both CTAs' operand tiles; each CTA's TMEM
// holds the accumulator of its own 128 rows and its own epilogue warps drain it.  A stage is 64 KB instead of 96 KB, so the
// ring holds THREE stages: the stage refill latency (~1.9 us against 0.8 us of MMAs per k-block) is what limited the
// single-CTA kernel with two.  Barriers the MMA thread waits on live in the leader (rank 0): the peer's converter and
// epilogue warps arrive remotely (shared::cluster address with the peer bit cleared), both CTAs' weight copies
// (tensor-map copies with .cta_group::2) complete their bytes on the leader's barrier, and the MMA completions are
// committed with a multicast to both CTAs' barriers.
constexpr int PAIR_STAGES = 3;
constexpr int PAIR_STAGE = 2 * PART + 2 * (128 * 128);

// KIND 0: 3xTF32 (k-blocks of 32 fp32; the landed tile is the hi operand, the converters derive lo).
// KIND 1: 2-term fp16 split (k-blocks of 64: TWO raw fp32 boxes land in the stage's A region and are converted IN PLACE
//         into the fp16 hi tile [0, PART) and lo tile [PART, 2 PART): a warp owns whole 8-row swizzle atoms of both boxes,
//         reads its 2 KB into registers, __syncwarp, writes the 1 KB hi + 1 KB lo atoms over the same bytes).
template <int EPI, int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
    tc_rows_pair_kernel(const __grid_constant__ RowsArgs g, const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b) {
  constexpr int KB = KIND == 1 ? TK16 : TK;           // K elements per k-block
  constexpr int KSTEP = KIND == 1 ? 16 : 8;           // K elements per MMA
  constexpr int S = PAIR_STAGES;
  constexpr int STAGE = PAIR_STAGE;
  constexpr int HB = 128 * 128;                       // bytes of one half (128 rows) of a B part
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* stg_all = reinterpret_cast<float*>(smem + S * STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * STAGE + STG_BYTES);
  // barriers (same offsets in both CTAs): full[4] empty[4] tfull[2] tempty[2] raw[4]
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + MAX_STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 2 * MAX_STAGES), bar_tempty = smem_u32(bars + 2 * MAX_STAGES + 2);
  const uint32_t bar_raw = smem_u32(bars + 2 * MAX_STAGES + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int64_t m_ptiles = (g.M + 2 * TM - 1) / (2 * TM);
  const int64_t total_ptiles = m_ptiles * g.n_tiles;
  if (t == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 2 * PAIR_PROD_WARPS + 1);   // both CTAs' converter warps + the leader loader's expect_tx
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_raw + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 2 * EPI_WARPS);      // both CTAs' epilogue warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int last_ksteps = (g.K - (g.nkb - 1) * KB + KSTEP - 1) / KSTEP;
  const int64_t my_ptiles = pair < total_ptiles ? (total_ptiles - pair + npairs - 1) / npairs : 0;

  if (warp < EPI_WARPS) {
    // ================= epilogue (this CTA's 128 rows) =================
    float* stg = stg_all + warp * 32 * STG_LD;
    const bool vec_ok = ((g.epi.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.epi.C) & 15) == 0);
    const bool vec_y = ((g.epi.ld_yprev & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.epi.yprev) & 15) == 0);
    uint32_t ti = 0;
    for (int64_t pt = pair; pt < total_ptiles; pt += npairs, ++ti) {
      const int64_t mt = pt / g.n_tiles;
      const int nt = int(pt - mt * g.n_tiles);
      const int64_t m0 = mt * 2 * TM + rank * TM;
      const uint32_t acc = ti & 1;
      if (EPI == EPI_DGRAD && g.epi.yprev != nullptr && g.epi.act_prev != MMSB_ACT_NONE) {
        for (int l = t; l < TM * 8; l += EPI_WARPS * 32) {
          const int64_t row = m0 + l / 8;
          const int col = nt * NT + (l % 8) * 32;
          if (row < g.epi.M && col < g.epi.N)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(g.epi.yprev + row * g.epi.ld_yprev + col));
        }
      }
      YPrev y_cur;
      load_yprev<EPI>(g.epi, y_cur, lane, m0 + (warp & 3) * 32, nt * NT + (warp >> 2) * CH, vec_y);
      mbar_wait(bar_tfull + 8 * acc, (ti >> 1) & 1);
      tc_fence_after();
      auto release = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(bar_tempty + 8 * acc);
      };
      if (g.dbg & 2) release();   // dev: no epilogue work
      else epilogue_tile<EPI, (PAIR_THREADS <= 512)>(g.epi, tmem + acc * NT, NT, stg, warp, lane, m0, nt * NT, vec_ok, vec_y, y_cur, release);
    }
  } else if (warp < EPI_WARPS + PAIR_PROD_WARPS) {
    // ================= converters (this CTA's 128 rows of A) =================
    const int p = t - EPI_WARPS * 32;
    const int64_t iters = my_ptiles * g.nkb;
    const int c = p & 7, r_base = p >> 3;
    constexpr int NCH = PART / (PAIR_PROD_THREADS * 16);      // 16-byte chunks per converter thread and k-block (4 or 8)
    constexpr int RSTEP = PAIR_PROD_THREADS / 8;              // generated operand: rows r_base + RSTEP i
    static_assert(NCH % 4 == 0, "whole groups of four chunks");
    for (int64_t it = 0; it < iters; ++it) {
      const uint32_t s = uint32_t(it % S);
      float hd[NCH];
      float4 hw4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g.hd) {
        const int64_t tl = it / g.nkb;
        const int kb = int(it - tl * g.nkb);
        const int64_t m0 = ((pair + tl * npairs) / g.n_tiles) * 2 * TM + rank * TM;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int64_t row = m0 + r_base + RSTEP * i;
          hd[i] = row < g.M ? __ldg(g.hd + row) : 0.f;
        }
        if (kb * TK + c * 4 < g.K) hw4 = load_cols4(g.hw, kb * TK + c * 4, g.K);
      }
      mbar_wait(bar_raw + 8 * s, uint32_t(it / S) & 1);
      const uint32_t hi = smem_u32(smem + s * STAGE), lo = hi + PART;
      if constexpr (KIND == 1) {
        // lane (r = lane & 7, j = lane >> 3) of a warp's 8-row atom: raw chunks 2j, 2j+1 of both boxes -> hi / lo chunks
        // j (k 8j..8j+7) and j + 4 (k 32+8j..); chunk positions are XOR-swizzled with the row (SWIZZLE_128B)
        const float sa = pow2f(f16_scale_exp(__ldg(g.epi.amax_a)));
        const int r = lane & 7, j = lane >> 3;
        const int cw = warp - EPI_WARPS;
#pragma unroll 1
        for (int grp = cw; grp < TM / 8; grp += PAIR_PROD_WARPS) {
          const uint32_t rowb = uint32_t(grp * 8 + r) * 128u;
          const uint32_t q0 = rowb + (uint32_t((2 * j) ^ r) << 4), q1 = rowb + (uint32_t((2 * j + 1) ^ r) << 4);
          const float4 a0 = lds128(hi + q0), a1 = lds128(hi + q1), b0 = lds128(lo + q0), b1 = lds128(lo + q1);
          __syncwarp();
          uint32_t h[4], l[4];
          split_f16x2(a0.x * sa, a0.y * sa, h[0], l[0]); split_f16x2(a0.z * sa, a0.w * sa, h[1], l[1]);
          split_f16x2(a1.x * sa, a1.y * sa, h[2], l[2]); split_f16x2(a1.z * sa, a1.w * sa, h[3], l[3]);
          const uint32_t c0 = rowb + (uint32_t(j ^ r) << 4), c1 = rowb + (uint32_t((j + 4) ^ r) << 4);
          sts128u(hi + c0, h[0], h[1], h[2], h[3]);
          sts128u(lo + c0, l[0], l[1], l[2], l[3]);
          split_f16x2(b0.x * sa, b0.y * sa, h[0], l[0]); split_f16x2(b0.z * sa, b0.w * sa, h[1], l[1]);
          split_f16x2(b1.x * sa, b1.y * sa, h[2], l[2]); split_f16x2(b1.z * sa, b1.w * sa, h[3], l[3]);
          sts128u(hi + c1, h[0], h[1], h[2], h[3]);
          sts128u(lo + c1, l[0], l[1], l[2], l[3]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(bar_full + 8 * s);
        continue;
      }
#pragma unroll
      for (int h = 0; h < NCH / 4; ++h) {
        uint32_t off[4];
        float hd4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ii = 4 * h + i;
          off[i] = uint32_t(p + ii * PAIR_PROD_THREADS) * 16u;
          if (g.hd) {
            const int r = r_base + RSTEP * ii;
            off[i] = uint32_t(r) * 128u + (uint32_t(c ^ (r & 7)) << 4);
          }
          hd4[i] = g.hd ? hd[ii] : 0.f;
        }
        convert_block<2>(hi, lo, off, g.hd != nullptr, hd4, hw4, g.hact, g.hact_param);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar_full + 8 * s);
    }
  } else if (warp == EPI_WARPS + PAIR_PROD_WARPS) {
    // ================= loader: this CTA's A k-block and its half of the weight k-block =================
    if (lane == 0) {
      const bool pf = g.prefetch != 0;
      uint32_t it = 0;
      for (int64_t pt = pair; pt < total_ptiles; pt += npairs) {
        const int64_t mt = pt / g.n_tiles;
        const int nt = int(pt - mt * g.n_tiles);
        const int m0 = int(mt * 2 * TM + rank * TM);
        // rows of the packed weight buffer (32 floats each): n-tile nt, k-block kb = [hi 256 rows][lo 256 rows]
        const int64_t tile_row0 = int64_t(nt) * NT * g.nkb * 2;
        for (int kb = 0; kb < g.nkb; ++kb, ++it) {
          const uint32_t s = it % S;
          mbar_wait(bar_empty + 8 * s, ((it / S) & 1) ^ 1);
          const uint32_t a_hi = smem_u32(smem + s * STAGE), b_hi = a_hi + 2 * PART;
          if constexpr (KIND == 1) {
            // two raw fp32 boxes of 32 k each (columns past K are zero-filled by the copy)
            mbar_arrive_expect_tx(bar_raw + 8 * s, uint32_t(2 * PART));
            tma_load_2d(a_hi, &tmap_a, kb * KB, m0, bar_raw + 8 * s);
            tma_load_2d(a_hi + PART, &tmap_a, kb * KB + TK, m0, bar_raw + 8 * s);
          } else {
            mbar_arrive_expect_tx(bar_raw + 8 * s, uint32_t(PART));
            tma_load_2d(a_hi, &tmap_a, kb * TK, m0, bar_raw + 8 * s);
          }
          if (KIND == 0 && pf && pt + npairs < total_ptiles)
            tma_prefetch_2d(&tmap_a, kb * TK, int(((pt + npairs) / g.n_tiles) * 2 * TM + rank * TM));
          if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * s, 4u * HB);     // both CTAs' hi + lo halves
          const int row_hi = int(tile_row0 + int64_t(kb) * 2 * NT + rank * 128);
          tma_load_2d_pair(b_hi, &tmap_b, 0, row_hi, bar_full + 8 * s);
          tma_load_2d_pair(b_hi + HB, &tmap_b, 0, row_hi + NT, bar_full + 8 * s);
        }
      }
    }
  } else if (warp == EPI_WARPS + PAIR_PROD_WARPS + 1 && rank == 0) {
    // ================= MMA issuer (leader CTA) =================
    if (lane == 0) {
      const uint32_t idesc = KIND == 1 ? make_idesc_f16(NT, 2 * TM)
                                       : ((1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(NT >> 3) << 17) | (uint32_t((2 * TM) >> 4) << 24));
      uint32_t it = 0, ti = 0;
      for (int64_t pt = pair; pt < total_ptiles; pt += npairs, ++ti) {
        const uint32_t acc = ti & 1;
        mbar_wait_spin(bar_tempty + 8 * acc, ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem + acc * NT;
        for (int kb = 0; kb < g.nkb; ++kb, ++it) {
          const uint32_t s = it % S;
          mbar_wait_spin(bar_full + 8 * s, (it / S) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + s * STAGE), b_hi = a_hi + 2 * PART;
          const uint64_t dah = make_desc(a_hi, 16, 1024), dal = make_desc(a_hi + PART, 16, 1024);
          const uint64_t dbh = make_desc(b_hi, 16, 1024), dbl = make_desc(b_hi + HB, 16, 1024);
          const int ksteps = (g.dbg & 8) ? 0 : (kb == g.nkb - 1 ? last_ksteps : KB / KSTEP);
          for (int j = 0; j < ksteps; ++j) {
            const uint64_t adv = uint64_t(j * 2);       // one MMA = 32 bytes along K in both kinds
            if constexpr (KIND == 1) {
              umma_f16_pair(d, dal + adv, dbh + adv, idesc, (kb | j) ? 1u : 0u);
              umma_f16_pair(d, dah + adv, dbl + adv, idesc, 1u);
              umma_f16_pair(d, dah + adv, dbh + adv, idesc, 1u);
            } else {
              umma_tf32_pair(d, dal + adv, dbh + adv, idesc, (kb | j) ? 1u : 0u);
              umma_tf32_pair(d, dah + adv, dbl + adv, idesc, 1u);
              umma_tf32_pair(d, dah + adv, dbh + adv, idesc, 1u);
            }
          }
          umma_commit_pair(bar_empty + 8 * s);
        }
        umma_commit_pair(bar_tfull + 8 * acc);
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

// ---- weight gradient ------------------------------------------------------------------------------------------
struct WgradArgs {
  const float* dz; int64_t lddz; int out_dim;
  const float* x; int64_t ldx; int in_dim;
  float* dw; int64_t lddw; float* db;
  int64_t rows; int64_t rows_per_split; int n_tiles;
  // generated dz (weight gradient below a fused head): dz[r, c] = hd[r] * hw[c] * act'(dz_mem[r, c]) with dz_mem = the
  // stored activations y of the layer; dhw[c] += sum_r hd[r] * y[r, c] (the head's own weight gradient)
  const float* hd; const float* hw; int hact; float hact_param; float* dhw;
  // transposed: the kernel computes dW^T = x^T dz (the fields above then hold: dz/lddz/out_dim = x and its width,
  // x/ldx/in_dim = dz and its width): used when in_dim <= 128 < out_dim so that the accumulator is 128 x 256 instead
  // of two 128 x 80 ones (a k-block of MMAs long enough to hide the staging hand-shake); the bias gradient is then
  // summed on the B side and the tile is reduced into dw with column stride lddw.
  int transposed;
};

// TMA: both operands are landed by tensor-map copies (box 32 floats x 32 rows = one MN-major panel,
// SWIZZLE_128B_ATOM_32B = the UMMA SWIZZLE_128B_BASE32B layout) straight into the stage's hi tiles (the tensor core
// truncates them to TF32); all 16 staging warps then only derive lo = rna(x - trunc(x)) and the bias sums from shared
// memory.  Used when both operands have 16-byte aligned rows and dz is a stored operand (not the generated head one).
// PAIR: the two 128-row m-tiles of a 256 x 256 weight gradient are the two CTAs of a cluster (blockIdx.x = rank):
// tcgen05.mma.cta_group::2 (M = 256) issued by rank 0 reads both CTAs' dz tiles and each CTA's HALF of the x tile, so
// a CTA stages 128 + 128 columns per k-block instead of 128 + 256 (x is no longer staged twice); the peer's staging
// warps arrive on the leader's barrier, completions are committed to both CTAs (see tc_rows_pair_kernel).
template <int NPARTS, bool TMA, bool PAIR>
__global__ void __launch_bounds__(WG_THREADS, 1) tc_wgrad_kernel(const WgradArgs g, const __grid_constant__ CUtensorMap tmap_a,
                                                                 const __grid_constant__ CUtensorMap tmap_b) {
  static_assert(!(TMA && PAIR), "the pair variant uses the register-staged producers");
  constexpr int S = num_stages(NPARTS);
  constexpr int STAGE = stage_bytes(NPARTS);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* stg_all = reinterpret_cast<float*>(smem + S * STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * STAGE + STG_BYTES);
  // barriers: full[4] empty[4] tfull raw[4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + MAX_STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 2 * MAX_STAGES);
  const uint32_t bar_raw = smem_u32(bars + 2 * MAX_STAGES + 4);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int mt = blockIdx.x / g.n_tiles, nt = blockIdx.x % g.n_tiles;
  const int m0 = mt * TM, n0 = nt * NT;
  const int w = tile_width(pad16(g.in_dim), nt);       // accumulator width (multiple of 16)
  // 16-byte chunks per staged row of x (whole 32-wide panels); a pair CTA stages its half of the 256 columns
  const int cpr = PAIR ? 32 : (w + 31) / 32 * 8;
  const int b_part = cpr * 16 * TK;                    // bytes of one B part: panels x 4096
  const int xcol0 = PAIR ? n0 + mt * 128 : n0;         // first x column this CTA stages
  const int64_t k_beg = int64_t(blockIdx.y) * g.rows_per_split;
  const int64_t k_end = min(g.rows, k_beg + g.rows_per_split);
  const int nkb = k_beg < k_end ? int((k_end - k_beg + TK - 1) / TK) : 0;

  if (t == 0) {
    for (int s = 0; s < S; ++s) {
      // one arrival per staging warp (of the group; of both CTAs in a pair)
      mbar_init(bar_full + 8 * s, TMA ? WG_STAGE_WARPS : (PAIR ? WG_STAGE_WARPS : WG_STAGE_WARPS / 2));
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_raw + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (PAIR) cluster_sync_all();
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (nkb > 0) {
    if (TMA && warp < WG_STAGE_WARPS) {
      // ================= converters: landed raw panels -> lo tiles, bias sums =================
      const int ca = t & 31, ka = t >> 5;      // dz tile: chunk ca of k-rows ka + 16 i (i < 2)
      const int cb = t & 63, kbb = t >> 6;     // x tile : chunk cb of k-rows kbb + 8 i (i < 4)
      const bool b_active = cb < cpr;
      const bool want_db = g.db != nullptr && nt == 0 && !g.transposed;
      const bool want_db_b = g.db != nullptr && mt == 0 && g.transposed && b_active;
      float4 colsum = make_float4(0.f, 0.f, 0.f, 0.f), bsum = colsum;
      auto lo_of = [](const float4& x) {
        float4 l;
        l.x = tf32_rna(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u));
        l.y = tf32_rna(x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u));
        l.z = tf32_rna(x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u));
        l.w = tf32_rna(x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
        return l;
      };
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t s = kb % S;
        mbar_wait(bar_raw + 8 * s, (kb / S) & 1);
        const uint32_t a_hi = smem_u32(smem + s * STAGE), b_hi = a_hi + NPARTS * PART;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t off = mn_offset(ca, ka + 16 * i);
          const float4 x = lds128(a_hi + off);
          colsum.x += x.x; colsum.y += x.y; colsum.z += x.z; colsum.w += x.w;
          if (NPARTS == 2) sts128(a_hi + PART + off, lo_of(x));
        }
        if (b_active) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t off = mn_offset(cb, kbb + 8 * i);
            const float4 x = lds128(b_hi + off);
            bsum.x += x.x; bsum.y += x.y; bsum.z += x.z; bsum.w += x.w;
            if (NPARTS == 2) sts128(b_hi + b_part + off, lo_of(x));
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * s);
      }
      if (want_db) {
        const int col = m0 + ca * 4;
        if (col < g.out_dim) atomicAdd(g.db + col, colsum.x);
        if (col + 1 < g.out_dim) atomicAdd(g.db + col + 1, colsum.y);
        if (col + 2 < g.out_dim) atomicAdd(g.db + col + 2, colsum.z);
        if (col + 3 < g.out_dim) atomicAdd(g.db + col + 3, colsum.w);
      }
      if (want_db_b) {
        const int col = n0 + cb * 4;
        if (col < g.in_dim) atomicAdd(g.db + col, bsum.x);
        if (col + 1 < g.in_dim) atomicAdd(g.db + col + 1, bsum.y);
        if (col + 2 < g.in_dim) atomicAdd(g.db + col + 2, bsum.z);
        if (col + 3 < g.in_dim) atomicAdd(g.db + col + 3, bsum.w);
      }
      if (warp < EPI_WARPS) {
        float* stg = stg_all + warp * 32 * STG_LD;
        EpiArgs e{};
        e.C = g.dw; e.ldc = g.lddw; e.M = g.out_dim; e.N = g.in_dim;
        if (g.transposed) { e.ldc = 1; e.cs = g.lddw; }
        mbar_wait(bar_tfull, 0);
        tc_fence_after();
        YPrev y_none;
        epilogue_tile<EPI_ATOMIC, false>(e, tmem, w, stg, warp, lane, m0, n0, false, false, y_none, []() {});
      }
    } else if (TMA && warp == WG_STAGE_WARPS + 1) {
      // ================= loader: one tensor-map copy per 32-wide panel =================
      if (lane == 0) {
        const int b_panels = cpr / 8;
        const uint32_t bytes = uint32_t((TM / 32 + b_panels) * 4096);
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t s = kb % S;
          mbar_wait(bar_empty + 8 * s, ((kb / S) & 1) ^ 1);
          const uint32_t a_hi = smem_u32(smem + s * STAGE), b_hi = a_hi + NPARTS * PART;
          const int k0 = int(k_beg + int64_t(kb) * TK);
          mbar_arrive_expect_tx(bar_raw + 8 * s, bytes);
          for (int p = 0; p < TM / 32; ++p) tma_load_2d(a_hi + p * 4096, &tmap_a, m0 + 32 * p, k0, bar_raw + 8 * s);
          for (int p = 0; p < b_panels; ++p) tma_load_2d(b_hi + p * 4096, &tmap_b, n0 + 32 * p, k0, bar_raw + 8 * s);
        }
      }
    } else if (!TMA && warp < WG_STAGE_WARPS) {
      // ================= producers: dz and x rows -> hi/lo -> MN-major swizzled shared memory =================
      // Two groups of eight warps alternate over the k-blocks, each owning its stages (see tc_rows_kernel): the loads
      // of a group's next k-block are in flight while the other group stores and the MMAs of its previous block run.
      constexpr int GROUPS = 2;
      static_assert(S % GROUPS == 0, "a staging group must always meet the same stages");
      const int grp = t >> 8, q = t & 255;
      const int ca = q & 31, ka = q >> 5;      // dz tile: 32 chunks per k-row, k-rows ka + 8 i (i < 4)
      // x tile: up to 64 chunks per k-row, k-rows kbb + 4 i (i < 8); pair: 32 chunks, k-rows kbb + 8 i (i < 4)
      const int cb = PAIR ? (q & 31) : (q & 63), kbb = PAIR ? (q >> 5) : (q >> 6);
      constexpr int NB = PAIR ? 4 : 8, KSTEP_B = PAIR ? 8 : 4;
      const bool vec_a = ((g.lddz & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.dz) & 15) == 0);
      const bool vec_b = ((g.ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.x) & 15) == 0);
      const bool b_active = cb < cpr;
      const bool want_db = g.db != nullptr && nt == 0 && !g.transposed;
      const bool want_db_b = g.db != nullptr && mt == 0 && g.transposed && b_active;
      float4 colsum = make_float4(0.f, 0.f, 0.f, 0.f), hsum = colsum, bsum = colsum;
      float4 va[4], vb[8];
      float hd[4];
      const float4 hw4 = g.hd && m0 + ca * 4 < g.out_dim ? load_cols4(g.hw, m0 + ca * 4, g.out_dim) : make_float4(0.f, 0.f, 0.f, 0.f);
      auto load_block = [&](int kb) {
        const int64_t k0 = k_beg + int64_t(kb) * TK;
        if (g.hd) {
#pragma unroll
          for (int i = 0; i < 4; ++i) hd[i] = k0 + ka + 8 * i < k_end ? __ldg(g.hd + k0 + ka + 8 * i) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) va[i] = load4(g.dz, g.lddz, k0 + ka + 8 * i, k_end, m0 + ca * 4, g.out_dim, vec_a);
        if (b_active) {
#pragma unroll
          for (int i = 0; i < NB; ++i) vb[i] = load4(g.x, g.ldx, k0 + kbb + KSTEP_B * i, k_end, xcol0 + cb * 4, g.in_dim, vec_b);
        }
      };
      int kb = grp;
      if (kb < nkb) load_block(kb);
      while (kb < nkb) {
        const uint32_t s = kb % S;
        mbar_wait(bar_empty + 8 * s, ((kb / S) & 1) ^ 1);
        uint8_t* a_hi = smem + s * STAGE;
        uint8_t* b_hi = a_hi + NPARTS * PART;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (g.hd) {
            const float4 y = va[i];
            const float d = hd[i];
            hsum.x = fmaf(d, y.x, hsum.x); hsum.y = fmaf(d, y.y, hsum.y); hsum.z = fmaf(d, y.z, hsum.z); hsum.w = fmaf(d, y.w, hsum.w);
            va[i] = gen_dz4_rt(y, d, hw4, g.hact, g.hact_param);
          }
          split_store4<NPARTS>(a_hi, a_hi + PART, mn_offset(ca, ka + 8 * i), va[i]);
          colsum.x += va[i].x; colsum.y += va[i].y; colsum.z += va[i].z; colsum.w += va[i].w;
        }
        if (b_active) {
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            split_store4<NPARTS>(b_hi, b_hi + b_part, mn_offset(cb, kbb + KSTEP_B * i), vb[i]);
            bsum.x += vb[i].x; bsum.y += vb[i].y; bsum.z += vb[i].z; bsum.w += vb[i].w;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {                                  // one arrival per warp
          if (PAIR) mbar_arrive_leader(bar_full + 8 * s);
          else mbar_arrive(bar_full + 8 * s);
        }
        kb += GROUPS;
        if (kb < nkb) load_block(kb);
      }
      if (want_db) {
        // 16 warps hold partial column sums for the same 128 columns: combine through the L2
        const int col = m0 + ca * 4;
        if (col < g.out_dim) atomicAdd(g.db + col, colsum.x);
        if (col + 1 < g.out_dim) atomicAdd(g.db + col + 1, colsum.y);
        if (col + 2 < g.out_dim) atomicAdd(g.db + col + 2, colsum.z);
        if (col + 3 < g.out_dim) atomicAdd(g.db + col + 3, colsum.w);
      }
      if (want_db_b) {
        const int col = n0 + cb * 4;
        if (col < g.in_dim) atomicAdd(g.db + col, bsum.x);
        if (col + 1 < g.in_dim) atomicAdd(g.db + col + 1, bsum.y);
        if (col + 2 < g.in_dim) atomicAdd(g.db + col + 2, bsum.z);
        if (col + 3 < g.in_dim) atomicAdd(g.db + col + 3, bsum.w);
      }
      if (g.hd && g.dhw && nt == 0) {
        const int col = m0 + ca * 4;
        if (col < g.out_dim) atomicAdd(g.dhw + col, hsum.x);
        if (col + 1 < g.out_dim) atomicAdd(g.dhw + col + 1, hsum.y);
        if (col + 2 < g.out_dim) atomicAdd(g.dhw + col + 2, hsum.z);
        if (col + 3 < g.out_dim) atomicAdd(g.dhw + col + 3, hsum.w);
      }
      if (warp < EPI_WARPS) {
        // ================= epilogue: partial tile -> dW with coalesced reductions =================
        float* stg = stg_all + warp * 32 * STG_LD;
        EpiArgs e{};
        e.C = g.dw; e.ldc = g.lddw; e.M = g.out_dim; e.N = g.in_dim;
        if (g.transposed) { e.ldc = 1; e.cs = g.lddw; }
        mbar_wait(bar_tfull, 0);
        tc_fence_after();
        YPrev y_none;
        epilogue_tile<EPI_ATOMIC, false>(e, tmem, w, stg, warp, lane, m0, n0, false, false, y_none, []() {});
      }
    } else if (warp == WG_STAGE_WARPS && (!PAIR || mt == 0)) {
      // ================= MMA issuer (pair: the leader CTA) =================
      if (lane == 0) {
        uint32_t idesc = make_idesc_tf32(w, true);
        if (PAIR) idesc = (1u << 4) | (2u << 7) | (2u << 10) | (3u << 15) | (uint32_t(NT >> 3) << 17) | (uint32_t((2 * TM) >> 4) << 24);
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t s = kb % S;
          mbar_wait_spin(bar_full + 8 * s, (kb / S) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + s * STAGE);
          const uint32_t b_hi = a_hi + NPARTS * PART;
          const uint64_t dah = make_desc(a_hi, 4096, 512, 1), dal = make_desc(a_hi + PART, 4096, 512, 1);
          const uint64_t dbh = make_desc(b_hi, 4096, 512, 1), dbl = make_desc(b_hi + b_part, 4096, 512, 1);
          // rows past k_end were staged as zeros, so every k-step of the last block may be issued
#pragma unroll
          for (int j = 0; j < TK / 8; ++j) {
            const uint64_t adv = uint64_t(j * (1024 >> 4));   // 8 k-rows = one 1024-byte atom
            if (PAIR) {
              if (NPARTS == 2) {
                umma_tf32_pair(tmem, dal + adv, dbh + adv, idesc, (kb | j) ? 1u : 0u);
                umma_tf32_pair(tmem, dah + adv, dbl + adv, idesc, 1u);
                umma_tf32_pair(tmem, dah + adv, dbh + adv, idesc, 1u);
              } else {
                umma_tf32_pair(tmem, dah + adv, dbh + adv, idesc, (kb | j) ? 1u : 0u);
              }
            } else if (NPARTS == 2) {
              umma_tf32(tmem, dal + adv, dbh + adv, idesc, (kb | j) ? 1u : 0u);
              umma_tf32(tmem, dah + adv, dbl + adv, idesc, 1u);
              umma_tf32(tmem, dah + adv, dbh + adv, idesc, 1u);
            } else {
              umma_tf32(tmem, dah + adv, dbh + adv, idesc, (kb | j) ? 1u : 0u);
            }
          }
          if (PAIR) umma_commit_pair(bar_empty + 8 * s);
          else umma_commit(bar_empty + 8 * s);
        }
        if (PAIR) umma_commit_pair(bar_tfull);
        else umma_commit(bar_tfull);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 0) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

template <int NPARTS, int EPI, bool TMA_A>
static int launch_rows_impl(const RowsArgs& g, const CUtensorMap& map, cudaStream_t s, const char* what) {
  static PerDeviceFlag flags;
  bool& configured = flags();
  auto kern = tc_rows_kernel<NPARTS, EPI, TMA_A>;
  if (!configured) {
    int rc = set_smem(kern, smem_bytes(NPARTS), what);
    if (rc) return rc;
    configured = true;
  }
  const int64_t grid = g.total_tiles < kNumSMs ? g.total_tiles : kNumSMs;
  kern<<<(unsigned)grid, THREADS, smem_bytes(NPARTS), s>>>(g, map);
  return check_launch(what);
}

// Tensor map of an MN-major operand of the weight-gradient kernel: [rows, cols] fp32, row stride ld floats, box 32 floats x
// 32 rows (one panel of a k-block), SWIZZLE_128B_ATOM_32B.
static bool make_mn_map(const float* p, int64_t ld, int64_t rows, int cols, CUtensorMap* map) {
  TensorMapEncodeFn encode = tensor_map_encoder();
  if (encode == nullptr) return false;
  if ((reinterpret_cast<uintptr_t>(p) & 15) != 0 || (ld & 3) != 0 || rows >= (int64_t(1) << 31)) return false;
  const cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  const cuuint64_t strides[1] = {cuuint64_t(ld) * sizeof(float)};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Tensor map of the A operand: [M rows, K floats] fp32, row stride lda floats, box 32 floats x 128 rows, SWIZZLE_128B.
static bool make_a_map(const RowsArgs& g, CUtensorMap* map) {
  static int use_tma = -1;
  if (use_tma < 0) { const char* e = getenv("MMSB_TC_TMA"); use_tma = e ? atoi(e) : 1; }
  if (!use_tma || (g.dbg & 1)) return false;
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) != 0 || (g.lda & 3) != 0 || g.M >= (int64_t(1) << 31)) return false;
  const cuuint64_t dims[2] = {cuuint64_t(g.K), cuuint64_t(g.M)};
  const cuuint64_t strides[1] = {cuuint64_t(g.lda) * sizeof(float)};
  const cuuint32_t box[2] = {cuuint32_t(TK), cuuint32_t(TM)};
  const cuuint32_t estr[2] = {1, 1};
  // the driver entry point is resolved through the runtime (no link-time dependency on libcuda: the library must load,
  // and export its symbols, on a machine without a driver)
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  static bool resolved = false;
  if (!resolved) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<EncodeFn>(fn);
    resolved = true;
  }
  if (encode == nullptr) return false;
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(g.A), dims, strides, box,
                                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}


template <int EPI, int KIND>
static int launch_rows_pair(const RowsArgs& g, const CUtensorMap& map_a, cudaStream_t s, const char* what) {
  CUtensorMap map_b;
  memset(&map_b, 0, sizeof(map_b));
  const int64_t total_rows = int64_t(pad16(g.N)) * g.nkb * 2;      // hi + lo rows (128 B each) of every k-block
  if (!make_packed_map(g.Bp, total_rows, &map_b)) return -1000;    // caller falls back to the single-CTA kernel
  static PerDeviceFlag flags;
  bool& configured = flags();
  auto kern = tc_rows_pair_kernel<EPI, KIND>;
  const int smem = PAIR_STAGES * PAIR_STAGE + STG_BYTES + 256 + 1024;
  if (!configured) {
    int rc = set_smem(kern, smem, what);
    if (rc) return rc;
    configured = true;
  }
  const int64_t ptiles = ceil_div(g.M, 2 * TM) * g.n_tiles;
  const int64_t pairs = ptiles < kNumSMs / 2 ? ptiles : kNumSMs / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(PAIR_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, g, map_a, map_b);
  if (e != cudaSuccess) {
    set_error("%s: cluster launch failed: %s", what, cudaGetErrorString(e));
    return MMSB_E_CUDA;
  }
  return check_launch(what);
}

template <int NPARTS, int EPI>
static int launch_rows(const RowsArgs& g_in, cudaStream_t s, const char* what) {
  RowsArgs g = g_in;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const bool tma = make_a_map(g, &map);
  // ring geometry: stage = A hi [+ lo] + the widest B tile of this launch (1024-byte multiple)
  const int n_pad = pad16(g.N);
  const int w_max = n_pad < NT ? n_pad : NT;
  g.stage_bytes = (NPARTS * (PART + w_max * 128) + 1023) / 1024 * 1024;
  int stages = num_stages(NPARTS) * stage_bytes(NPARTS) / g.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  { const char* e = getenv("MMSB_TC_MAX_STAGES"); if (e && atoi(e) >= 2 && stages > atoi(e)) stages = atoi(e); }   // dev
  if (!tma && (stages & 1)) --stages;       // register-staged producers: each of the two groups owns its stages
  g.stages = stages;
  {
    // CTA pairs (cta_group::2): 3xTF32, TMA-fed operand, every accumulator 256 wide, enough rows to fill the machine
    static int use_pair = -1;
    if (use_pair < 0) { const char* e = getenv("MMSB_TC_PAIR"); use_pair = e ? atoi(e) : 1; }
    if (use_pair && NPARTS == 2 && tma && n_pad % NT == 0 && g.M >= 2 * TM * (kNumSMs / 2) && !(g.dbg & 5)) {
      const int rc = launch_rows_pair<EPI, 0>(g, map, s, what);
      if (rc != -1000) return rc;
    }
  }
  if (tma) return launch_rows_impl<NPARTS, EPI, true>(g, map, s, what);
  return launch_rows_impl<NPARTS, EPI, false>(g, map, s, what);
}

// precision 2 applies where the CTA-pair kernel does: TMA-fed A (16-byte aligned rows), accumulators 256 wide, enough
// rows to fill the machine; the caller (ops.py) routes every other layer shape to the 3xTF32 kernels.
static bool f16_rows_eligible(int64_t n, int n_dim, const float* a, int64_t lda) {
  return pad16(n_dim) % NT == 0 && n >= 2 * TM * (kNumSMs / 2) && (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (lda & 3) == 0;
}
template <int EPI>
static int launch_rows_f16(const RowsArgs& g_in, cudaStream_t s, const char* what) {
  RowsArgs g = g_in;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const bool elig = f16_rows_eligible(g.M, g.N, g.A, g.lda);
  if (!elig || !make_a_map(g, &map)) {
    set_error("%s: precision 2 (fp16 split) needs 16-byte aligned operand rows, an output width that is a multiple of 256 "
              "and at least %d rows (rows %lld, width %d, lda %lld, base %% 16 = %d, tensor map %s)", what,
              2 * TM * (kNumSMs / 2), (long long)g.M, g.N, (long long)g.lda, int(reinterpret_cast<uintptr_t>(g.A) & 15),
              elig ? "could not be encoded" : "not tried");
    return MMSB_E_INVALID_ARGUMENT;
  }
  const int rc = launch_rows_pair<EPI, 1>(g, map, s, what);
  if (rc == -1000) {
    set_error("%s: could not encode the tensor map of the packed weights", what);
    return MMSB_E_CUDA;
  }
  return rc;
}

template <int NPARTS, bool TMA, bool PAIR>
static int launch_wgrad_impl(const WgradArgs& g, const CUtensorMap& ma, const CUtensorMap& mb, dim3 grid, cudaStream_t s,
                             const char* what) {
  static PerDeviceFlag flags;
  bool& configured = flags();
  auto kern = tc_wgrad_kernel<NPARTS, TMA, PAIR>;
  if (!configured) {
    int rc = set_smem(kern, smem_bytes(NPARTS), what);
    if (rc) return rc;
    configured = true;
  }
  if (PAIR) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(WG_THREADS);
    cfg.dynamicSmemBytes = smem_bytes(NPARTS);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, g, ma, mb);
    if (e != cudaSuccess) {
      set_error("%s: cluster launch failed: %s", what, cudaGetErrorString(e));
      return MMSB_E_CUDA;
    }
    return check_launch(what);
  }
  kern<<<grid, WG_THREADS, smem_bytes(NPARTS), s>>>(g, ma, mb);
  return check_launch(what);
}

template <int NPARTS>
static int launch_wgrad(const WgradArgs& g, dim3 grid, cudaStream_t s, const char* what) {
  CUtensorMap ma, mb;
  memset(&ma, 0, sizeof(ma));
  memset(&mb, 0, sizeof(mb));
  static int use_tma = -1;
  // off by default: measured 12 % slower than the register-staged producers (twelve 4 KB panel copies per k-block)
  if (use_tma < 0) { const char* e = getenv("MMSB_TC_TMA_WGRAD"); use_tma = e ? atoi(e) : 0; }
  const bool tma = use_tma && g.hd == nullptr && make_mn_map(g.dz, g.lddz, g.rows, g.out_dim, &ma) &&
                   make_mn_map(g.x, g.ldx, g.rows, g.in_dim, &mb);
  if (tma) return launch_wgrad_impl<NPARTS, true, false>(g, ma, mb, grid, s, what);
  static int use_pair = -1;
  if (use_pair < 0) { const char* e = getenv("MMSB_TC_PAIR_WGRAD"); use_pair = e ? atoi(e) : 1; }
  // CTA pair: exactly two m-tiles and one full 256-wide n-tile (grid.x = 2 = the cluster)
  if (use_pair && !g.transposed && grid.x == 2 && g.n_tiles == 1 && pad16(g.in_dim) == NT && g.out_dim > TM)
    return launch_wgrad_impl<NPARTS, false, true>(g, ma, mb, grid, s, what);
  return launch_wgrad_impl<NPARTS, false, false>(g, ma, mb, grid, s, what);
}

}  // namespace tc
}  // namespace mmsb

using namespace mmsb;

static bool valid_precision(int p) { return p == 1 || p == 2 || p == 3; }

extern "C" int64_t mmsb_linear_packed_size(int32_t n_dim, int32_t k_dim, int32_t precision) {
  if (n_dim <= 0 || k_dim <= 0 || !valid_precision(precision)) return -1;
  if (precision == 2) return tc::packed_floats_f16(n_dim, k_dim) + 4;      // + the trailer that holds max |w|
  return tc::packed_floats(n_dim, k_dim, precision == 3 ? 2 : 1);
}

extern "C" int mmsb_linear_pack_weight(const float* w, int64_t ldw, int32_t out_dim, int32_t in_dim, int32_t transpose,
                                       int32_t precision, float* packed, mmsb_stream_t stream) {
  MMSB_REQUIRE(w && packed, "linear_pack_weight: null pointer");
  MMSB_REQUIRE(out_dim > 0 && in_dim > 0 && ldw >= in_dim, "linear_pack_weight: bad shape out=%d in=%d ldw=%lld", out_dim,
               in_dim, (long long)ldw);
  MMSB_REQUIRE(valid_precision(precision), "linear_pack_weight: precision must be 1 (TF32), 2 (fp16 split) or 3 (3xTF32), got %d", precision);
  const int n = transpose ? in_dim : out_dim, k = transpose ? out_dim : in_dim;
  if (precision == 2) {
    // max |w| into the buffer's trailer (zeroed first), then the scaled hi / lo split
    float* amax = packed + tc::packed_floats_f16(n, k);
    cudaError_t e = cudaMemsetAsync(amax, 0, 4 * sizeof(float), as_stream(stream));
    if (e != cudaSuccess) { set_error("linear_pack_weight: memset failed: %s", cudaGetErrorString(e)); return MMSB_E_CUDA; }
    tc::amax_kernel<<<64, 256, 0, as_stream(stream)>>>(w, ldw, out_dim, in_dim, amax);
    if (int rc = check_launch("linear_pack_weight(amax)")) return rc;
    const int64_t pairs = int64_t(tc::pad16(n)) * ((k + tc::TK16 - 1) / tc::TK16) * 32;
    const int blocks2 = int(ceil_div(pairs, 256) < 4 * kNumSMs ? ceil_div(pairs, 256) : 4 * kNumSMs);
    tc::pack_weight_f16_kernel<<<blocks2, 256, 0, as_stream(stream)>>>(w, ldw, n, k, transpose, amax,
                                                                        reinterpret_cast<uint32_t*>(packed));
    return check_launch("linear_pack_weight");
  }
  const int64_t total = int64_t(tc::pad16(n)) * ((k + tc::TK - 1) / tc::TK) * tc::TK;
  const int blocks = int(ceil_div(total, 256) < 4 * kNumSMs ? ceil_div(total, 256) : 4 * kNumSMs);
  tc::pack_weight_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w, ldw, n, k, transpose, precision == 3 ? 2 : 1, packed);
  return check_launch("linear_pack_weight");
}

// Epilogue variant (MMSB_TC_DIRECT): 2 = fragment layout (tcgen05.ld.16x256b: a quad of lanes owns a 32-byte sector of a
// row, no shared-memory transpose), 0 = 32x32b + transpose through shared memory (default; the two measure the same).
static int epilogue_mode() {
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("MMSB_TC_DIRECT"); mode = e ? atoi(e) : 0; }
  return mode;
}

// ---- internal launch helpers shared by the plain and the fused-head entry points ----------------------------------
static int rows_fwd(const float* x, int64_t ldx, const float* packed_w, const float* b, float* y, int64_t ldy, int64_t n,
                    int in_dim, int out_dim, int act, float act_param, int precision, const float* head_w,
                    const float* head_b, float* head_out, cudaStream_t stream, const char* what,
                    const float* x_amax = nullptr, float* y_amax = nullptr) {
  tc::RowsArgs g{};
  g.A = x; g.lda = ldx; g.M = n; g.K = in_dim; g.Bp = packed_w; g.N = out_dim;
  g.epi.C = y; g.epi.ldc = ldy; g.epi.M = n; g.epi.N = out_dim; g.epi.bias = b; g.epi.act = act; g.epi.act_param = act_param;
  g.epi.head_w = head_w; g.epi.head_b = head_b; g.epi.head_out = head_out;
  g.epi.direct = epilogue_mode();
  g.n_tiles = int(ceil_div(tc::pad16(out_dim), tc::NT)); g.nkb = int(ceil_div(in_dim, tc::TK));
  g.total_tiles = ceil_div(n, tc::TM) * g.n_tiles;
  { const char* e = getenv("MMSB_TC_DEBUG"); g.dbg = e ? atoi(e) : 0; }
  { const char* e = getenv("MMSB_TC_PREFETCH"); g.prefetch = e ? atoi(e) : 0; }
  g.epi.amax_out = y_amax;
  if (precision == 2) {
    if (x_amax == nullptr) { set_error("%s: precision 2 needs the operand's max |x| (x_amax)", what); return MMSB_E_INVALID_ARGUMENT; }
    g.nkb = int(ceil_div(in_dim, tc::TK16));
    g.epi.amax_a = x_amax;
    g.epi.amax_b = packed_w + tc::packed_floats_f16(out_dim, in_dim);
    return tc::launch_rows_f16<tc::EPI_FWD>(g, stream, what);
  }
  return precision == 3 ? tc::launch_rows<2, tc::EPI_FWD>(g, stream, what) : tc::launch_rows<1, tc::EPI_FWD>(g, stream, what);
}

static int rows_dgrad(const float* a, int64_t lda, const float* packed_wt, float* dx, int64_t lddx, const float* y_prev,
                      int64_t ld_yprev, int act_prev, float act_prev_param, int64_t n, int in_dim, int out_dim,
                      int precision, const float* r1_d, const float* r1_w, const float* hd, const float* hw, int hact,
                      float hact_param, cudaStream_t stream, const char* what, int accumulate = 0) {
  tc::RowsArgs g{};
  g.A = a; g.lda = lda; g.M = n; g.K = out_dim; g.Bp = packed_wt; g.N = in_dim;
  g.epi.C = dx; g.epi.ldc = lddx; g.epi.M = n; g.epi.N = in_dim;
  g.epi.yprev = y_prev; g.epi.ld_yprev = ld_yprev; g.epi.act_prev = act_prev; g.epi.act_prev_param = act_prev_param;
  g.epi.r1_d = r1_d; g.epi.r1_w = r1_w;
  g.epi.accumulate = accumulate;
  g.epi.direct = epilogue_mode();
  g.hd = hd; g.hw = hw; g.hact = hact; g.hact_param = hact_param;
  g.n_tiles = int(ceil_div(tc::pad16(in_dim), tc::NT)); g.nkb = int(ceil_div(out_dim, tc::TK));
  g.total_tiles = ceil_div(n, tc::TM) * g.n_tiles;
  { const char* e = getenv("MMSB_TC_DEBUG"); g.dbg = e ? atoi(e) : 0; }
  { const char* e = getenv("MMSB_TC_PREFETCH"); g.prefetch = e ? atoi(e) : 0; }
  return precision == 3 ? tc::launch_rows<2, tc::EPI_DGRAD>(g, stream, what) : tc::launch_rows<1, tc::EPI_DGRAD>(g, stream, what);
}

static int rows_wgrad(const float* dz, int64_t lddz, const float* x, int64_t ldx, float* dw, float* db, int64_t n, int in_dim,
                      int out_dim, int precision, const float* hd, const float* hw, int hact, float hact_param, float* dhw,
                      cudaStream_t stream, const char* what) {
  tc::WgradArgs g{};
  g.dz = dz; g.lddz = lddz; g.out_dim = out_dim; g.x = x; g.ldx = ldx; g.in_dim = in_dim;
  g.dw = dw; g.lddw = in_dim; g.db = db; g.rows = n;
  g.hd = hd; g.hw = hw; g.hact = hact; g.hact_param = hact_param; g.dhw = dhw;
  if (hd == nullptr && in_dim <= tc::TM && out_dim > tc::TM) {
    g.transposed = 1;
    g.dz = x; g.lddz = ldx; g.out_dim = in_dim; g.x = dz; g.ldx = lddz; g.in_dim = out_dim;
  }
  const int m_tiles = int(ceil_div(g.out_dim, tc::TM));
  g.n_tiles = int(ceil_div(tc::pad16(g.in_dim), tc::NT));
  const int tiles = m_tiles * g.n_tiles;
  int64_t splits = kNumSMs / tiles;
  if (splits < 1) splits = 1;
  const int64_t max_splits = ceil_div(n, 4 * tc::TK);
  if (splits > max_splits) splits = max_splits;
  g.rows_per_split = ceil_div(ceil_div(n, splits), tc::TK) * tc::TK;
  dim3 grid((unsigned)tiles, (unsigned)ceil_div(n, g.rows_per_split));
  return precision == 3 ? tc::launch_wgrad<2>(g, grid, stream, what) : tc::launch_wgrad<1>(g, grid, stream, what);
}

extern "C" int mmsb_linear_fwd_tc(const float* x, int64_t ldx, const float* packed_w, const float* b, float* y, int64_t ldy,
                                  int64_t n, int32_t in_dim, int32_t out_dim, int32_t act, float act_param, int32_t precision,
                                  const float* x_amax, float* y_amax, mmsb_stream_t stream) {
  MMSB_REQUIRE(x && packed_w && y, "linear_fwd_tc: null pointer");
  MMSB_REQUIRE(n >= 0 && in_dim > 0 && out_dim > 0 && ldx >= in_dim && ldy >= out_dim, "linear_fwd_tc: bad shape");
  MMSB_REQUIRE(act >= MMSB_ACT_NONE && act <= MMSB_ACT_SIGMOID, "linear_fwd_tc: unknown activation %d", act);
  MMSB_REQUIRE(valid_precision(precision), "linear_fwd_tc: precision must be 1, 2 or 3, got %d", precision);
  if (n == 0) return MMSB_OK;
  return rows_fwd(x, ldx, packed_w, b, y, ldy, n, in_dim, out_dim, act, act_param, precision, nullptr, nullptr, nullptr,
                  as_stream(stream), "linear_fwd_tc", x_amax, y_amax);
}

extern "C" int mmsb_linear_fwd_head_tc(const float* x, int64_t ldx, const float* packed_w, const float* b, float* y,
                                       int64_t ldy, int64_t n, int32_t in_dim, int32_t out_dim, int32_t act, float act_param,
                                       int32_t precision, const float* head_w, const float* head_b, float* head_out,
                                       const float* x_amax, float* y_amax, mmsb_stream_t stream) {
  MMSB_REQUIRE(x && packed_w && head_w && head_out, "linear_fwd_head_tc: null pointer");
  MMSB_REQUIRE(n >= 0 && in_dim > 0 && out_dim > 0 && out_dim <= tc::NT && ldx >= in_dim && (!y || ldy >= out_dim),
               "linear_fwd_head_tc: bad shape (the fused head needs out_dim <= 256)");
  MMSB_REQUIRE(act >= MMSB_ACT_NONE && act <= MMSB_ACT_SIGMOID, "linear_fwd_head_tc: unknown activation %d", act);
  MMSB_REQUIRE(valid_precision(precision), "linear_fwd_head_tc: precision must be 1, 2 or 3, got %d", precision);
  if (n == 0) return MMSB_OK;
  return rows_fwd(x, ldx, packed_w, b, y, ldy, n, in_dim, out_dim, act, act_param, precision, head_w, head_b, head_out,
                  as_stream(stream), "linear_fwd_head_tc", x_amax, y_amax);
}

extern "C" int mmsb_linear_bwd_data_tc(const float* dz, int64_t lddz, const float* packed_wt, float* dx, int64_t lddx,
                                       const float* y_prev, int64_t ld_yprev, int32_t act_prev, float act_prev_param,
                                       int64_t n, int32_t in_dim, int32_t out_dim, int32_t precision, int32_t accumulate,
                                       mmsb_stream_t stream) {
  MMSB_REQUIRE(dz && packed_wt && dx, "linear_bwd_data_tc: null pointer");
  MMSB_REQUIRE(n >= 0 && in_dim > 0 && out_dim > 0 && lddz >= out_dim && lddx >= in_dim, "linear_bwd_data_tc: bad shape");
  MMSB_REQUIRE(valid_precision(precision), "linear_bwd_data_tc: precision must be 1 or 3, got %d", precision);
  if (n == 0) return MMSB_OK;
  return rows_dgrad(dz, lddz, packed_wt, dx, lddx, y_prev, ld_yprev, act_prev, act_prev_param, n, in_dim, out_dim, precision,
                    nullptr, nullptr, nullptr, nullptr, 0, 0.f, as_stream(stream), "linear_bwd_data_tc", accumulate != 0);
}

extern "C" int mmsb_linear_bwd_data_rank1_tc(const float* dz, int64_t lddz, const float* packed_wt, float* dx, int64_t lddx,
                                             const float* y_prev, int64_t ld_yprev, int32_t act_prev, float act_prev_param,
                                             int64_t n, int32_t in_dim, int32_t out_dim, int32_t precision,
                                             const float* head_d, const float* head_w, mmsb_stream_t stream) {
  MMSB_REQUIRE(dz && packed_wt && dx && head_d && head_w, "linear_bwd_data_rank1_tc: null pointer");
  MMSB_REQUIRE(n >= 0 && in_dim > 0 && out_dim > 0 && lddz >= out_dim && lddx >= in_dim, "linear_bwd_data_rank1_tc: bad shape");
  MMSB_REQUIRE(valid_precision(precision), "linear_bwd_data_rank1_tc: precision must be 1 or 3, got %d", precision);
  if (n == 0) return MMSB_OK;
  return rows_dgrad(dz, lddz, packed_wt, dx, lddx, y_prev, ld_yprev, act_prev, act_prev_param, n, in_dim, out_dim, precision,
                    head_d, head_w, nullptr, nullptr, 0, 0.f, as_stream(stream), "linear_bwd_data_rank1_tc");
}

extern "C" int mmsb_linear_bwd_data_head_tc(const float* y, int64_t ldy, int32_t act, float act_param, const float* head_d,
                                            const float* head_w, const float* packed_wt, float* dx, int64_t lddx,
                                            const float* y_prev, int64_t ld_yprev, int32_t act_prev, float act_prev_param,
                                            int64_t n, int32_t in_dim, int32_t out_dim, int32_t precision,
                                            mmsb_stream_t stream) {
  MMSB_REQUIRE(y && head_d && head_w && packed_wt && dx, "linear_bwd_data_head_tc: null pointer");
  MMSB_REQUIRE(n >= 0 && in_dim > 0 && out_dim > 0 && ldy >= out_dim && lddx >= in_dim, "linear_bwd_data_head_tc: bad shape");
  MMSB_REQUIRE(valid_precision(precision), "linear_bwd_data_head_tc: precision must be 1 or 3, got %d", precision);
  if (n == 0) return MMSB_OK;
  return rows_dgrad(y, ldy, packed_wt, dx, lddx, y_prev, ld_yprev, act_prev, act_prev_param, n, in_dim, out_dim, precision,
                    nullptr, nullptr, head_d, head_w, act, act_param, as_stream(stream), "linear_bwd_data_head_tc");
}

extern "C" int mmsb_linear_bwd_weight_tc(const float* dz, int64_t lddz, const float* x, int64_t ldx, float* dw, float* db,
                                         int64_t n, int32_t in_dim, int32_t out_dim, int32_t precision,
                                         mmsb_stream_t stream) {
  MMSB_REQUIRE(dz && x && dw, "linear_bwd_weight_tc: null pointer");
  MMSB_REQUIRE(n >= 0 && in_dim > 0 && out_dim > 0 && lddz >= out_dim && ldx >= in_dim, "linear_bwd_weight_tc: bad shape");
  MMSB_REQUIRE(valid_precision(precision), "linear_bwd_weight_tc: precision must be 1 or 3, got %d", precision);
  if (n == 0) return MMSB_OK;
  return rows_wgrad(dz, lddz, x, ldx, dw, db, n, in_dim, out_dim, precision, nullptr, nullptr, 0, 0.f, nullptr,
                    as_stream(stream), "linear_bwd_weight_tc");
}

extern "C" int mmsb_linear_bwd_weight_head_tc(const float* y, int64_t ldy, int32_t act, float act_param, const float* head_d,
                                              const float* head_w, const float* x, int64_t ldx, float* dw, float* db,
                                              float* dhead_w, int64_t n, int32_t in_dim, int32_t out_dim, int32_t precision,
                                              mmsb_stream_t stream) {
  MMSB_REQUIRE(y && head_d && head_w && x && dw, "linear_bwd_weight_head_tc: null pointer");
  MMSB_REQUIRE(n >= 0 && in_dim > 0 && out_dim > 0 && ldy >= out_dim && ldx >= in_dim, "linear_bwd_weight_head_tc: bad shape");
  MMSB_REQUIRE(valid_precision(precision), "linear_bwd_weight_head_tc: precision must be 1 or 3, got %d", precision);
  if (n == 0) return MMSB_OK;
  return rows_wgrad(y, ldy, x, ldx, dw, db, n, in_dim, out_dim, precision, head_d, head_w, act, act_param, dhead_w,
                    as_stream(stream), "linear_bwd_weight_head_tc");
}

extern "C" int mmsb_amax(const float* x, int64_t ldx, int64_t n, int32_t cols, float* amax, mmsb_stream_t stream) {
  MMSB_REQUIRE(amax && n >= 0 && cols >= 1 && ldx >= cols, "amax: bad arguments");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(x != nullptr, "amax: NULL pointer");
  if (ldx == cols && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && ((n * cols) & 3) == 0) {
    tc::amax4_kernel<<<4 * kNumSMs, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), n * cols / 4, amax);
  } else {
    tc::amax_kernel<<<4 * kNumSMs, 256, 0, as_stream(stream)>>>(x, ldx, n, cols, amax);
  }
  return check_launch("amax");
}
