// A11 — dense layers of the field MLPs, fp32 SIMT path (parity mode: plain IEEE fp32 accumulation).
// ref: src/field_components/mlp.py:152-171.  The tcgen05 path (mmsb_linear_*_tc) lives in mlp_tc.cu.
//
// One register-tiled SGEMM (128x128x16 tiles, 8x8 per thread, double-buffered shared memory, LDS.128
// fragments) serves the three products of a layer:
//   forward   y  = act(x W^T + b)      A = x  [n,in]  (K contiguous)   B = W [out,in] (K contiguous)
//   dgrad     dx = dz W (* act'(y_prev)) A = dz [n,out] (K contiguous)   B = W [out,in] (N contiguous)
//   wgrad     dW += dz^T x             A = dz [n,out] (M contiguous)   B = x [n,in]  (N contiguous), split-K
// Layers with out_dim <= 16 (sdf-only last layer, modality heads, density head) use skinny
// row-streaming kernels instead of padding a 128-wide tile.
#include "common.cuh"

namespace mmsb {

__device__ __forceinline__ float act_fwd(float z, int act, float p) {
  switch (act) {
    case MMSB_ACT_RELU: return fmaxf(z, 0.f);
    case MMSB_ACT_SOFTPLUS: {
      const float zb = z * p;
      return zb > 20.f ? z : log1pf(expf(zb)) / p;
    }
    case MMSB_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    default: return z;
  }
}
// derivative expressed through the layer OUTPUT y
__device__ __forceinline__ float act_bwd_from_y(float y, int act, float p) {
  switch (act) {
    case MMSB_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case MMSB_ACT_SOFTPLUS: {
      const float yb = y * p;
      return yb > 20.f ? 1.f : -expm1f(-yb);
    }
    case MMSB_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

struct GemmArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int64_t M; int64_t N; int64_t K;
  const float* bias; int act; float act_param;
  const float* yprev; int64_t ld_yprev; int act_prev; float act_prev_param;
  int64_t k_per_split;
};

enum { EPI_FWD = 0, EPI_DGRAD = 1, EPI_ATOMIC = 2 };

// Loads an (extent x BK) operand tile into registers.  KC: element (r, k) at p[r*ld + k];
// otherwise element (r, k) at p[k*ld + r].
template <bool KC>
__device__ __forceinline__ void load_tile(const float* __restrict__ p, int64_t ld, int64_t r0, int64_t rmax,
                                          int64_t k0, int64_t kmax, float (&reg)[8]) {
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = t + 256 * i;
    int r, k;
    if (KC) { k = idx & (BK - 1); r = idx >> 4; } else { r = idx & (BM - 1); k = idx >> 7; }
    const int64_t gr = r0 + r, gk = k0 + k;
    float v = 0.f;
    if (gr < rmax && gk < kmax) v = KC ? __ldg(p + gr * ld + gk) : __ldg(p + gk * ld + gr);
    reg[i] = v;
  }
}
template <bool KC>
__device__ __forceinline__ void store_tile(float (*s)[BM + PAD], const float (&reg)[8]) {
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = t + 256 * i;
    int r, k;
    if (KC) { k = idx & (BK - 1); r = idx >> 4; } else { r = idx & (BM - 1); k = idx >> 7; }
    s[k][r] = reg[i];
  }
}

template <bool A_KC, bool B_KC, int EPI>
__global__ void __launch_bounds__(256, 2) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int64_t m0 = int64_t(blockIdx.x) * BM, n0 = int64_t(blockIdx.y) * BN;
  const int64_t kbeg = int64_t(blockIdx.z) * g.k_per_split;
  const int64_t kend = min(g.K, kbeg + g.k_per_split);
  if (kbeg >= kend) return;
  const int nk = int((kend - kbeg + BK - 1) / BK);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  load_tile<A_KC>(g.A, g.lda, m0, g.M, kbeg, kend, ra);
  load_tile<B_KC>(g.B, g.ldb, n0, g.N, kbeg, kend, rb);
  store_tile<A_KC>(As[0], ra);
  store_tile<B_KC>(Bs[0], rb);
  __syncthreads();
  int cur = 0;
  for (int kt = 0; kt < nk; ++kt) {
    const bool more = kt + 1 < nk;
    if (more) {
      const int64_t k0 = kbeg + int64_t(kt + 1) * BK;
      load_tile<A_KC>(g.A, g.lda, m0, g.M, k0, kend, ra);
      load_tile<B_KC>(g.B, g.ldb, n0, g.N, k0, kend, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tile<A_KC>(As[cur ^ 1], ra);
      store_tile<B_KC>(Bs[cur ^ 1], rb);
    }
    __syncthreads();
    cur ^= 1;
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= g.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int64_t gn0 = n0 + (jh ? 64 + tx * 4 : tx * 4);
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t gn = gn0 + j;
        if (gn >= g.N) continue;
        if (EPI == EPI_FWD) {
          float z = v[j];
          if (g.bias) z += __ldg(g.bias + gn);
          g.C[gm * g.ldc + gn] = act_fwd(z, g.act, g.act_param);
        } else if (EPI == EPI_DGRAD) {
          float d = v[j];
          if (g.yprev) d *= act_bwd_from_y(__ldg(g.yprev + gm * g.ld_yprev + gn), g.act_prev, g.act_prev_param);
          g.C[gm * g.ldc + gn] = d;
        } else {
          atomicAdd(g.C + gm * g.ldc + gn, v[j]);
        }
      }
    }
  }
}

// ---- skinny kernels (out_dim <= 16) ---------------------------------------------------------
constexpr int SK_MAX_OUT = 16;

// forward: one warp per row, lanes stride over in_dim; weights through the read-only path (<= 16 KiB).
__global__ void __launch_bounds__(256) skinny_fwd_kernel(const float* __restrict__ x, int64_t ldx,
                                                         const float* __restrict__ w, const float* __restrict__ b,
                                                         float* __restrict__ y, int64_t ldy, int64_t n, int in_dim,
                                                         int out_dim, int act, float act_param) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  float acc[SK_MAX_OUT];
#pragma unroll
  for (int j = 0; j < SK_MAX_OUT; ++j) acc[j] = 0.f;
  const float* xr = x + row * ldx;
  for (int k = lane; k < in_dim; k += 32) {
    const float xv = __ldg(xr + k);
#pragma unroll
    for (int j = 0; j < SK_MAX_OUT; ++j)
      if (j < out_dim) acc[j] = fmaf(xv, __ldg(w + j * in_dim + k), acc[j]);
  }
#pragma unroll
  for (int j = 0; j < SK_MAX_OUT; ++j)
    if (j < out_dim) acc[j] = warp_sum(acc[j]);
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < SK_MAX_OUT; ++j)
      if (j < out_dim) {
        float z = acc[j];
        if (b) z += __ldg(b + j);
        y[row * ldy + j] = act_fwd(z, act, act_param);
      }
  }
}

// dgrad: one thread per (row, k)
__global__ void __launch_bounds__(256) skinny_dgrad_kernel(const float* __restrict__ dz, int64_t lddz,
                                                           const float* __restrict__ w, float* __restrict__ dx,
                                                           int64_t lddx, const float* __restrict__ yprev,
                                                           int64_t ld_yprev, int act_prev, float act_prev_param,
                                                           int64_t n, int in_dim, int out_dim) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * in_dim) return;
  const int64_t row = t / in_dim;
  const int k = int(t - row * in_dim);
  float acc = 0.f;
  for (int j = 0; j < out_dim; ++j) acc = fmaf(__ldg(dz + row * lddz + j), __ldg(w + j * in_dim + k), acc);
  if (yprev) acc *= act_bwd_from_y(__ldg(yprev + row * ld_yprev + k), act_prev, act_prev_param);
  dx[row * lddx + k] = acc;
}

// wgrad: block = 256 threads over k columns (looped), rows strided by gridDim.x chunks; one atomic per
// (block, j, k).
__global__ void __launch_bounds__(256) skinny_wgrad_kernel(const float* __restrict__ dz, int64_t lddz,
                                                           const float* __restrict__ x, int64_t ldx,
                                                           float* __restrict__ dw, int64_t n, int in_dim, int out_dim,
                                                           int64_t rows_per_block) {
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(n, r0 + rows_per_block);
  for (int k = threadIdx.x; k < in_dim; k += blockDim.x) {
    float acc[SK_MAX_OUT];
#pragma unroll
    for (int j = 0; j < SK_MAX_OUT; ++j) acc[j] = 0.f;
    for (int64_t r = r0; r < r1; ++r) {
      const float xv = __ldg(x + r * ldx + k);
#pragma unroll
      for (int j = 0; j < SK_MAX_OUT; ++j)
        if (j < out_dim) acc[j] = fmaf(__ldg(dz + r * lddz + j), xv, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < SK_MAX_OUT; ++j)
      if (j < out_dim) atomicAdd(dw + j * in_dim + k, acc[j]);
  }
}

// db[j] += sum_rows dz[row, j]: block (32 cols x 8 row-lanes), rows chunked over gridDim.y.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dz, int64_t lddz, float* __restrict__ db,
                                                     int64_t n, int dim, int64_t rows_per_block) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  const int64_t r0 = int64_t(blockIdx.y) * rows_per_block;
  const int64_t r1 = min(n, r0 + rows_per_block);
  float acc = 0.f;
  if (c < dim)
    for (int64_t r = r0 + rl; r < r1; r += 8) acc += __ldg(dz + r * lddz + c);
  red[rl][threadIdx.x & 31] = acc;
  __syncthreads();
  if (rl == 0 && c < dim) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x & 31];
    atomicAdd(db + c, s);
  }
}

__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dy, int64_t lddy,
                                                      const float* __restrict__ y, int64_t ldy, float* __restrict__ dz,
                                                      int64_t lddz, int64_t n, int dim, int act, float p) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * dim) return;
  const int64_t row = t / dim;
  const int c = int(t - row * dim);
  dz[row * lddz + c] = __ldg(dy + row * lddy + c) * act_bwd_from_y(__ldg(y + row * ldy + c), act, p);
}

// 16 bytes per thread when all three operands have 16-byte aligned rows and dim % 4 == 0 (the padded rows of the
// pipeline): the kernel is a pure HBM stream (12 B per element)
__global__ void __launch_bounds__(256) act_bwd4_kernel(const float* __restrict__ dy, int64_t lddy,
                                                       const float* __restrict__ y, int64_t ldy, float* __restrict__ dz,
                                                       int64_t lddz, int64_t n, int dim4, int act, float p) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * dim4) return;
  const int64_t row = t / dim4;
  const int c = int(t - row * dim4) * 4;
  const float4 g = __ldg(reinterpret_cast<const float4*>(dy + row * lddy + c));
  const float4 v = __ldg(reinterpret_cast<const float4*>(y + row * ldy + c));
  float4 o;
  o.x = g.x * act_bwd_from_y(v.x, act, p); o.y = g.y * act_bwd_from_y(v.y, act, p);
  o.z = g.z * act_bwd_from_y(v.z, act, p); o.w = g.w * act_bwd_from_y(v.w, act, p);
  *reinterpret_cast<float4*>(dz + row * lddz + c) = o;
}

// dst[r, 0:w] = src[r, 0:w] for row-strided operands (the "copy" pieces of an assembled MLP input row)
__global__ void __launch_bounds__(256) copy_rows4_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst,
                                                         int64_t ldd, int64_t n, int w4) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * w4) return;
  const int64_t row = t / w4;
  const int c = int(t - row * w4) * 4;
  *reinterpret_cast<float4*>(dst + row * ldd + c) = __ldg(reinterpret_cast<const float4*>(src + row * lds + c));
}
__global__ void __launch_bounds__(256) copy_rows_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst,
                                                        int64_t ldd, int64_t n, int w) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n * w) return;
  const int64_t row = t / w;
  const int c = int(t - row * w);
  dst[row * ldd + c] = __ldg(src + row * lds + c);
}

static __host__ bool aligned16(const void* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld & 3) == 0; }

}  // namespace mmsb

using namespace mmsb;

static bool valid_act(int a) { return a >= MMSB_ACT_NONE && a <= MMSB_ACT_SIGMOID; }

extern "C" int mmsb_linear_fwd(const float* x, int64_t ldx, const float* w, const float* b, float* y, int64_t ldy,
                               int64_t n, int32_t in_dim, int32_t out_dim, int32_t act, float act_param,
                               mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && in_dim >= 1 && out_dim >= 1 && ldx >= in_dim && ldy >= out_dim,
               "linear_fwd: bad sizes n=%lld in=%d out=%d ldx=%lld ldy=%lld", (long long)n, in_dim, out_dim,
               (long long)ldx, (long long)ldy);
  MMSB_REQUIRE(valid_act(act), "linear_fwd: unknown activation %d", act);
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(x && w && y, "linear_fwd: NULL pointer");
  cudaStream_t s = as_stream(stream);
  if (out_dim <= SK_MAX_OUT) {
    skinny_fwd_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, s>>>(x, ldx, w, b, y, ldy, n, in_dim, out_dim, act,
                                                                      act_param);
    return check_launch("linear_fwd(skinny)");
  }
  GemmArgs g{};
  g.A = x; g.lda = ldx; g.B = w; g.ldb = in_dim; g.C = y; g.ldc = ldy;
  g.M = n; g.N = out_dim; g.K = in_dim; g.bias = b; g.act = act; g.act_param = act_param;
  g.k_per_split = in_dim;
  dim3 grid((unsigned)ceil_div(n, BM), (unsigned)ceil_div(out_dim, BN), 1);
  sgemm_kernel<true, true, EPI_FWD><<<grid, 256, 0, s>>>(g);
  return check_launch("linear_fwd");
}

extern "C" int mmsb_act_bwd(const float* dy, int64_t lddy, const float* y, int64_t ldy, float* dz, int64_t lddz,
                            int64_t n, int32_t dim, int32_t act, float act_param, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && dim >= 1 && lddy >= dim && ldy >= dim && lddz >= dim, "act_bwd: bad sizes");
  MMSB_REQUIRE(valid_act(act), "act_bwd: unknown activation %d", act);
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(dy && y && dz, "act_bwd: NULL pointer");
  if ((dim & 3) == 0 && aligned16(dy, lddy) && aligned16(y, ldy) && aligned16(dz, lddz)) {
    act_bwd4_kernel<<<(unsigned)ceil_div(n * (dim / 4), 256), 256, 0, as_stream(stream)>>>(dy, lddy, y, ldy, dz, lddz, n,
                                                                                          dim / 4, act, act_param);
    return check_launch("act_bwd");
  }
  act_bwd_kernel<<<(unsigned)ceil_div(n * dim, 256), 256, 0, as_stream(stream)>>>(dy, lddy, y, ldy, dz, lddz, n, dim,
                                                                                  act, act_param);
  return check_launch("act_bwd");
}

extern "C" int mmsb_copy_rows(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t n, int32_t width,
                              mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && width >= 1 && lds >= width && ldd >= width, "copy_rows: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(src && dst, "copy_rows: NULL pointer");
  if ((width & 3) == 0 && aligned16(src, lds) && aligned16(dst, ldd))
    copy_rows4_kernel<<<(unsigned)ceil_div(n * (width / 4), 256), 256, 0, as_stream(stream)>>>(src, lds, dst, ldd, n, width / 4);
  else
    copy_rows_kernel<<<(unsigned)ceil_div(n * width, 256), 256, 0, as_stream(stream)>>>(src, lds, dst, ldd, n, width);
  return check_launch("copy_rows");
}

extern "C" int mmsb_linear_bwd_data(const float* dz, int64_t lddz, const float* w, float* dx, int64_t lddx,
                                    const float* y_prev, int64_t ld_yprev, int32_t act_prev, float act_prev_param,
                                    int64_t n, int32_t in_dim, int32_t out_dim, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && in_dim >= 1 && out_dim >= 1 && lddz >= out_dim && lddx >= in_dim &&
                   (!y_prev || ld_yprev >= in_dim),
               "linear_bwd_data: bad sizes");
  MMSB_REQUIRE(valid_act(act_prev), "linear_bwd_data: unknown activation %d", act_prev);
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(dz && w && dx, "linear_bwd_data: NULL pointer");
  cudaStream_t s = as_stream(stream);
  if (out_dim <= SK_MAX_OUT) {
    skinny_dgrad_kernel<<<(unsigned)ceil_div(n * in_dim, 256), 256, 0, s>>>(dz, lddz, w, dx, lddx, y_prev, ld_yprev,
                                                                            act_prev, act_prev_param, n, in_dim, out_dim);
    return check_launch("linear_bwd_data(skinny)");
  }
  GemmArgs g{};
  g.A = dz; g.lda = lddz; g.B = w; g.ldb = in_dim; g.C = dx; g.ldc = lddx;
  g.M = n; g.N = in_dim; g.K = out_dim;
  g.yprev = y_prev; g.ld_yprev = ld_yprev; g.act_prev = act_prev; g.act_prev_param = act_prev_param;
  g.k_per_split = out_dim;
  dim3 grid((unsigned)ceil_div(n, BM), (unsigned)ceil_div(in_dim, BN), 1);
  sgemm_kernel<true, false, EPI_DGRAD><<<grid, 256, 0, s>>>(g);
  return check_launch("linear_bwd_data");
}

extern "C" int mmsb_linear_bwd_weight(const float* dz, int64_t lddz, const float* x, int64_t ldx, float* dw, float* db,
                                      int64_t n, int32_t in_dim, int32_t out_dim, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && in_dim >= 1 && out_dim >= 1 && lddz >= out_dim && ldx >= in_dim, "linear_bwd_weight: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(dz && x && (dw || db), "linear_bwd_weight: NULL pointer");
  cudaStream_t s = as_stream(stream);
  if (db) {
    const int64_t rows_per_block = 2048;
    dim3 grid((unsigned)ceil_div(out_dim, 32), (unsigned)ceil_div(n, rows_per_block));
    colsum_kernel<<<grid, 256, 0, s>>>(dz, lddz, db, n, out_dim, rows_per_block);
    if (int e = check_launch("linear_bwd_weight(bias)")) return e;
  }
  if (!dw) return MMSB_OK;
  if (out_dim <= SK_MAX_OUT) {
    int64_t blocks = ceil_div(n, 512);
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    const int64_t rows_per_block = ceil_div(n, blocks);
    skinny_wgrad_kernel<<<(unsigned)ceil_div(n, rows_per_block), 256, 0, s>>>(dz, lddz, x, ldx, dw, n, in_dim, out_dim,
                                                                             rows_per_block);
    return check_launch("linear_bwd_weight(skinny)");
  }
  GemmArgs g{};
  g.A = dz; g.lda = lddz; g.B = x; g.ldb = ldx; g.C = dw; g.ldc = in_dim;
  g.M = out_dim; g.N = in_dim; g.K = n;
  const int64_t tiles = ceil_div(out_dim, BM) * ceil_div(in_dim, BN);
  int64_t splits = ceil_div(int64_t(4) * kNumSMs, tiles);
  const int64_t max_splits = ceil_div(n, 8 * BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t kps = ceil_div(n, splits);
  kps = ceil_div(kps, BK) * BK;
  g.k_per_split = kps;
  dim3 grid((unsigned)ceil_div(out_dim, BM), (unsigned)ceil_div(in_dim, BN), (unsigned)ceil_div(n, kps));
  sgemm_kernel<false, false, EPI_ATOMIC><<<grid, 256, 0, s>>>(g);
  return check_launch("linear_bwd_weight");
}
