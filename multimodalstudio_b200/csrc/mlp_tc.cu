// A11 — dense layers on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// One kernel, three products (forward, dgrad, wgrad) of a layer, fp32 in / fp32 out, fp32-accurate:
// every fp32 operand x is split in shared memory into hi = tf32(x) and lo = x - hi and the product is
// accumulated as hi*hi + lo*hi + hi*lo (3 x kind::tf32, fp32 accumulation in TMEM), which keeps the
// 1e-5 parity band of the fp32 reference while running on the tensor pipe.
//
// Tile: D[128 x BN] (BN <= 256) per CTA, K in blocks of 32 fp32 (one 128-byte swizzle row), two shared-memory
// stages.  All 8 warps stage operands (global -> registers -> hi/lo split -> K-major SWIZZLE_128B shared
// memory; transposing on the fly where the global layout is MN-major), one elected thread issues the MMAs,
// tcgen05.commit signals stage reuse / accumulator completion through mbarriers, all 8 warps drain TMEM with
// tcgen05.ld (32 lanes x 32 columns per instruction) and apply the fused epilogue (bias + activation,
// activation derivative of the previous layer, or split-K atomics for the weight gradient).
#include "common.cuh"

namespace mmsb {

// activation helpers (same definitions as mlp_simt.cu)
__device__ __forceinline__ float tc_act_fwd(float z, int act, float p) {
  switch (act) {
    case MMSB_ACT_RELU: return fmaxf(z, 0.f);
    case MMSB_ACT_SOFTPLUS: { const float zb = z * p; return zb > 20.f ? z : log1pf(expf(zb)) / p; }
    case MMSB_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    default: return z;
  }
}
__device__ __forceinline__ float tc_act_bwd_from_y(float y, int act, float p) {
  switch (act) {
    case MMSB_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case MMSB_ACT_SOFTPLUS: { const float yb = y * p; return yb > 20.f ? 1.f : -expm1f(-yb); }
    case MMSB_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}

struct TcArgs {
  const float* A; int64_t lda;   // A_KC: A(m,k) = A[m*lda + k]; else A(m,k) = A[k*lda + m]
  const float* B; int64_t ldb;   // B_KC: B(n,k) = B[n*ldb + k]; else B(n,k) = B[k*ldb + n]
  float* C; int64_t ldc;
  int64_t M; int64_t N; int64_t K;
  const float* bias; int act; float act_param;
  const float* yprev; int64_t ld_yprev; int act_prev; float act_prev_param;
  int64_t k_per_split;
  int bn;                        // N tile (multiple of 16, <= 256)
};

constexpr int TM = 128, TK = 32, TC_THREADS = 256, TC_STAGES = 2;
constexpr int A_TILE_BYTES = TM * 128;           // 128 rows x 128 B (one operand, hi or lo)
constexpr int B_TILE_BYTES = 256 * 128;          // up to 256 rows

__host__ __device__ constexpr int tc_stage_bytes() { return 2 * A_TILE_BYTES + 2 * B_TILE_BYTES; }
constexpr int TC_SMEM_BYTES = TC_STAGES * tc_stage_bytes() + 1024 /*align*/ + 64 /*barriers*/;

enum { TC_EPI_FWD = 0, TC_EPI_DGRAD = 1, TC_EPI_ATOMIC = 2 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100: version 1): 8-row atoms of 1024 B.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4);        // start address
  d |= uint64_t(1) << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= uint64_t(1024 >> 4) << 32;               // stride byte offset between 8-row groups
  d |= uint64_t(1) << 46;                       // descriptor version (Blackwell)
  d |= uint64_t(2) << 61;                       // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_tf32(int n) {
  // c_format F32 (1 << 4), a/b format TF32 (2 << 7, 2 << 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(TM >> 4) << 24);
}

// byte offset of element (row r, k) inside a [rows x 32 fp32] K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128(int r, int k) {
  return uint32_t(r) * 128u + ((uint32_t(k >> 2) ^ uint32_t(r & 7)) << 4) + uint32_t(k & 3) * 4u;
}
// hi = x rounded to nearest tf32 (unbiased, unlike the hardware's truncation), lo = tf32(x - hi)
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_store(uint8_t* hi, uint8_t* lo, uint32_t off, float v) {
  const float h = tf32_rna(v);
  *reinterpret_cast<float*>(hi + off) = h;
  *reinterpret_cast<float*>(lo + off) = tf32_rna(v - h);
}

// Stages one operand k-block: rows [r0, r0+nrows) x k [k0, k0+32) -> hi/lo tiles.
template <bool KC>
__device__ __forceinline__ void stage_operand(const float* __restrict__ p, int64_t ld, int64_t r0, int64_t rmax, int nrows,
                                              int64_t k0, int64_t kmax, uint8_t* hi, uint8_t* lo) {
  const int t = threadIdx.x;
  if (KC) {
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ((k0 & 3) == 0);
    // thread -> (row = t/8 + 32*i, 16-byte chunk = t%8)
    const int c = t & 7;
    for (int r = t >> 3; r < nrows; r += 32) {
      const int64_t gr = r0 + r, gk = k0 + c * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < rmax) {
        const float* src = p + gr * ld + gk;
        if (vec && gk + 3 < kmax) {
          v = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          if (gk < kmax) v.x = __ldg(src);
          if (gk + 1 < kmax) v.y = __ldg(src + 1);
          if (gk + 2 < kmax) v.z = __ldg(src + 2);
          if (gk + 3 < kmax) v.w = __ldg(src + 3);
        }
      }
      const uint32_t off = uint32_t(r) * 128u + (uint32_t(c ^ (r & 7)) << 4);
      float4 h, l;
      h.x = tf32_rna(v.x); l.x = tf32_rna(v.x - h.x);
      h.y = tf32_rna(v.y); l.y = tf32_rna(v.y - h.y);
      h.z = tf32_rna(v.z); l.z = tf32_rna(v.z - h.z);
      h.w = tf32_rna(v.w); l.w = tf32_rna(v.w - h.w);
      *reinterpret_cast<float4*>(hi + off) = h;
      *reinterpret_cast<float4*>(lo + off) = l;
    }
  } else {
    // MN-contiguous in global memory: lanes run along the rows (coalesced), warps along k; transposed store
    const int lane = t & 31, w = t >> 5;
    for (int k = w; k < TK; k += 8) {
      const int64_t gk = k0 + k;
      for (int r = lane; r < nrows; r += 32) {
        const int64_t gr = r0 + r;
        float v = 0.f;
        if (gr < rmax && gk < kmax) v = __ldg(p + gk * ld + gr);
        split_store(hi, lo, sw128(r, k), v);
      }
    }
  }
}

template <bool A_KC, bool B_KC, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(TcArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * tc_stage_bytes());   // [0..1] stage free, [2] accumulator done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int bn = g.bn;
  const int64_t m0 = int64_t(blockIdx.x) * TM, n0 = int64_t(blockIdx.y) * bn;
  const int64_t kbeg = int64_t(blockIdx.z) * g.k_per_split;
  const int64_t kend = min(g.K, kbeg + g.k_per_split);
  if (kbeg >= kend) return;
  const int nkb = int((kend - kbeg + TK - 1) / TK);
  const int64_t nb_left = ((g.N - n0 + 15) / 16) * 16;
  const int nrows_b = int(nb_left < int64_t(bn) ? nb_left : int64_t(bn));   // rows of B actually staged (rest of the tile unused)
  uint32_t tmem_cols = 32;
  while (tmem_cols < uint32_t(bn)) tmem_cols <<= 1;

  if (t == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t idesc = make_idesc_tf32(nrows_b);

  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb & 1;
    uint8_t* st = smem + s * tc_stage_bytes();
    uint8_t *a_hi = st, *a_lo = st + A_TILE_BYTES, *b_hi = st + 2 * A_TILE_BYTES, *b_lo = st + 2 * A_TILE_BYTES + B_TILE_BYTES;
    if (kb >= 2) mbar_wait(smem_u32(&bars[s]), uint32_t(((kb >> 1) - 1) & 1));   // MMAs that read this stage are done
    const int64_t k0 = kbeg + int64_t(kb) * TK;
    stage_operand<A_KC>(g.A, g.lda, m0, g.M, TM, k0, kend, a_hi, a_lo);
    stage_operand<B_KC>(g.B, g.ldb, n0, g.N, nrows_b, k0, kend, b_hi, b_lo);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t dah = make_desc(smem_u32(a_hi)), dal = make_desc(smem_u32(a_lo));
      const uint64_t dbh = make_desc(smem_u32(b_hi)), dbl = make_desc(smem_u32(b_lo));
#pragma unroll
      for (int j = 0; j < TK / 8; ++j) {
        const uint64_t adv = uint64_t((j * 32) >> 4);     // 8 tf32 = 32 bytes along K inside the swizzle row
        umma_tf32(tmem_d, dal + adv, dbh + adv, idesc, (kb | j) ? 1u : 0u);
        umma_tf32(tmem_d, dah + adv, dbl + adv, idesc, 1u);
        umma_tf32(tmem_d, dah + adv, dbh + adv, idesc, 1u);
      }
      umma_commit(smem_u32(&bars[s]));
      if (kb == nkb - 1) umma_commit(smem_u32(&bars[2]));
    }
  }

  // ---- epilogue: TMEM -> registers -> global --------------------------------------------------
  mbar_wait(smem_u32(&bars[2]), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int q = warp & 3;                       // TMEM lane quarter this warp may access
  const int half = warp >> 2;                   // warps w and w+4 share a quarter: split the column chunks
  const int64_t gm = m0 + q * 32 + lane;
  const int nchunks = (nrows_b + 31) / 32;
  for (int c = half; c < nchunks; c += 2) {
    uint32_t v[32];
    const uint32_t taddr = tmem_d + (uint32_t(q * 32) << 16) + uint32_t(c * 32);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (gm < g.M) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int64_t gn = n0 + c * 32 + j;
        if (gn < g.N) {
          float x = __uint_as_float(v[j]);
          if (EPI == TC_EPI_FWD) {
            if (g.bias) x += __ldg(g.bias + gn);
            g.C[gm * g.ldc + gn] = tc_act_fwd(x, g.act, g.act_param);
          } else if (EPI == TC_EPI_DGRAD) {
            if (g.yprev) x *= tc_act_bwd_from_y(__ldg(g.yprev + gm * g.ld_yprev + gn), g.act_prev, g.act_prev_param);
            g.C[gm * g.ldc + gn] = x;
          } else {
            atomicAdd(g.C + gm * g.ldc + gn, x);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
  }
}

template <bool A_KC, bool B_KC, int EPI>
static int launch_tc(const TcArgs& g, dim3 grid, cudaStream_t s, const char* what) {
  static bool configured = false;
  auto kern = tc_gemm_kernel<A_KC, B_KC, EPI>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e));
      return MMSB_E_CUDA;
    }
    configured = true;
  }
  kern<<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(g);
  return check_launch(what);
}

static int pick_bn(int64_t n) {
  int64_t r = ((n + 15) / 16) * 16;
  return int(r > 256 ? 256 : r);
}

// ---- entry points used by mlp_simt.cu's dispatcher ---------------------------------------------
int tc_linear_fwd(const float* x, int64_t ldx, const float* w, const float* b, float* y, int64_t ldy, int64_t n,
                  int in_dim, int out_dim, int act, float act_param, cudaStream_t s) {
  TcArgs g{};
  g.A = x; g.lda = ldx; g.B = w; g.ldb = in_dim; g.C = y; g.ldc = ldy;
  g.M = n; g.N = out_dim; g.K = in_dim; g.bias = b; g.act = act; g.act_param = act_param;
  g.k_per_split = in_dim; g.bn = pick_bn(out_dim);
  dim3 grid((unsigned)ceil_div(n, TM), (unsigned)ceil_div(out_dim, g.bn), 1);
  return launch_tc<true, true, TC_EPI_FWD>(g, grid, s, "linear_fwd(tcgen05)");
}

int tc_linear_bwd_data(const float* dz, int64_t lddz, const float* w, float* dx, int64_t lddx, const float* y_prev,
                       int64_t ld_yprev, int act_prev, float act_prev_param, int64_t n, int in_dim, int out_dim,
                       cudaStream_t s) {
  TcArgs g{};
  g.A = dz; g.lda = lddz; g.B = w; g.ldb = in_dim; g.C = dx; g.ldc = lddx;
  g.M = n; g.N = in_dim; g.K = out_dim;
  g.yprev = y_prev; g.ld_yprev = ld_yprev; g.act_prev = act_prev; g.act_prev_param = act_prev_param;
  g.k_per_split = out_dim; g.bn = pick_bn(in_dim);
  dim3 grid((unsigned)ceil_div(n, TM), (unsigned)ceil_div(in_dim, g.bn), 1);
  return launch_tc<true, false, TC_EPI_DGRAD>(g, grid, s, "linear_bwd_data(tcgen05)");
}

int tc_linear_bwd_weight(const float* dz, int64_t lddz, const float* x, int64_t ldx, float* dw, int64_t n, int in_dim,
                         int out_dim, cudaStream_t s) {
  TcArgs g{};
  g.A = dz; g.lda = lddz; g.B = x; g.ldb = ldx; g.C = dw; g.ldc = in_dim;
  g.M = out_dim; g.N = in_dim; g.K = n; g.bn = pick_bn(in_dim);
  const int64_t tiles = ceil_div(out_dim, TM) * ceil_div(in_dim, g.bn);
  int64_t splits = ceil_div(int64_t(2) * kNumSMs, tiles);
  const int64_t max_splits = ceil_div(n, 4 * TK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t kps = ceil_div(ceil_div(n, splits), TK) * TK;
  g.k_per_split = kps;
  dim3 grid((unsigned)ceil_div(out_dim, TM), (unsigned)ceil_div(in_dim, g.bn), (unsigned)ceil_div(n, kps));
  return launch_tc<false, false, TC_EPI_ATOMIC>(g, grid, s, "linear_bwd_weight(tcgen05)");
}

}  // namespace mmsb
