// Shared pieces of the tcgen05 kernels (mlp_tc.cu, sdf_fused.cu): tile constants, PTX wrappers, the fp16-split helpers,
// activation functions, CTA-pair helpers and the host-side tensor-map / launch-attribute helpers.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace mmsb {
namespace tc {


// The fragment-layout epilogue (tcgen05.ld.16x256b, no shared-memory transpose) measures the same as the transposed one;
// compiling both into the kernels only costs instruction-cache footprint, so it is a build-time option
// (make EXTRA=-DMMSB_TC_FRAG_EPILOGUE=1, then MMSB_TC_DIRECT=2 selects it).
#ifndef MMSB_TC_PAIR_PROD_WARPS
#define MMSB_TC_PAIR_PROD_WARPS 4
#endif
#ifndef MMSB_TC_EPI_WARPS
#define MMSB_TC_EPI_WARPS 8
#endif
#ifndef MMSB_TC_EPI_ILP
#define MMSB_TC_EPI_ILP 4
#endif
#ifndef MMSB_TC_FRAG_EPILOGUE
#define MMSB_TC_FRAG_EPILOGUE 0
#endif

constexpr int TM = 128;            // rows of one accumulator (UMMA M)
constexpr int TK = 32;             // fp32 per k-block = one 128-byte swizzle row
constexpr int NT = 256;            // widest accumulator (UMMA N)
constexpr int PART = TM * 128;     // bytes of one A part (hi or lo) of a stage
constexpr int BPART = NT * 128;    // bytes reserved for one B part of a stage
constexpr int EPI_WARPS = MMSB_TC_EPI_WARPS, PROD_WARPS = 8;   // epilogue warps: a multiple of 4 (TMEM lane quadrants); 2 producer groups
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int THREADS = (EPI_WARPS + PROD_WARPS + 2) * 32;   // rows kernel: + B-loader warp + MMA warp
// pair kernels: 4 converter warps -> 14 warps per CTA: at most 4 warps per SM sub-partition, i.e. up to 128 registers per
// thread (18 warps put 5 on one sub-partition: 96), which the dgrad epilogue needs to read TMEM ahead
constexpr int PAIR_PROD_WARPS = MMSB_TC_PAIR_PROD_WARPS;
constexpr int PAIR_PROD_THREADS = PAIR_PROD_WARPS * 32;
constexpr int PAIR_THREADS = (EPI_WARPS + PAIR_PROD_WARPS + 2) * 32;
constexpr int WG_STAGE_WARPS = 16;                           // weight-gradient kernel: 2 groups of 8 staging warps
constexpr int WG_THREADS = (WG_STAGE_WARPS + 2) * 32;        // + MMA warp + TMA loader warp
constexpr int CH = 16;                                       // accumulator columns per epilogue chunk
constexpr int STG_LD = 16;                                   // floats per row of the epilogue transpose buffer (chunks XOR-swizzled)
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_LD * 4;
constexpr int MAX_STAGES = 4;

__host__ __device__ constexpr int stage_bytes(int nparts) { return nparts * (PART + BPART); }
__host__ __device__ constexpr int num_stages(int nparts) { return nparts == 2 ? 2 : 4; }
__host__ __device__ constexpr int smem_bytes(int nparts) {
  return num_stages(nparts) * stage_bytes(nparts) + STG_BYTES + 256 /*barriers*/ + 1024 /*alignment*/;
}

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// Explicit shared-space accesses: the staging pointers are derived from an integer-aligned base, which hides the
// address space from the compiler (it would emit generic LD / ST: slower path, long-scoreboard latency).
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// Bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the GPU.
// MMSB_TC_WAIT_BACKOFF_NS > 0: the polling loop sleeps between tries (see mbar_wait_backoff); the MMA issuers keep the
// tight loop (mbar_wait_spin): their wake-up latency is on the tensor pipe's critical path.
#ifndef MMSB_TC_WAIT_BACKOFF_NS
#define MMSB_TC_WAIT_BACKOFF_NS 0
#endif
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait_impl(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done) {
      if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
      if (spins > (SLEEP_NS > 0 ? (1u << 22) : (1u << 24))) {
        printf("mms_b200 tcgen05: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
               bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_impl<MMSB_TC_WAIT_BACKOFF_NS>(bar, parity); }
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) { mbar_wait_impl<0>(bar, parity); }
// The same wait with a back-off: a warp that polls in a tight loop takes issue slots from the warps that work (measured
// on the fused SDF kernel: a third of all issued instructions were polling).  SLEEP_NS ~ the latency the waiter can
// afford: tens of ns on the MMA hand-shake path, hundreds for prefetching roles.
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) { mbar_wait_impl<SLEEP_NS>(bar, parity); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 2-D tensor-map copy (TMA): box {TK floats, TM rows} at (k0, m0) -> shared memory in the map's swizzle, bytes onto bar
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int k0, int m0, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(k0), "r"(m0), "r"(bar)
               : "memory");
}
// L2 prefetch of one tensor-map box (no shared memory, no barrier): issued one tile ahead of the copies so that the
// ring's refill latency is an L2 hit instead of an HBM access (the ring alone keeps too few bytes in flight per SM)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int k0, int m0) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(k0), "r"(m0) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Shared-memory matrix descriptor (sm_100: version 1), SWIZZLE_128B.
//  K-major : rows of 128 B (32 fp32 along K), 8-row atoms of 1024 B; SBO = 1024, LBO unused.
//  MN-major (32-bit operands only come as SWIZZLE_128B_BASE32B): rows of 128 B (32 fp32 along M/N) per k, whose
//            32-byte chunks are XOR-swizzled with k % 4; 4-k atoms of 512 B (SBO), 32-wide panels LBO apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4);
  d |= uint64_t(lbo_bytes >> 4) << 16;
  d |= uint64_t(sbo_bytes >> 4) << 32;
  d |= uint64_t(1) << 46;   // descriptor version (Blackwell)
  d |= uint64_t(layout) << 61;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
// c_format F32 (1 << 4), a/b format TF32 (2 << 7, 2 << 10), a/b major at bits 15/16 (1 = MN-major),
// N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc_tf32(int n, bool mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? (3u << 15) : 0u) | (uint32_t(n >> 3) << 17) |
         (uint32_t(TM >> 4) << 24);
}
// fp32 -> TF32 (10 explicit mantissa bits), round to nearest, ties away from zero: what cvt.rna.tf32.f32 computes for
// finite inputs, written as one integer add and one mask (full-rate pipes; the cvt instruction is a quarter-rate
// conversion and the operand producers execute 2 of them per element on the MMA's critical hand-shake path).
__device__ __forceinline__ float tf32_rna(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// ---- precision 2: 2-term fp16 split --------------------------------------------------------------------------------
// x (fp32) is scaled by a power of two s derived from the tensor's max |x| (s * amax in [2^14, 2^15): no fp16 overflow,
// and every element down to amax * 2^-18 keeps a normal `lo`), then hi = fp16(s x) (11 significant bits), lo = fp16(s x -
// hi): 22 significant bits like 3xTF32, products hi*hi + hi*lo + lo*hi as three kind::f16 MMAs (twice the TF32 rate,
// K = 16 per instruction), and the accumulator is multiplied by the exact 1 / (s_a s_b) in the epilogue.
__device__ __forceinline__ int f16_scale_exp(float amax) {
  const uint32_t e = (__float_as_uint(amax) >> 23) & 0xFFu;      // biased exponent of amax >= 0
  if (e == 0u || e == 0xFFu) return 0;                           // zero / denormal / non-finite: no scaling
  int k = 14 + 127 - int(e);
  return k > 120 ? 120 : (k < -120 ? -120 : k);
}
__device__ __forceinline__ float pow2f(int k) { return __uint_as_float(uint32_t(k + 127) << 23); }
// two scaled fp32 values -> packed fp16x2 hi and lo (element 0 in the low half)
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  a = fminf(fmaxf(a, -65504.f), 65504.f);
  b = fminf(fmaxf(b, -65504.f), 65504.f);
  // round to 11 significant bits with one integer add and one mask (exact for fp16-normal magnitudes; below 2^-14 the
  // conversion rounds again, an absolute error of at most 2^-25 of the scaled value, i.e. 2^-39 of the tensor's max)
  const float ha = __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xFFFFE000u);
  const float hb = __uint_as_float((__float_as_uint(b) + 0x1000u) & 0xFFFFE000u);
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(hb), "f"(ha));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hb), "f"(a - ha));
}
// the same split for values the caller has bounded below 2^16 in magnitude (no clamp)
__device__ __forceinline__ void split_f16x2_bounded(float a, float b, uint32_t& hi, uint32_t& lo) {
  const float ha = __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xFFFFE000u);
  const float hb = __uint_as_float((__float_as_uint(b) + 0x1000u) & 0xFFFFE000u);
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(hb), "f"(ha));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hb), "f"(a - ha));
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// c_format F32 (1 << 4), a/b format F16 (0), K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc_f16(int n, int m) {
  return (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
constexpr int TK16 = 64;           // fp16 elements per k-block = one 128-byte swizzle row

// byte offset of 16-byte chunk c (4 fp32 along M/N) of k-row k in an MN-major SWIZZLE_128B_BASE32B tile whose 32-wide
// panels hold TK k-rows each
__device__ __forceinline__ uint32_t mn_offset(int c, int k) {
  return uint32_t(c >> 3) * uint32_t(TK * 128) + uint32_t(k) * 128u + (uint32_t(((c & 7) >> 1) ^ (k & 3)) << 5) +
         (uint32_t(c & 1) << 4);
}
template <int NPARTS>
__device__ __forceinline__ void split_store4(uint8_t* hi, uint8_t* lo, uint32_t off, const float4& v) {
  float4 h;
  h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
  sts128(smem_u32(hi) + off, h);
  if (NPARTS == 2) {
    float4 l;
    l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
    sts128(smem_u32(lo) + off, l);
  }
}
// 16 bytes at (row r, k..k+3) of P[r*ld + k]; zero outside [0,rmax) x [0,kmax)
__device__ __forceinline__ float4 load4(const float* __restrict__ p, int64_t ld, int64_t r, int64_t rmax, int64_t k,
                                        int64_t kmax, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < rmax && k < kmax) {
    const float* src = p + r * ld + k;
    if (vec && k + 3 < kmax) {
      v = __ldg(reinterpret_cast<const float4*>(src));
    } else {
      v.x = __ldg(src);
      if (k + 1 < kmax) v.y = __ldg(src + 1);
      if (k + 2 < kmax) v.z = __ldg(src + 2);
      if (k + 3 < kmax) v.w = __ldg(src + 3);
    }
  }
  return v;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 256 bits (8 fp32 columns) per repetition, the mma C-fragment layout: thread t holds (row t / 4, columns
// 2 (t % 4), +1) in v[0..1] and (row t / 4 + 8, same columns) in v[2..3]; the second repetition (v[4..7]) is the next 8
// columns.  A quad of lanes therefore owns a whole 32-byte sector of a row: global stores need no transpose.
__device__ __forceinline__ void tmem_ld_frag(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// The epilogues evaluate Softplus through the MUFU units (ex2 / lg2 approximations, ~2^-22 relative): their error is
// far below the 3xTF32 product error, and the IEEE expf / log1pf sequences would cost more issue slots per tile than
// the whole MMA mainloop leaves free.
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float act_fwd(float z, int act, float p) {
  switch (act) {
    case MMSB_ACT_RELU: return fmaxf(z, 0.f);
    case MMSB_ACT_SOFTPLUS: {
      // max(z, 0) + log1p(exp(-|beta z|)) / beta; equals z exactly for beta z > 20 like torch's threshold
      const float e = fast_ex2(-fabsf(z * p) * 1.4426950408889634f);
      return fmaxf(z, 0.f) + fast_lg2(1.f + e) * (0.6931471805599453f / p);
    }
    case MMSB_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    default: return z;
  }
}
__device__ __forceinline__ float act_bwd_from_y(float y, int act, float p) {
  switch (act) {
    case MMSB_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case MMSB_ACT_SOFTPLUS: return 1.f - fast_ex2(-y * p * 1.4426950408889634f);      // sigmoid(beta z) = 1 - exp(-beta y)
    case MMSB_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}

// ---- packed-weight geometry ------------------------------------------------------------------------------
__host__ __device__ inline int pad16(int n) { return (n + 15) / 16 * 16; }
__host__ __device__ inline int tile_width(int n_pad, int nt) { return n_pad - nt * NT < NT ? n_pad - nt * NT : NT; }
__host__ __device__ inline int64_t packed_floats(int n, int k, int nparts) {
  return int64_t(pad16(n)) * TK * ((k + TK - 1) / TK) * nparts;
}

__host__ __device__ inline int64_t packed_floats_f16(int n, int k) {
  return int64_t(pad16(n)) * ((k + TK16 - 1) / TK16) * 64;      // hi 128 B + lo 128 B per (row, k-block); + 4 trailer floats
}
// ---- CTA pairs ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the pair's even CTA
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(leader_bar & kPeerBitMask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(uint16_t(3))
               : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// ---- host helpers ---------------------------------------------------------------------------------------
// cudaFuncSetAttribute applies to the CURRENT device: a process that drives several GPUs must configure each one.
// One flag per (kernel instantiation, device ordinal); set once, racing threads at worst repeat an idempotent call.
struct PerDeviceFlag {
  bool done[64] = {};
  bool& operator()() {
    int dev = 0;
    cudaGetDevice(&dev);
    return done[dev & 63];
  }
};

template <typename K>
static int set_smem(K kern, int bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute failed: %s", what, cudaGetErrorString(e));
    return MMSB_E_CUDA;
  }
  return MMSB_OK;
}

typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// the driver entry point is resolved through the runtime (no link-time dependency on libcuda: the library must load, and
// export its symbols, on a machine without a driver)
static TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn encode = nullptr;
  static bool resolved = false;
  if (!resolved) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<TensorMapEncodeFn>(fn);
    resolved = true;
  }
  return encode;
}

// Tensor map over the packed weight buffer seen as rows of 32 floats (128 B, already swizzled by the packer): box = 128
// rows (one CTA's half of a hi or lo tile).
static bool make_packed_map(const float* packed, int64_t total_rows, CUtensorMap* map) {
  TensorMapEncodeFn encode = tensor_map_encoder();
  if (encode == nullptr || (reinterpret_cast<uintptr_t>(packed) & 15) != 0 || total_rows >= (int64_t(1) << 31)) return false;
  const cuuint64_t dims[2] = {32, cuuint64_t(total_rows)};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(packed), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace mmsb
