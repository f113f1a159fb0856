// A21/A22 — mosaick channel select fused with the L1 / skip-saturation loss, and the geometry
// (eikonal + curvature) reductions.
// ref: src/pipelines/raw_pipeline.py:112-122, src/data/datasets.py:229-250,
//      src/model_components/losses.py:97-119,143-164,213-265
#include "common.cuh"

namespace mmsb {

__device__ __forceinline__ float block_sum_256(float v) {
  __shared__ float red[8];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
  if (w == 0) t = warp_sum(t);
  return t;   // valid in thread 0
}

__device__ __forceinline__ int band_of(const int32_t* __restrict__ coords, const int32_t* __restrict__ pattern, int ph,
                                       int pw, int64_t i) {
  // mosaick_mask[y, x] with the pattern tiled from the image origin (datasets.py:243-249)
  const int y = coords[3 * i + 1], x = coords[3 * i + 2];
  return pattern[(y % ph) * pw + (x % pw)];
}

// pattern != NULL : raw frames, one supervised channel per pixel, target [n]
// pattern == NULL : demosaicked frames, every channel supervised, target [n, c]
__global__ void __launch_bounds__(256) mosaick_l1_fwd_kernel(const int32_t* __restrict__ coords,
                                                             const int32_t* __restrict__ pattern, int ph, int pw,
                                                             const float* __restrict__ rendered, int c,
                                                             const float* __restrict__ target, float sat_threshold,
                                                             const int64_t* __restrict__ sat_index,
                                                             int64_t* __restrict__ band_out, float* __restrict__ selected,
                                                             float* __restrict__ loss_sum, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  float acc = 0.f;
  float satv = 0.f;
  bool has_sat = false;
  if (sat_index) {
    const int64_t si = *sat_index;
    const int64_t total = pattern ? n : n * c;
    has_sat = si < total;
    if (has_sat) satv = target[si];
  }
  if (i < n) {
    if (pattern) {
      const int b = band_of(coords, pattern, ph, pw, i);
      if (band_out) band_out[i] = b;
      const float sel = rendered[i * c + b];
      if (selected) selected[i] = sel;
      const float t = target[i];
      const float pred = (has_sat && t > sat_threshold) ? satv : sel;
      acc = fabsf(pred - t);
    } else {
      for (int k = 0; k < c; ++k) {
        const float t = target[i * c + k];
        const float pred = (has_sat && t > sat_threshold) ? satv : rendered[i * c + k];
        acc += fabsf(pred - t);
      }
    }
  }
  acc = block_sum_256(acc);
  if (threadIdx.x == 0 && acc != 0.f) atomicAdd(loss_sum, acc);
}

__global__ void __launch_bounds__(256) mosaick_l1_bwd_kernel(const int32_t* __restrict__ coords,
                                                             const int32_t* __restrict__ pattern, int ph, int pw,
                                                             const float* __restrict__ rendered, int c,
                                                             const float* __restrict__ target, float sat_threshold,
                                                             const int64_t* __restrict__ sat_index,
                                                             const float* __restrict__ d_loss, float inv_count,
                                                             float* __restrict__ d_rendered, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float scale = __ldg(d_loss) * inv_count;
  bool has_sat = false;
  if (sat_index) has_sat = *sat_index < (pattern ? n : n * c);
  const int b = pattern ? band_of(coords, pattern, ph, pw, i) : -1;
  for (int k = 0; k < c; ++k) {
    float g = 0.f;
    if (!pattern || k == b) {
      const float t = pattern ? target[i] : target[i * c + k];
      if (!(has_sat && t > sat_threshold)) {
        const float d = rendered[i * c + k] - t;
        g = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
      }
    }
    d_rendered[i * c + k] = g;
  }
}

__global__ void sat_init_kernel(int64_t* sat_index, int64_t n) { *sat_index = n; }
__global__ void __launch_bounds__(256) sat_find_kernel(const float* __restrict__ target, float thr,
                                                       int64_t* __restrict__ sat_index, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n && target[i] > thr) atomicMin(reinterpret_cast<unsigned long long*>(sat_index), (unsigned long long)i);
}

// sums[0] += sum (|g|-1)^2, sums[1] += sum |lap|, sums[2] += number of rows counted
__global__ void __launch_bounds__(256) geometry_loss_fwd_kernel(const float* __restrict__ gradients,
                                                                const float* __restrict__ hessians,
                                                                const uint8_t* __restrict__ ray_mask, int s,
                                                                float* __restrict__ sums, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  float e = 0.f, cu = 0.f, cnt = 0.f;
  if (i < n && (!ray_mask || ray_mask[i / s])) {
    const float gx = gradients[3 * i], gy = gradients[3 * i + 1], gz = gradients[3 * i + 2];
    const float d = sqrtf(gx * gx + gy * gy + gz * gz) - 1.f;
    e = d * d;
    if (hessians) cu = fabsf(hessians[3 * i] + hessians[3 * i + 1] + hessians[3 * i + 2]);
    cnt = 1.f;
  }
  e = block_sum_256(e);
  cu = block_sum_256(cu);
  cnt = block_sum_256(cnt);
  if (threadIdx.x == 0 && cnt != 0.f) {
    atomicAdd(sums, e);
    atomicAdd(sums + 1, cu);
    atomicAdd(sums + 2, cnt);
  }
}

// d_gradients = d_eik * 2 (|g|-1) g/|g| / count ; d_hessians = d_curv * sign(lap) / count
__global__ void __launch_bounds__(256) geometry_loss_bwd_kernel(const float* __restrict__ gradients,
                                                                const float* __restrict__ hessians,
                                                                const uint8_t* __restrict__ ray_mask, int s,
                                                                const float* __restrict__ sums,
                                                                const float* __restrict__ d_eik,
                                                                const float* __restrict__ d_curv,
                                                                float* __restrict__ d_gradients,
                                                                float* __restrict__ d_hessians, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool on = !ray_mask || ray_mask[i / s];
  const float inv = 1.f / fmaxf(sums[2], 1.f);
  float ox = 0.f, oy = 0.f, oz = 0.f, h = 0.f;
  if (on) {
    const float gx = gradients[3 * i], gy = gradients[3 * i + 1], gz = gradients[3 * i + 2];
    const float nr = sqrtf(gx * gx + gy * gy + gz * gz);
    if (nr > 0.f && d_eik) {
      const float k = __ldg(d_eik) * 2.f * (nr - 1.f) / nr * inv;
      ox = k * gx; oy = k * gy; oz = k * gz;
    }
    if (hessians && d_curv) {
      const float lap = hessians[3 * i] + hessians[3 * i + 1] + hessians[3 * i + 2];
      h = (lap > 0.f ? 1.f : (lap < 0.f ? -1.f : 0.f)) * __ldg(d_curv) * inv;
    }
  }
  d_gradients[3 * i] = ox; d_gradients[3 * i + 1] = oy; d_gradients[3 * i + 2] = oz;
  if (d_hessians) { d_hessians[3 * i] = h; d_hessians[3 * i + 1] = h; d_hessians[3 * i + 2] = h; }
}

// ---- A23 (next): fused AdamW and the global-norm reduction ----------------------------------
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v,
                                                    const float* __restrict__ grad_scale, float lr, float beta1,
                                                    float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                                                    int64_t n) {
  const float gs = grad_scale ? __ldg(grad_scale) : 1.f;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float gr = g[i] * gs;
    float pv = p[i];
    pv *= (1.f - lr * wd);                               // decoupled weight decay (torch.optim.AdamW)
    const float mi = beta1 * m[i] + (1.f - beta1) * gr;
    const float vi = beta2 * v[i] + (1.f - beta2) * gr * gr;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pv - (lr / bc1) * (mi / denom);
  }
}

// Same update with the per-step scalars in device memory (hyper = {lr, 1 - beta1^t, sqrt(1 - beta2^t), prescale}) and
// the clip coefficient min(1, max_norm / (sqrt(sumsq) * prescale + 1e-6)) evaluated in the kernel: nothing step-dependent
// is a launch argument, so the launch can be replayed from a CUDA graph.  prescale (1 / world size for DDP's gradient
// mean, 1 otherwise) is folded in here instead of a separate pass over the all-reduced gradient: g := g * prescale
// before the clip, i.e. the norm that is clipped is the norm of the averaged gradient.
__global__ void __launch_bounds__(256) adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        const float* __restrict__ grad_sumsq, float max_norm,
                                                        const float* __restrict__ hyper, float beta1, float beta2,
                                                        float eps, float wd, int64_t n) {
  const float lr = __ldg(hyper), bc1 = __ldg(hyper + 1), bc2_sqrt = __ldg(hyper + 2), pre = __ldg(hyper + 3);
  float gs = pre;
  if (grad_sumsq) gs = pre * fminf(__fdiv_rn(max_norm, __fadd_rn(__fmul_rn(__fsqrt_rn(__ldg(grad_sumsq)), pre), 1e-6f)), 1.f);
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float gr = g[i] * gs;
    float pv = p[i];
    pv *= (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gr;
    const float vi = beta2 * v[i] + (1.f - beta2) * gr * gr;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pv - (lr / bc1) * (mi / denom);
  }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  float acc = 0.f;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float v = x[i];
    acc = fmaf(v, v, acc);
  }
  acc = block_sum_256(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

}  // namespace mmsb

using namespace mmsb;

extern "C" int mmsb_mosaick_l1_fwd(const int32_t* coords, const int32_t* pattern, int32_t ph, int32_t pw,
                                   const float* rendered, int32_t c, const float* target, float sat_threshold,
                                   const int64_t* sat_index, int64_t* band_out, float* selected, float* loss_sum,
                                   int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && c >= 1 && (!pattern || (ph >= 1 && pw >= 1 && coords)), "mosaick_l1_fwd: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(rendered && target && loss_sum, "mosaick_l1_fwd: NULL pointer");
  mosaick_l1_fwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
      coords, pattern, ph, pw, rendered, c, target, sat_threshold, sat_index, band_out, selected, loss_sum, n);
  return check_launch("mosaick_l1_fwd");
}

extern "C" int mmsb_mosaick_l1_bwd(const int32_t* coords, const int32_t* pattern, int32_t ph, int32_t pw,
                                   const float* rendered, int32_t c, const float* target, float sat_threshold,
                                   const int64_t* sat_index, const float* d_loss, float inv_count, float* d_rendered,
                                   int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && c >= 1 && (!pattern || (ph >= 1 && pw >= 1 && coords)), "mosaick_l1_bwd: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(rendered && target && d_loss && d_rendered, "mosaick_l1_bwd: NULL pointer");
  mosaick_l1_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
      coords, pattern, ph, pw, rendered, c, target, sat_threshold, sat_index, d_loss, inv_count, d_rendered, n);
  return check_launch("mosaick_l1_bwd");
}

extern "C" int mmsb_first_saturated(const float* target, float sat_threshold, int64_t* sat_index, int64_t n,
                                    mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && sat_index, "first_saturated: bad arguments");
  sat_init_kernel<<<1, 1, 0, as_stream(stream)>>>(sat_index, n);
  if (int e = check_launch("first_saturated(init)")) return e;
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(target != nullptr, "first_saturated: NULL pointer");
  sat_find_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(target, sat_threshold, sat_index, n);
  return check_launch("first_saturated");
}

extern "C" int mmsb_geometry_loss_fwd(const float* gradients, const float* hessians, const uint8_t* ray_mask, int32_t s,
                                      float* sums, int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && s >= 1, "geometry_loss_fwd: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(gradients && sums, "geometry_loss_fwd: NULL pointer");
  geometry_loss_fwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(gradients, hessians, ray_mask, s,
                                                                                      sums, n);
  return check_launch("geometry_loss_fwd");
}

extern "C" int mmsb_geometry_loss_bwd(const float* gradients, const float* hessians, const uint8_t* ray_mask, int32_t s,
                                      const float* sums, const float* d_eik, const float* d_curv, float* d_gradients,
                                      float* d_hessians, int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && s >= 1, "geometry_loss_bwd: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(gradients && sums && d_gradients, "geometry_loss_bwd: NULL pointer");
  geometry_loss_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
      gradients, hessians, ray_mask, s, sums, d_eik, d_curv, d_gradients, d_hessians, n);
  return check_launch("geometry_loss_bwd");
}

extern "C" int mmsb_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                               const float* grad_scale, float lr, float beta1, float beta2, float eps,
                               float weight_decay, int32_t step, int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && step >= 1, "adamw_step: bad arguments (step is 1-based)");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(param && grad && exp_avg && exp_avg_sq, "adamw_step: NULL pointer");
  const double bc1 = 1.0 - pow(double(beta1), double(step));
  const double bc2 = 1.0 - pow(double(beta2), double(step));
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  adamw_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, grad_scale, lr, beta1,
                                                                beta2, eps, weight_decay, float(bc1), float(sqrt(bc2)), n);
  return check_launch("adamw_step");
}

extern "C" int mmsb_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                   const float* grad_sumsq, float max_norm, const float* hyper, float beta1, float beta2,
                                   float eps, float weight_decay, int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0, "adamw_step_dev: bad arguments");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(param && grad && exp_avg && exp_avg_sq && hyper, "adamw_step_dev: NULL pointer");
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  adamw_dev_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, grad_sumsq, max_norm,
                                                                    hyper, beta1, beta2, eps, weight_decay, n);
  return check_launch("adamw_step_dev");
}

extern "C" int mmsb_sumsq(const float* x, float* sumsq, int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && sumsq, "sumsq: bad arguments");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(x != nullptr, "sumsq: NULL pointer");
  int64_t blocks = ceil_div(n, 1024);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, sumsq, n);
  return check_launch("sumsq");
}
