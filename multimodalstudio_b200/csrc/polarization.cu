// A17 — polarization head post-processing: Stokes vector -> the four polarizer intensities seen by the camera.
// ref: src/field_components/field_heads.py:90-106 (leaky_relu on S0), src/model_components/polarizer.py:39-101
// (align_polarization_filters: Mueller rotation by 2 theta, theta = acos(clamp(n_r . up)) - pi/2 with
// n_r = normalize(d x z); stokes_to_intensity: I = 1/2 [[1,1,0],[1,0,1],[1,-1,0],[1,0,-1]] S').
// One thread per sample row; forward and backward (d stokes, d directions, d up) without the [N,3,3] matrices the
// reference materialises.
#include "common.cuh"

namespace mmsb {

constexpr float kLeaky = 0.01f;       // torch.nn.functional.leaky_relu default negative_slope
constexpr float kClamp = 1.0f - 1e-4f;

struct PolGeom {
  float nx, ny, r, u, c, s;
  bool clamped;
};

__device__ __forceinline__ PolGeom pol_geometry(float dx, float dy, float ux, float uy) {
  PolGeom g;
  // d x (0,0,1) = (d_y, -d_x, 0); F.normalize: v / max(|v|, 1e-12)
  const float norm = sqrtf(dy * dy + dx * dx);
  g.r = fmaxf(norm, 1e-12f);
  g.nx = dy / g.r;
  g.ny = -dx / g.r;
  const float raw = g.nx * ux + g.ny * uy;
  g.clamped = raw < -kClamp || raw > kClamp;
  g.u = fminf(fmaxf(raw, -kClamp), kClamp);
  const float theta = acosf(g.u) - 1.5707963267948966f;
  g.c = cosf(2.f * theta);
  g.s = sinf(2.f * theta);
  return g;
}

__global__ void __launch_bounds__(256) polarization_fwd_kernel(const float* __restrict__ stokes, int64_t lds,
                                                               const float* __restrict__ dirs,
                                                               const float* __restrict__ up, float* __restrict__ out,
                                                               int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s0r = __ldg(stokes + i * lds), s1 = __ldg(stokes + i * lds + 1), s2 = __ldg(stokes + i * lds + 2);
  const float s0 = s0r > 0.f ? s0r : kLeaky * s0r;
  const PolGeom g = pol_geometry(__ldg(dirs + 3 * i), __ldg(dirs + 3 * i + 1), __ldg(up + 3 * i), __ldg(up + 3 * i + 1));
  const float a1 = g.c * s1 + g.s * s2;
  const float a2 = -g.s * s1 + g.c * s2;
  float4 o;
  o.x = 0.5f * (s0 + a1);
  o.y = 0.5f * (s0 + a2);
  o.z = 0.5f * (s0 - a1);
  o.w = 0.5f * (s0 - a2);
  reinterpret_cast<float4*>(out)[i] = o;
}

__global__ void __launch_bounds__(256) polarization_bwd_kernel(const float* __restrict__ stokes, int64_t lds,
                                                               const float* __restrict__ dirs,
                                                               const float* __restrict__ up,
                                                               const float* __restrict__ d_out,
                                                               float* __restrict__ d_stokes, float* __restrict__ d_dirs,
                                                               float* __restrict__ d_up, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s0r = __ldg(stokes + i * lds), s1 = __ldg(stokes + i * lds + 1), s2 = __ldg(stokes + i * lds + 2);
  const float dx = __ldg(dirs + 3 * i), dy = __ldg(dirs + 3 * i + 1);
  const float ux = __ldg(up + 3 * i), uy = __ldg(up + 3 * i + 1);
  const PolGeom g = pol_geometry(dx, dy, ux, uy);
  const float4 go = __ldg(reinterpret_cast<const float4*>(d_out) + i);
  const float g0 = 0.5f * (go.x + go.y + go.z + go.w);
  const float ga1 = 0.5f * (go.x - go.z);
  const float ga2 = 0.5f * (go.y - go.w);
  d_stokes[3 * i] = g0 * (s0r > 0.f ? 1.f : kLeaky);
  d_stokes[3 * i + 1] = g.c * ga1 - g.s * ga2;
  d_stokes[3 * i + 2] = g.s * ga1 + g.c * ga2;
  if (d_dirs == nullptr && d_up == nullptr) return;
  // a1 = c s1 + s s2, a2 = -s s1 + c s2
  const float dc = ga1 * s1 + ga2 * s2;
  const float ds = ga1 * s2 - ga2 * s1;
  const float dtheta = -2.f * g.s * dc + 2.f * g.c * ds;
  const float du = g.clamped ? 0.f : -dtheta / sqrtf(fmaxf(1.f - g.u * g.u, 1e-20f));
  if (d_up) {
    d_up[3 * i] = du * g.nx;
    d_up[3 * i + 1] = du * g.ny;
    d_up[3 * i + 2] = 0.f;
  }
  if (d_dirs) {
    // n_r = v / r, v = (d_y, -d_x, 0): dv = (dn - n (n . dn)) / r (zero below the 1e-12 clamp of F.normalize)
    const float dnx = du * ux, dny = du * uy;
    const float dot = g.nx * dnx + g.ny * dny;
    const bool live = g.r > 1e-12f;
    const float dvx = live ? (dnx - g.nx * dot) / g.r : dnx / g.r;
    const float dvy = live ? (dny - g.ny * dot) / g.r : dny / g.r;
    d_dirs[3 * i] = -dvy;
    d_dirs[3 * i + 1] = dvx;
    d_dirs[3 * i + 2] = 0.f;
  }
}

}  // namespace mmsb

using namespace mmsb;

extern "C" int mmsb_polarization_fwd(const float* stokes, int64_t ld_stokes, const float* directions, const float* up_directions,
                                     float* out, int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && ld_stokes >= 3, "polarization_fwd: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(stokes && directions && up_directions && out, "polarization_fwd: NULL pointer");
  MMSB_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "polarization_fwd: out must be 16-byte aligned");
  polarization_fwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(stokes, ld_stokes, directions,
                                                                                     up_directions, out, n);
  return check_launch("polarization_fwd");
}

extern "C" int mmsb_polarization_bwd(const float* stokes, int64_t ld_stokes, const float* directions, const float* up_directions,
                                     const float* d_out, float* d_stokes, float* d_directions, float* d_up_directions,
                                     int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && ld_stokes >= 3, "polarization_bwd: bad sizes");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(stokes && directions && up_directions && d_out && d_stokes, "polarization_bwd: NULL pointer");
  MMSB_REQUIRE((reinterpret_cast<uintptr_t>(d_out) & 15) == 0, "polarization_bwd: d_out must be 16-byte aligned");
  polarization_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
      stokes, ld_stokes, directions, up_directions, d_out, d_stokes, d_directions, d_up_directions, n);
  return check_launch("polarization_bwd");
}
