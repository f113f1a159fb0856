// A1/A2 — ray generation with per-modality intrinsics, OpenCV distortion (Newton undistort) and the
// SO3xR3 pose refinement, forward and backward to the pose parameters.
// ref: src/cameras/camera_optimizers.py:86-119, src/cameras/lie_groups.py:28-63,
//      src/model_components/ray_generators.py:54-81, src/cameras/cameras.py:534-703,
//      src/cameras/camera_utils.py:279-383, src/utils/poses.py:53-67
#include "common.cuh"

namespace mmsb {

// 10 Newton iterations on the OpenCV (k1,k2,k3,k4,p1,p2) model; step zeroed when |det| <= 1e-3.
__device__ __forceinline__ void undistort(float xd, float yd, const float* __restrict__ dp, float& xo, float& yo) {
  const float k1 = dp[0], k2 = dp[1], k3 = dp[2], k4 = dp[3], p1 = dp[4], p2 = dp[5];
  float x = xd, y = yd;
#pragma unroll 1
  for (int it = 0; it < 10; ++it) {
    const float r = x * x + y * y;
    const float d = 1.0f + r * (k1 + r * (k2 + r * (k3 + r * k4)));
    const float fx = d * x + 2.f * p1 * x * y + p2 * (r + 2.f * x * x) - xd;
    const float fy = d * y + 2.f * p2 * x * y + p1 * (r + 2.f * y * y) - yd;
    const float d_r = k1 + r * (2.0f * k2 + r * (3.0f * k3 + r * 4.0f * k4));
    const float d_x = 2.0f * x * d_r, d_y = 2.0f * y * d_r;
    const float fx_x = d + d_x * x + 2.0f * p1 * y + 6.0f * p2 * x;
    const float fx_y = d_y * x + 2.0f * p1 * x + 2.0f * p2 * y;
    const float fy_x = d_x * y + 2.0f * p2 * y + 2.0f * p1 * x;
    const float fy_y = d + d_y * y + 2.0f * p2 * x + 6.0f * p1 * y;
    const float den = fy_x * fx_y - fx_x * fy_y;
    const float xn = fx * fy_y - fy * fx_y, yn = fy * fx_x - fx * fy_x;
    const bool ok = fabsf(den) > 1e-3f;
    x += ok ? xn / den : 0.f;
    y += ok ? yn / den : 0.f;
  }
  xo = x; yo = y;
}

struct Pose {
  float R[3][3];   // composed rotation c2w * delta
  float t[3];      // composed translation
};

struct ExpMap {
  float R2[3][3];
  float K[3][3], K2[3][3];
  float theta, f1, f2;
  bool clamped;
};

__device__ __forceinline__ ExpMap exp_so3(const float* __restrict__ w) {
  ExpMap e;
  const float nrm2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  e.clamped = nrm2 < 1e-4f;
  e.theta = sqrtf(fmaxf(nrm2, 1e-4f));
  const float inv = 1.0f / e.theta;
  e.f1 = inv * sinf(e.theta);
  e.f2 = inv * inv * (1.0f - cosf(e.theta));
  const float K[3][3] = {{0.f, -w[2], w[1]}, {w[2], 0.f, -w[0]}, {-w[1], w[0], 0.f}};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      e.K[i][j] = K[i][j];
      e.K2[i][j] = K[i][0] * K[0][j] + K[i][1] * K[1][j] + K[i][2] * K[2][j];
    }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) e.R2[i][j] = e.f1 * e.K[i][j] + e.f2 * e.K2[i][j] + (i == j ? 1.f : 0.f);
  return e;
}

__device__ __forceinline__ void cam_dirs(const int32_t* __restrict__ coords, const float* __restrict__ intr,
                                         const float* __restrict__ dist, float pixel_offset, int64_t i, int& cam,
                                         float (&v)[3][3]) {
  cam = coords[3 * i];
  const float y = float(coords[3 * i + 1]) + pixel_offset, x = float(coords[3 * i + 2]) + pixel_offset;
  const float fx = intr[4 * cam], fy = intr[4 * cam + 1], cx = intr[4 * cam + 2], cy = intr[4 * cam + 3];
  // ref: cameras.py:574-576 (pixel, +1 in x, +1 in y)
  float pts[3][2] = {{(x - cx) / fx, -(y - cy) / fy}, {(x - cx + 1.f) / fx, -(y - cy) / fy}, {(x - cx) / fx, -(y - cy + 1.f) / fy}};
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    float u = pts[p][0], w = pts[p][1];
    if (dist) undistort(u, w, dist + 6 * cam, u, w);
    v[p][0] = u; v[p][1] = w; v[p][2] = -1.0f;      // cameras.py:620-622
  }
}

__device__ __forceinline__ Pose compose(const float* __restrict__ c2w, int cam, const float* __restrict__ pose_adjust,
                                        int n_pose, ExpMap* em_out) {
  Pose P;
  const float* M = c2w + 12 * cam;
  if (!pose_adjust) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) P.R[i][j] = M[4 * i + j];
      P.t[i] = M[4 * i + 3];
    }
    return P;
  }
  const float* pa = pose_adjust + 6 * (n_pose == 1 ? 0 : cam);
  const ExpMap e = exp_so3(pa + 3);
  if (em_out) *em_out = e;
  // ref: poses.py:63-67  R = R1 R2, t = t1 + R1 t2
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) P.R[i][j] = M[4 * i] * e.R2[0][j] + M[4 * i + 1] * e.R2[1][j] + M[4 * i + 2] * e.R2[2][j];
    P.t[i] = M[4 * i + 3] + (M[4 * i] * pa[0] + M[4 * i + 1] * pa[1] + M[4 * i + 2] * pa[2]);
  }
  return P;
}

__global__ void __launch_bounds__(256) raygen_fwd_kernel(const int32_t* __restrict__ coords, const float* __restrict__ c2w,
                                                         const float* __restrict__ intr, const float* __restrict__ dist,
                                                         const float* __restrict__ pose_adjust, int n_pose,
                                                         float pixel_offset, float* __restrict__ origins,
                                                         float* __restrict__ directions, float* __restrict__ up,
                                                         float* __restrict__ pixel_area, float* __restrict__ dir_norm,
                                                         int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cam;
  float v[3][3];
  cam_dirs(coords, intr, dist, pixel_offset, i, cam, v);
  const Pose P = compose(c2w, cam, pose_adjust, n_pose, nullptr);
  float dw[3][3];
  float nrm0 = 0.f;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    float w[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) w[a] = v[p][0] * P.R[a][0] + v[p][1] * P.R[a][1] + v[p][2] * P.R[a][2];
    const float nr = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    if (p == 0) nrm0 = nr;
    const float den = fmaxf(nr, 1e-12f);
#pragma unroll
    for (int a = 0; a < 3; ++a) dw[p][a] = w[a] / den;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    origins[3 * i + a] = P.t[a];
    directions[3 * i + a] = dw[0][a];
    if (up) up[3 * i + a] = P.R[a][1];     // R (0,1,0)^T, cameras.py:680-682
  }
  if (dir_norm) dir_norm[i] = nrm0;
  if (pixel_area) {
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float ex = dw[0][a] - dw[1][a], ey = dw[0][a] - dw[2][a];
      sx += ex * ex; sy += ey * ey;
    }
    pixel_area[i] = sqrtf(sx) * sqrtf(sy);
  }
}

__global__ void __launch_bounds__(256) raygen_bwd_kernel(const int32_t* __restrict__ coords, const float* __restrict__ c2w,
                                                         const float* __restrict__ intr, const float* __restrict__ dist,
                                                         const float* __restrict__ pose_adjust, int n_pose,
                                                         float pixel_offset, const float* __restrict__ d_origins,
                                                         const float* __restrict__ d_directions,
                                                         const float* __restrict__ d_up, float* __restrict__ d_pose,
                                                         int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int cam = 0;
  if (i < n) {
    float v[3][3];
    cam_dirs(coords, intr, dist, pixel_offset, i, cam, v);
    ExpMap e;
    const Pose P = compose(c2w, cam, pose_adjust, n_pose, &e);
    const float* M = c2w + 12 * cam;
    // dL/dR (composed) from the normalised main direction and the up vector
    float GR[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    if (d_directions) {
      float w[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) w[a] = v[0][0] * P.R[a][0] + v[0][1] * P.R[a][1] + v[0][2] * P.R[a][2];
      const float nr = fmaxf(sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]), 1e-12f);
      const float dh[3] = {w[0] / nr, w[1] / nr, w[2] / nr};
      const float gd[3] = {d_directions[3 * i], d_directions[3 * i + 1], d_directions[3 * i + 2]};
      const float dot = dh[0] * gd[0] + dh[1] * gd[1] + dh[2] * gd[2];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float gw = (gd[a] - dh[a] * dot) / nr;
#pragma unroll
        for (int b = 0; b < 3; ++b) GR[a][b] += gw * v[0][b];
      }
    }
    if (d_up) {
#pragma unroll
      for (int a = 0; a < 3; ++a) GR[a][1] += d_up[3 * i + a];
    }
    // translation: t = t1 + R1 t2  ->  d t2 = R1^T dO
    if (d_origins) {
#pragma unroll
      for (int b = 0; b < 3; ++b)
        g[b] = M[b] * d_origins[3 * i] + M[4 + b] * d_origins[3 * i + 1] + M[8 + b] * d_origins[3 * i + 2];
    }
    // G = dL/dR2 = R1^T GR
    float G[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) G[a][b] = M[a] * GR[0][b] + M[4 + a] * GR[1][b] + M[8 + a] * GR[2][b];
    float df1 = 0.f, df2 = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) { df1 += G[a][b] * e.K[a][b]; df2 += G[a][b] * e.K2[a][b]; }
    // dL/dK = f1 G + f2 (G K^T + K^T G)
    float DK[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        float gk = 0.f, kg = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) { gk += G[a][c] * e.K[b][c]; kg += e.K[c][a] * G[c][b]; }
        DK[a][b] = e.f1 * G[a][b] + e.f2 * (gk + kg);
      }
    g[3] = DK[2][1] - DK[1][2];
    g[4] = DK[0][2] - DK[2][0];
    g[5] = DK[1][0] - DK[0][1];
    if (!e.clamped) {
      const float th = e.theta, sn = sinf(th), cs = cosf(th);
      const float f1p = (th * cs - sn) / (th * th);
      const float f2p = (th * sn - 2.f * (1.f - cs)) / (th * th * th);
      const float dth = df1 * f1p + df2 * f2p;
      const float* wv = pose_adjust + 6 * (n_pose == 1 ? 0 : cam) + 3;
      g[3] += dth * wv[0] / th; g[4] += dth * wv[1] / th; g[5] += dth * wv[2] / th;
    }
  }
  if (n_pose == 1) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float s = warp_sum(g[k]);
      if (lane == 0) atomicAdd(d_pose + k, s);
    }
  } else if (i < n) {
#pragma unroll
    for (int k = 0; k < 6; ++k) atomicAdd(d_pose + 6 * cam + k, g[k]);
  }
}

}  // namespace mmsb

using namespace mmsb;

extern "C" int mmsb_raygen_fwd(const int32_t* coords, const float* c2w, const float* intr, const float* dist,
                               const float* pose_adjust, int32_t n_pose, int32_t n_cam, float pixel_offset,
                               float* origins, float* directions, float* up, float* pixel_area, float* dir_norm,
                               int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && n_cam >= 1, "raygen_fwd: bad sizes");
  MMSB_REQUIRE(!pose_adjust || n_pose == 1 || n_pose == n_cam, "raygen_fwd: n_pose must be 1 or n_cam");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(coords && c2w && intr && origins && directions, "raygen_fwd: NULL pointer");
  raygen_fwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
      coords, c2w, intr, dist, pose_adjust, n_pose, pixel_offset, origins, directions, up, pixel_area, dir_norm, n);
  return check_launch("raygen_fwd");
}

extern "C" int mmsb_raygen_bwd(const int32_t* coords, const float* c2w, const float* intr, const float* dist,
                               const float* pose_adjust, int32_t n_pose, int32_t n_cam, float pixel_offset,
                               const float* d_origins, const float* d_directions, const float* d_up, float* d_pose,
                               int64_t n, mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && n_cam >= 1, "raygen_bwd: bad sizes");
  MMSB_REQUIRE(pose_adjust && (n_pose == 1 || n_pose == n_cam), "raygen_bwd: needs pose_adjust with n_pose 1 or n_cam");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(coords && c2w && intr && d_pose, "raygen_bwd: NULL pointer");
  raygen_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
      coords, c2w, intr, dist, pose_adjust, n_pose, pixel_offset, d_origins, d_directions, d_up, d_pose, n);
  return check_launch("raygen_bwd");
}
