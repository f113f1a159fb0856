// SURVEY 8(f) row 2 — on-device pixel sampling and target gather.
// ref: src/cameras/pixel_samplers.py:71-89 (UniformPixelSampler: three torch.randint draws (camera, y, x) per batch on
// the CPU), src/data/dataloaders.py:164-167 (advanced-index gather of the pixel values, then a host-to-device copy).
// Here: one thread per ray draws (camera, y, x) from a counter-based generator (Philox-4x32-10 keyed by the seed,
// counter = (ray, step, stream id)) and gathers the target from a device-resident frame stack, so a step has no
// host-to-device crossing.  The draws are uniform like the reference's but not the same sequence (torch's CPU
// generator is a Mersenne twister): distribution parity, not bit parity — tests check range, determinism, uniformity
// and that targets equal frames[cam, y, x] exactly.
#include "common.cuh"

namespace mmsb {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

__global__ void __launch_bounds__(256) sample_pixels_kernel(uint64_t seed, uint32_t step, uint32_t stream_id, int n_cam, int h,
                                                            int w, const float* __restrict__ frames, int channels,
                                                            int32_t* __restrict__ coords, float* __restrict__ targets,
                                                            int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 r = philox4x32_10(make_uint4(uint32_t(i), uint32_t(uint64_t(i) >> 32), step, stream_id),
                                make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
  // multiply-shift maps a 32-bit draw onto [0, range): bias < range / 2^32
  const int cam = int(__umulhi(r.x, uint32_t(n_cam)));
  const int y = int(__umulhi(r.y, uint32_t(h)));
  const int x = int(__umulhi(r.z, uint32_t(w)));
  coords[3 * i] = cam;
  coords[3 * i + 1] = y;
  coords[3 * i + 2] = x;
  if (frames != nullptr) {
    const float* px = frames + ((int64_t(cam) * h + y) * w + x) * channels;
    for (int c = 0; c < channels; ++c) targets[i * channels + c] = __ldg(px + c);
  }
}

}  // namespace mmsb

using namespace mmsb;

extern "C" int mmsb_sample_pixels(uint64_t seed, int32_t step, int32_t stream_id, int32_t n_cam, int32_t height, int32_t width,
                                  const float* frames, int32_t channels, int32_t* coords, float* targets, int64_t n,
                                  mmsb_stream_t stream) {
  MMSB_REQUIRE(n >= 0 && n_cam >= 1 && height >= 1 && width >= 1 && step >= 0, "sample_pixels: bad sizes");
  MMSB_REQUIRE(!frames || (channels >= 1 && targets), "sample_pixels: frames need channels >= 1 and a target buffer");
  if (n == 0) return MMSB_OK;
  MMSB_REQUIRE(coords != nullptr, "sample_pixels: NULL pointer");
  sample_pixels_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(seed, uint32_t(step), uint32_t(stream_id),
                                                                                  n_cam, height, width, frames, channels,
                                                                                  coords, targets, n);
  return check_launch("sample_pixels");
}
