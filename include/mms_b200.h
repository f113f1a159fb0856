/*
 * mms_b200.h — C ABI of libmms_b200.so: the B200-native (sm_100a) per-ray rendering hot path of
 * MMS-FW (LTTM/MultimodalStudio).  This is the drop-in boundary.
 *
 * The reference is pure Python; its only native slot is the tiny-cuda-nn torch binding
 * (reference src/field_components/encodings.py:218,377 and src/field_components/mlp.py:223,277).
 * Every entry point below replaces the arithmetic of one reference operator; the operator it
 * replaces is cited as `ref: <file>:<lines>` (paths relative to /root/reference/src).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++ / torch types.
 *  - Every pointer is a DEVICE pointer unless the name ends in `_host`.  The caller owns every
 *    buffer (inputs, outputs, workspaces); the library never allocates or frees device memory and
 *    keeps no mutable global state.
 *  - All floating tensors are fp32, contiguous in their last dimension; a leading dimension
 *    (`ld*`, in elements) is given where a tensor may be a column slice of a wider row.
 *  - `stream` is a cudaStream_t passed as void*.  Functions only enqueue work on that stream, never
 *    synchronise, never touch the default stream, and are CUDA-graph capturable.  Ragged counts
 *    (number of rays inside the sphere) live in device memory.
 *  - Return value: 0 on success, a negative MMSB_E_* code otherwise; a human-readable message for
 *    the calling thread is available from mmsb_last_error().
 *  - Re-entrant: may be called from the autograd engine's worker threads.
 */
#ifndef MMS_B200_H
#define MMS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMSB_OK 0
#define MMSB_E_INVALID_ARGUMENT (-1) /* bad shape / size / enum: Python wrapper raises ValueError   */
#define MMSB_E_UNSUPPORTED (-2)      /* valid but not built (e.g. F not in {1,2,4,8}): ValueError     */
#define MMSB_E_CUDA (-3)             /* launch failure: RuntimeError                                 */

typedef void* mmsb_stream_t;

const char* mmsb_version(void);
const char* mmsb_last_error(void);
/* Number of kernels this library has launched from this process (all threads); bench.py reads it
 * around the timed region to report `gpu_launches`. */
int64_t mmsb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * A9/A10  multi-resolution hash-grid encoding
 * ref: field_components/encodings.py:184-310 (HashEncoding.hash_fn/pytorch_fwd),
 *      field_components/feature_structures.py:78-88 (FeatureGrid.forward: rescale + level mask)
 * ---------------------------------------------------------------------------------------------- */
#define MMSB_MAX_LEVELS 32
#define MMSB_INTERP_LINEAR 0     /* reference torch path (pinned) */
#define MMSB_INTERP_SMOOTHSTEP 1 /* tcnn default; parity unpinned (tiny-cuda-nn not in the tree) */

typedef struct {
  int32_t num_levels;              /* L, 1..32 */
  int32_t features_per_level;      /* F in {1,2,4,8} */
  int32_t log2_hashmap_size;       /* table rows per level = 1 << log2_hashmap_size */
  int32_t interpolation;           /* MMSB_INTERP_* */
  float radius;                    /* FeatureGrid rescale x' = (x + radius) / (2 radius); <= 0: none */
  float resolution[MMSB_MAX_LEVELS]; /* res_l = floor(min_res * growth^l), computed on the host     */
} MmsbHashGridDesc;

/* out[n, col] (row stride ld_out) = interp(table, x[n]) * mask[col]   for col in [0, L*F)
 * x: [n, 3] row stride ldx.  table: [L << log2, F].  mask: [L*F] or NULL (= all ones).
 * idx_out: optional int64 [n, L, 8] flat table rows of the 8 corners in the reference's order
 * (hashed_0..hashed_7, encodings.py:274-281) for bit-exact index parity; NULL to skip. */
int mmsb_hashgrid_fwd(const MmsbHashGridDesc* desc, const float* x, int64_t ldx, const float* table,
                      const float* mask, float* out, int64_t ld_out, int64_t* idx_out, int64_t n,
                      mmsb_stream_t stream);

/* dtable[row] += dout[n, lF+f] * mask * w_corner   (fp32 atomics; caller zero-fills dtable)
 * dx[n, :] = d(out)/d(x) contracted with dout (caller-owned, overwritten), NULL to skip.
 * dtable may be NULL to skip the scatter (input-gradient only). */
int mmsb_hashgrid_bwd(const MmsbHashGridDesc* desc, const float* x, int64_t ldx, const float* table,
                      const float* mask, const float* dout, int64_t ld_dout, float* dtable, float* dx,
                      int64_t lddx, int64_t n, mmsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A8  NeRF sinusoidal encoding.  ref: field_components/encodings.py:161-182
 * out[n] = [x (if include_input), sin(x_d * f_k) (d major, k minor), sin(x_d * f_k + pi/2)]
 * freqs_host: [num_freqs] host floats (2 ** linspace(min,max,K), computed by the caller).
 * ---------------------------------------------------------------------------------------------- */
#define MMSB_MAX_FREQS 16
int mmsb_nerf_encoding_fwd(const float* x, int64_t ldx, int32_t in_dim, const float* freqs_host,
                           int32_t num_freqs, int32_t include_input, float* out, int64_t ld_out,
                           int64_t n, mmsb_stream_t stream);
/* dx[n, d] (+)= sum_k f_k * (cos(x f_k) dout_sin + cos(x f_k + pi/2) dout_cos) (+ dout_x).
 * accumulate != 0 adds into dx instead of overwriting. */
int mmsb_nerf_encoding_bwd(const float* x, int64_t ldx, int32_t in_dim, const float* freqs_host,
                           int32_t num_freqs, int32_t include_input, const float* dout,
                           int64_t ld_dout, float* dx, int64_t lddx, int32_t accumulate, int64_t n,
                           mmsb_stream_t stream);

/* A15  real spherical harmonics, levels 1..5 (1,4,9,16,25 outputs).
 * ref: utils/math.py:21-82 (components_from_spherical_harmonics; the pinned in-tree definition —
 * the tcnn SphericalHarmonics op of encodings.py:377-392 is not in the tree: parity unpinned). */
int mmsb_sh_encoding_fwd(const float* dirs, int64_t ldx, int32_t levels, float* out, int64_t ld_out,
                         int64_t n, mmsb_stream_t stream);
int mmsb_sh_encoding_bwd(const float* dirs, int64_t ldx, int32_t levels, const float* dout,
                         int64_t ld_dout, float* ddirs, int64_t lddx, int64_t n, mmsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A11  dense layers of the field MLPs.  ref: field_components/mlp.py:152-171
 * y = act(x W^T + b); W is the effective [out, in] row-major weight (weight-norm already applied
 * by the caller, mlp.py:206-209), b: [out] or NULL.
 * ---------------------------------------------------------------------------------------------- */
#define MMSB_ACT_NONE 0
#define MMSB_ACT_RELU 1
#define MMSB_ACT_SOFTPLUS 2 /* nn.Softplus(beta=act_param, threshold=20) */
#define MMSB_ACT_SIGMOID 3

/* Layers with out_dim <= 16 take a row-streaming path; wider ones a 128x128 register-tiled GEMM. */
int mmsb_linear_fwd(const float* x, int64_t ldx, const float* w, const float* b, float* y, int64_t ldy,
                    int64_t n, int32_t in_dim, int32_t out_dim, int32_t act, float act_param,
                    mmsb_stream_t stream);
/* dz = dy * act'(y)  (in place allowed: dz == dy).  y is the layer OUTPUT (post-activation). */
int mmsb_act_bwd(const float* dy, int64_t lddy, const float* y, int64_t ldy, float* dz, int64_t lddz,
                 int64_t n, int32_t dim, int32_t act, float act_param, mmsb_stream_t stream);
/* dst[r, 0:width] = src[r, 0:width], both row-strided (lds, ldd in floats): places an already materialised tensor
 * (e.g. the geometry features, surface_model.py:124-127 -> radiance_field.py:90-101 `torch.cat`) into its column range
 * of an assembled MLP input row. */
int mmsb_copy_rows(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t n, int32_t width,
                   mmsb_stream_t stream);
/* dx = dz W  ([n,out] x [out,in]); if y_prev != NULL the previous layer's activation derivative is
 * fused into the epilogue: dx *= act_prev'(y_prev) (y_prev: [n, in] with stride ld_yprev). */
int mmsb_linear_bwd_data(const float* dz, int64_t lddz, const float* w, float* dx, int64_t lddx,
                         const float* y_prev, int64_t ld_yprev, int32_t act_prev, float act_prev_param,
                         int64_t n, int32_t in_dim, int32_t out_dim, mmsb_stream_t stream);
/* dw[out,in] += dz^T x,  db[out] += sum_n dz  (split over n with fp32 atomics: the CALLER zero-fills
 * dw/db, or leaves the running sum of earlier calls there to accumulate).  dw or db may be NULL. */
int mmsb_linear_bwd_weight(const float* dz, int64_t lddz, const float* x, int64_t ldx, float* dw,
                           float* db, int64_t n, int32_t in_dim, int32_t out_dim, mmsb_stream_t stream);

/* A11 on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM; csrc/mlp_tc.cu).
 * precision: 3 = "3xTF32" (hi/lo split, three TF32 MMAs per product: fp32-accurate, the parity mode),
 *            2 = 2-term fp16 split (hi/lo fp16 operands under a per-tensor power-of-two scale, three kind::f16 MMAs per
 *                product at twice the TF32 rate: fp32-accurate like 3xTF32).  Needs max |operand| in device memory
 *                (`*_amax`: mmsb_amax, or the `y_amax` a producing tensor-core kernel wrote) and applies to the
 *                CTA-pair shapes only (16-byte aligned operand rows, output width a multiple of 256, >= 18944 rows);
 *                other shapes return MMSB_E_INVALID_ARGUMENT and are run with precision 3 by the caller,
 *            1 = single-pass TF32 (the reference's GPU runs use fp16 autocast, mlp.py:152-171 under
 *                torch.autocast; 1e-2 band).
 * Weights are consumed pre-split and pre-swizzled ("packed"): pack once per optimiser step and layer, reuse
 * for every evaluation of the step.  mmsb_linear_packed_size returns the number of floats of the packed
 * operand for a logical [n_dim x k_dim] B matrix (forward: n_dim = out, k_dim = in; dgrad: n_dim = in,
 * k_dim = out), or -1 on bad arguments. */
int64_t mmsb_linear_packed_size(int32_t n_dim, int32_t k_dim, int32_t precision);
/* w: effective [out_dim, in_dim] weight, row stride ldw.  transpose = 0 packs B = W (forward),
 * transpose = 1 packs B = W^T (dgrad). */
int mmsb_linear_pack_weight(const float* w, int64_t ldw, int32_t out_dim, int32_t in_dim, int32_t transpose,
                            int32_t precision, float* packed, mmsb_stream_t stream);
/* amax[0] = max(amax[0], max |x|) over a row-strided [n, cols] matrix (the caller zero-fills amax once). */
int mmsb_amax(const float* x, int64_t ldx, int64_t n, int32_t cols, float* amax, mmsb_stream_t stream);
/* y = act(x W^T + b) with packed_w = pack(W, transpose = 0).  x_amax: device scalar >= max |x| (precision 2 only, else
 * NULL); y_amax (optional, any precision): device scalar updated with max(y_amax, max |y|) — zero-filled by the caller. */
int mmsb_linear_fwd_tc(const float* x, int64_t ldx, const float* packed_w, const float* b, float* y, int64_t ldy,
                       int64_t n, int32_t in_dim, int32_t out_dim, int32_t act, float act_param, int32_t precision,
                       const float* x_amax, float* y_amax, mmsb_stream_t stream);
/* dx = dz W (* act_prev'(y_prev) when y_prev != NULL) with packed_wt = pack(W, transpose = 1);
 * accumulate != 0: dx += ... (a second gradient path into the same rows, e.g. the geometry-feature path of the SDF
 * network's centre rows on top of the sdf-head path that covers every row). */
int mmsb_linear_bwd_data_tc(const float* dz, int64_t lddz, const float* packed_wt, float* dx, int64_t lddx,
                            const float* y_prev, int64_t ld_yprev, int32_t act_prev, float act_prev_param,
                            int64_t n, int32_t in_dim, int32_t out_dim, int32_t precision, int32_t accumulate,
                            mmsb_stream_t stream);
/* dw[out,in] += dz^T x, db[out] += sum_n dz (db may be NULL); split over n, fp32 reductions in the L2. */
int mmsb_linear_bwd_weight_tc(const float* dz, int64_t lddz, const float* x, int64_t ldx, float* dw, float* db,
                              int64_t n, int32_t in_dim, int32_t out_dim, int32_t precision, mmsb_stream_t stream);

/* The same three products for a layer that is followed by a ONE-output head (the sdf column of the last SDF layer,
 * surface_field.py:70-72 / surface_model.py:143-146: the tap and sampler evaluations keep only output 0), fused so that
 * the head never runs as a layer of its own:
 *   forward : y = act(x W^T + b) (y == NULL: not stored), head_out[row] += y[row, :] . head_w (+ head_b[0]);
 *             head_out must be zeroed by the caller (two commutative additions per row: order-independent);
 *   dgrad   : the operand dz = head_d[row] * head_w[col] * act'(y[row, col]) is generated from the stored y;
 *   wgrad   : likewise, and dhead_w[col] += sum_row head_d[row] * y[row, col] (the head's weight gradient);
 *   rank1   : plain dgrad of the layer ABOVE a head-carrying layer, whose input gradient also receives the head's
 *             rank-1 term: dx = (dz W + head_d[row] * head_w[col]) * act_prev'(y_prev). */
int mmsb_linear_fwd_head_tc(const float* x, int64_t ldx, const float* packed_w, const float* b, float* y, int64_t ldy,
                            int64_t n, int32_t in_dim, int32_t out_dim, int32_t act, float act_param, int32_t precision,
                            const float* head_w, const float* head_b, float* head_out, const float* x_amax, float* y_amax,
                            mmsb_stream_t stream);
int mmsb_linear_bwd_data_head_tc(const float* y, int64_t ldy, int32_t act, float act_param, const float* head_d,
                                 const float* head_w, const float* packed_wt, float* dx, int64_t lddx, const float* y_prev,
                                 int64_t ld_yprev, int32_t act_prev, float act_prev_param, int64_t n, int32_t in_dim,
                                 int32_t out_dim, int32_t precision, mmsb_stream_t stream);
int mmsb_linear_bwd_weight_head_tc(const float* y, int64_t ldy, int32_t act, float act_param, const float* head_d,
                                   const float* head_w, const float* x, int64_t ldx, float* dw, float* db, float* dhead_w,
                                   int64_t n, int32_t in_dim, int32_t out_dim, int32_t precision, mmsb_stream_t stream);
int mmsb_linear_bwd_data_rank1_tc(const float* dz, int64_t lddz, const float* packed_wt, float* dx, int64_t lddx,
                                  const float* y_prev, int64_t ld_yprev, int32_t act_prev, float act_prev_param, int64_t n,
                                  int32_t in_dim, int32_t out_dim, int32_t precision, const float* head_d,
                                  const float* head_w, mmsb_stream_t stream);

/* A12, the whole forward of the SDF network (surface_field.py:99-116: x -> two hidden layers -> output 0 of the last
 * layer; mlp.py:152-171) as ONE persistent kernel on CTA pairs: the hidden activations stay on chip between the layers
 * (layer 0's epilogue writes fp16 hi / lo operand chunks of h0 into shared memory, layer 1's MMAs consume them).
 *   x [n, in_dim] (16-byte aligned rows, 64 < in_dim <= 80), hidden = 256, act = ReLU or Softplus(beta = act_param);
 *   w0 [hidden, in_dim] contiguous (its row norms bound the scale of h0), packed_w0 / packed_w1 = mmsb_linear_pack_weight
 *   (precision 2) of W0 [hidden, in_dim] / W1 [hidden, hidden]; head_w [hidden] / head_b [1] = row 0 of the last layer;
 *   products = 3: 2-term fp16 split, three MMAs per product (fp32-accurate, per-row power-of-two scales);
 *   products = 1: one fp16 MMA per product (the fast mode; 1e-2 band);
 *   sdf [n] is written (not accumulated).  h0 / h1 (NULL: not stored) receive the hidden activations for the backward
 *   kernels; h1_group = g > 1 stores only the rows r with r % g == 0 at h1[r / g] (the centre rows of the grouped layout,
 *   whose geometry features the caller evaluates, surface_model.py:143-146). */
int mmsb_sdf_net_fwd_fused(const float* x, int64_t ldx, int64_t n, int32_t in_dim, int32_t hidden, const float* w0,
                           const float* packed_w0, const float* b0, const float* packed_w1, const float* b1,
                           const float* head_w, const float* head_b, int32_t act, float act_param, int32_t products,
                           float* h0, int64_t ldh0, float* h1, int64_t ldh1, int32_t h1_group, float* sdf,
                           mmsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A17  polarization head: Stokes vector of a sample -> intensities behind the 0 / 45 / 90 / 135 degree polarizers.
 * ref: field_components/field_heads.py:90-106, model_components/polarizer.py:39-101.
 * stokes [n,3] (row stride ld_stokes) = raw head output (leaky_relu on S0 applied here), directions / up_directions
 * [n,3] contiguous, out [n,4].  bwd: d_directions / d_up_directions may be NULL (poses not optimised).
 * ---------------------------------------------------------------------------------------------- */
int mmsb_polarization_fwd(const float* stokes, int64_t ld_stokes, const float* directions, const float* up_directions,
                          float* out, int64_t n, mmsb_stream_t stream);
int mmsb_polarization_bwd(const float* stokes, int64_t ld_stokes, const float* directions, const float* up_directions,
                          const float* d_out, float* d_stokes, float* d_directions, float* d_up_directions, int64_t n,
                          mmsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A1/A2  ray generation.  ref: cameras/camera_optimizers.py:86-119, cameras/lie_groups.py:28-63,
 * model_components/ray_generators.py:54-81, cameras/cameras.py:460-703,
 * cameras/camera_utils.py:279-383, utils/poses.py:53-67
 * coords: int32 [n,3] = (camera, y, x).  c2w: [n_cam,3,4].  intr: [n_cam,4] = fx,fy,cx,cy.
 * dist: [n_cam,6] = k1,k2,k3,k4,p1,p2 or NULL.  pose_adjust: [n_pose,6] = t(3), so3(3) or NULL
 * (mode "off"); n_pose is 1 (shared_optimization) or n_cam.
 * Outputs: origins/directions/up [n,3], pixel_area/dir_norm [n,1].
 * ---------------------------------------------------------------------------------------------- */
int mmsb_raygen_fwd(const int32_t* coords, const float* c2w, const float* intr, const float* dist,
                    const float* pose_adjust, int32_t n_pose, int32_t n_cam, float pixel_offset,
                    float* origins, float* directions, float* up, float* pixel_area, float* dir_norm,
                    int64_t n, mmsb_stream_t stream);
/* d_pose[n_pose,6] += gradients of (origins, directions, up) w.r.t. pose_adjust (atomics; caller
 * zero-fills).  pixel_area / dir_norm gradients are not propagated (never used by a loss). */
int mmsb_raygen_bwd(const int32_t* coords, const float* c2w, const float* intr, const float* dist,
                    const float* pose_adjust, int32_t n_pose, int32_t n_cam, float pixel_offset,
                    const float* d_origins, const float* d_directions, const float* d_up,
                    float* d_pose, int64_t n, mmsb_stream_t stream);

/* A3  ray / sphere collider.  ref: model_components/scene_colliders.py:60-80,96-113
 * near/far/mask for the foreground; if bg_near/bg_far != NULL also the background interval
 * (near := far where hit, far += 3).  mask: uint8 [n]. */
int mmsb_sphere_collide(const float* origins, const float* directions, float radius, float* nears,
                        float* fars, uint8_t* mask, float* bg_nears, float* bg_fars, int64_t n,
                        mmsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A4  spaced (uniform / linear-disparity) sampler.  ref: model_components/ray_samplers.py:183-296
 * lin: [num_samples+1] = torch.linspace(0,1,num_samples+1) (device).  t_rand: NULL (eval),
 * [n,1] (single_jitter, ld_rand = 1... stride 1 per ray) or [n,num_samples+1].
 * Outputs: spacing bins [n, num_samples+1] and euclidean bins [n, num_samples+1].
 * ---------------------------------------------------------------------------------------------- */
#define MMSB_SPACING_UNIFORM 0
#define MMSB_SPACING_DISPARITY 1
int mmsb_spaced_bins(const float* nears, const float* fars, const float* lin, const float* t_rand,
                     int32_t rand_per_ray, int32_t num_samples, int32_t spacing, float* spacing_bins,
                     float* euclid_bins, int64_t n, mmsb_stream_t stream);

/* A5/A6/A7  one NeuS up-sampling round, fused: fixed-inv_s alphas -> weights -> pdf/cdf ->
 * searchsorted(side=right) -> inverse-cdf lerp -> sorted merge.
 * ref: model_components/ray_samplers.py:516-551 (alphas), cameras/rays.py:201-217 (weights),
 *      ray_samplers.py:316-422 (PDFSampler), ray_samplers.py:38-68 (merge_ray_samples)
 * bins: [n, m+1] spacing bins (sorted), sdf: [n, m] at the bin starts, u: [n, k+1] the stratified
 * samples (host builds linspace + rand/(k+1) exactly like the reference), nears/fars: [n].
 * Outputs: cdf_ws [n, m+1] (required workspace; holds the cdf on return), inds int64 [n, k+1]
 * (optional; the searchsorted result),
 * new_bins [n, k+1] (spacing), merged_bins [n, m+k+1], merged_index int64 [n, m+k] = the
 * `sorted_index` of torch.sort(cat(starts_old, starts_new)) with old-before-new on ties. */
int mmsb_neus_upsample(const float* bins, const float* sdf, const float* u, const float* nears,
                       const float* fars, float inv_s, float histogram_padding, float eps, int32_t m,
                       int32_t k, float* cdf_ws, int64_t* inds_out, float* new_bins,
                       float* merged_bins, int64_t* merged_index, int64_t n, mmsb_stream_t stream);

/* A6 on its own: inverse-cdf lerp of stratified samples (bit-exact parity surface).
 * ref: ray_samplers.py:394-403.  cdf/bins [n, num_edges], u [n, q] -> inds int64 [n, q] (optional),
 * new_bins [n, q]. */
int mmsb_pdf_inverse(const float* cdf, const float* bins, const float* u, int32_t num_edges, int32_t q,
                     int64_t* inds, float* new_bins, int64_t n, mmsb_stream_t stream);

/* A5 merge of per-sample values after an up-sampling round: out = gather(cat(a, b), index).
 * ref: ray_samplers.py:486-488.  a [n,m], b [n,k], index int64 [n,m+k] -> out [n,m+k] */
int mmsb_merge_rows(const float* a, int32_t m, const float* b, int32_t k, const int64_t* index,
                    float* out, int64_t n, mmsb_stream_t stream);

/* searchsorted(side="right") on its own (bit-exact index parity test surface).
 * ref: ray_samplers.py:394.  cdf [n, m], u [n, q] -> inds int64 [n, q] */
int mmsb_searchsorted_right(const float* cdf, const float* u, int64_t* inds, int32_t m, int32_t q,
                            int64_t n, mmsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A14  NeuS alphas + transmittance weights, one warp per ray (warp scan), forward and backward.
 * ref: model_components/volume_rendering.py:177-213, field_components/single_variance.py:34-36
 * sdf [n,s], grad [n,s,3], dirs [n,3], deltas [n,s]; inv_s: device scalar (already
 * exp(10 s).clip(1e-6,1e6)); anneal: cos_anneal_ratio.  weights [n,s].
 * mask: uint8 [n] or NULL — rays outside the sphere get zero weights / zero gradients, which is what
 * the reference's boolean compaction + masked scatter amounts to (models/base_model.py:88-93,
 * renderers.py:105-135) without a host synchronisation.
 * ---------------------------------------------------------------------------------------------- */
int mmsb_neus_weights_fwd(const float* sdf, const float* grad, const float* dirs, const float* deltas,
                          const float* inv_s, const uint8_t* mask, float anneal, float* weights,
                          int32_t s, int64_t n, mmsb_stream_t stream);
/* Gradients w.r.t. sdf, grad, dirs (optional), deltas (optional), inv_s (atomic into d_inv_s[1],
 * caller zero-fills). */
int mmsb_neus_weights_bwd(const float* sdf, const float* grad, const float* dirs, const float* deltas,
                          const float* inv_s, const uint8_t* mask, float anneal, const float* d_weights,
                          float* d_sdf, float* d_grad, float* d_dirs, float* d_deltas, float* d_inv_s,
                          int32_t s, int64_t n, mmsb_stream_t stream);

/* A18 (weights part)  density -> alpha -> weights for the background field.
 * ref: cameras/rays.py:138-151,201-217.  density/deltas [n,s] -> weights [n,s]; backward to density. */
int mmsb_density_weights_fwd(const float* density, const float* deltas, float* weights, int32_t s,
                             int64_t n, mmsb_stream_t stream);
int mmsb_density_weights_bwd(const float* density, const float* deltas, const float* d_weights,
                             float* d_density, float* d_deltas, int32_t s, int64_t n,
                             mmsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A19  per-modality compositing (RadianceRenderer + accumulation + depth + normals), warp per ray.
 * ref: model_components/renderers.py:76-136,149-243
 * weights [n,s]; values [n,s,c]; background [n,c] or NULL (black);
 * out_color [n,c] = sum_s w v + bg (1 - sum_s w).  Optional: normals [n,s,3] -> out_normals [n,3];
 * starts/ends [n,s] -> out_depth [n] (UNCLIPPED sum w (start+end)/2; the caller applies the global
 * min/max clip of renderers.py:215); out_acc [n].
 * ---------------------------------------------------------------------------------------------- */
int mmsb_composite_fwd(const float* weights, const float* values, const float* background, int32_t c,
                       const float* normals, const float* starts, const float* ends, float* out_color,
                       float* out_normals, float* out_depth, float* out_acc, int32_t s, int64_t n,
                       mmsb_stream_t stream);
/* d_weights [n,s] (overwritten), d_values [n,s,c], d_background [n,c] (optional) from d_color [n,c]
 * (+ optional d_acc [n], d_normals[n,3] with normals, d_depth[n] with starts/ends; NULL = zero). */
int mmsb_composite_bwd(const float* weights, const float* values, const float* background, int32_t c,
                       const float* d_color, const float* normals, const float* d_out_normals,
                       const float* starts, const float* ends, const float* d_out_depth,
                       const float* d_out_acc, float* d_weights, float* d_values, float* d_background,
                       float* d_normals, int32_t s, int64_t n, mmsb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A21/A22  mosaick channel select fused with the L1 / skip-saturation loss.
 * ref: pipelines/raw_pipeline.py:112-122, data/datasets.py:229-250,
 *      model_components/losses.py:97-105,158-164
 * coords int32 [n,3] (cam,y,x); pattern int32 [ph*pw] row-major (band = pattern[y%ph][x%pw]);
 * rendered [n,c]; target [n] (raw: one value per pixel).
 * pattern == NULL: demosaicked frames (presets `grid`, `mlp`): all c channels supervised, target [n,c].
 * Outputs: band int64 [n] (optional, index parity), selected [n] (optional; the gathered channel),
 * loss_sum[1] += sum |pred - target| (caller zero-fills and divides by the element count).
 * Skip-saturation (sat_index != NULL): where target > sat_threshold the prediction is replaced by
 * target[*sat_index] (the first saturated target in flattened order, see mmsb_first_saturated) and
 * gets no gradient; *sat_index == number of targets means "none saturated".
 * ---------------------------------------------------------------------------------------------- */
int mmsb_mosaick_l1_fwd(const int32_t* coords, const int32_t* pattern, int32_t ph, int32_t pw,
                        const float* rendered, int32_t c, const float* target, float sat_threshold,
                        const int64_t* sat_index, int64_t* band_out, float* selected, float* loss_sum,
                        int64_t n, mmsb_stream_t stream);
/* d_rendered [n,c] (overwritten; zero except the selected band) = scale * sign(sel - target),
 * scale = *d_loss (device scalar) * inv_count. */
int mmsb_mosaick_l1_bwd(const int32_t* coords, const int32_t* pattern, int32_t ph, int32_t pw,
                        const float* rendered, int32_t c, const float* target, float sat_threshold,
                        const int64_t* sat_index, const float* d_loss, float inv_count,
                        float* d_rendered, int64_t n, mmsb_stream_t stream);
/* sat_index[0] = index of the first target > sat_threshold (flattened order), or n when none.
 * ref: losses.py:160-163 */
int mmsb_first_saturated(const float* target, float sat_threshold, int64_t* sat_index, int64_t n,
                         mmsb_stream_t stream);

/* A22 geometry losses.  ref: losses.py:113-119 (eikonal), 143-150 (curvature)
 * gradients/hessians [n,3] (n = rays*s samples, hessians optional); ray_mask uint8 [n/s] or NULL.
 * sums[3] += { sum (|g|-1)^2, sum |hxx+hyy+hzz|, number of unmasked samples } (caller zero-fills;
 * several modalities accumulate into one sums buffer = the reference's torch.cat, losses.py:245-248).
 * Backward: d_gradients = *d_eik * 2(|g|-1) g/|g| / sums[2]; d_hessians = *d_curv * sign(lap) / sums[2]. */
int mmsb_geometry_loss_fwd(const float* gradients, const float* hessians, const uint8_t* ray_mask,
                           int32_t s, float* sums, int64_t n, mmsb_stream_t stream);
int mmsb_geometry_loss_bwd(const float* gradients, const float* hessians, const uint8_t* ray_mask,
                           int32_t s, const float* sums, const float* d_eik, const float* d_curv,
                           float* d_gradients, float* d_hessians, int64_t n, mmsb_stream_t stream);

/* A13  numerical SDF gradients / Hessian diagonal / normals from the 4 tetrahedron taps.
 * ref: model_components/surface_model.py:137-153,98
 * sdf_c [n] centre, sdf_t [4,n] taps (k1..k4 order).  With delta = numerical_gradients_delta/sqrt(3)
 * evaluated in double by the caller like the reference's Python scalars: four_delta = float(4*delta),
 * delta_sq = float(delta**2).  gradients [n,3], hessians [n,3] (optional), normals [n,3] (optional). */
int mmsb_sdf_taps_fwd(const float* sdf_c, const float* sdf_t, float four_delta, float delta_sq,
                      float* gradients, float* hessians, float* normals, int64_t n, mmsb_stream_t stream);
int mmsb_sdf_taps_bwd(const float* sdf_t, float four_delta, float delta_sq, const float* d_gradients,
                      const float* d_hessians, const float* d_normals, float* d_sdf_c, float* d_sdf_t,
                      int64_t n, mmsb_stream_t stream);

/* A23 (next row 1)  fused AdamW step over a flat fp32 parameter.
 * ref: engine/optimizers.py:96-105, configs/method_configs.py:260-269 (torch.optim.AdamW semantics)
 * grad_scale: device scalar multiplied into the gradient (global-norm clip coefficient) or NULL. */
int mmsb_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                    const float* grad_scale, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int32_t step, int64_t n, mmsb_stream_t stream);
/* The same update, replayable from a CUDA graph: the step-dependent scalars live in device memory,
 * hyper = {lr, 1 - beta1^t, sqrt(1 - beta2^t), prescale}; the gradient is read as grad * prescale (1 / world size
 * after a summing all-reduce = DDP's mean, 1 otherwise); grad_sumsq (or NULL) + max_norm give the clip coefficient
 * min(1, max_norm / (sqrt(sumsq) * prescale + 1e-6)) of clip_grad_norm_ (pipelines/base_pipeline.py:232-248). */
int mmsb_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                        const float* grad_sumsq, float max_norm, const float* hyper, float beta1, float beta2,
                        float eps, float weight_decay, int64_t n, mmsb_stream_t stream);
/* sumsq[0] += sum grad^2 (for clip_grad_norm_, pipelines/base_pipeline.py:232-248). */
int mmsb_sumsq(const float* x, float* sumsq, int64_t n, mmsb_stream_t stream);

/* SURVEY 8(f) row 2  on-device pixel sampling + target gather.
 * ref: cameras/pixel_samplers.py:71-89 (uniform (camera, y, x) per ray), data/dataloaders.py:164-167 (target gather).
 * coords int32 [n,3]; frames [n_cam, height, width, channels] fp32 device-resident (or NULL: indices only) ->
 * targets [n, channels].  Counter-based (Philox-4x32-10): the draw of ray i depends only on (seed, step, stream_id, i). */
int mmsb_sample_pixels(uint64_t seed, int32_t step, int32_t stream_id, int32_t n_cam, int32_t height, int32_t width,
                       const float* frames, int32_t channels, int32_t* coords, float* targets, int64_t n,
                       mmsb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMS_B200_H */
