/*
 * pv2_b200.h — C ABI of libpv2_b200.so: the B200 (sm_100a) implementation of the PonderV2
 * pretraining hot path.  This is the drop-in boundary underneath
 *   B1  spconv.pytorch            (reference call sites: ponder/models/sparse_unet/spconv_unet_v1m1_base.py:11-278)
 *   B2  smooth_sampler._C         (reference: libs/smooth-sampler/smooth_sampler/csrc/smooth_sampler.cpp:36-97)
 *   B3  render_utils (NeuSModel)  (reference: ponder/models/ponder/render_utils/**)
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller (torch) owns
 *     all buffers including workspaces; nothing in here allocates, frees or synchronises.
 *   - `stream` is a cudaStream_t passed as void*; work is stream-ordered.  Calls on different streams / devices may
 *     overlap; the only process-wide state is (a) the option table below, (b) the launch counter, (c) per-(kernel, device)
 *     "opt-in shared memory attribute set" flags, and (d) a few PV2_* development environment switches read once.
 *   - return value: 0 on success, a negative PV2_E* code for argument errors, a positive
 *     cudaError_t for launch errors.  pv2_error_string() maps either to text.
 *   - feature matrices are row-major [rows, channels] ("channels-last").
 *   - dtype codes: PV2_F32 = 0, PV2_BF16 = 1, PV2_F64 = 2 (sampler KAT only).
 */
#ifndef PV2_B200_H_
#define PV2_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PV2_F32 0
#define PV2_BF16 1
#define PV2_F64 2

#define PV2_EINVAL (-1)   /* bad argument (null pointer, negative size, unsupported shape) */
#define PV2_EWORKSPACE (-2) /* workspace too small */
#define PV2_EUNSUPPORTED (-3)

int pv2_version(void);
const char* pv2_error_string(int code);
/* Total number of CUDA kernels this library has launched in the calling process (statistics only). */
int64_t pv2_launch_count(void);
/* Number of SMs the library sized its persistent grids for (148 on B200); 0 if no device. */
int pv2_sm_count(void);
/* Kernel-selection switches for A/B measurements and tests (defaults from the environment, read once: PV2_GG_TMA,
 * PV2_GG_BX3, PV2_WGRAD_MN, PV2_LINEAR_BX3, PV2_GG_KSPLIT_MAX, PV2_GG_GROUPS, PV2_GG_BX3_SPLIT).  "gg_tma": bf16 gather
 * through TMA gather4 (-1 auto by size, 0 off, 1 on); "gg_bx3": fp32 gather-GEMM as bf16x3 (0 / 1); "wgrad_mn": MN-major
 * bf16 weight-gradient kernel (0 / 1); "linear_bx3": render-MLP linears on the bf16x3 kernel (0 / 1); "gg_ksplit_max": cap
 * of the split-K factor on the small levels (0 = none); "gg_bx3_split": small levels on the split-K bf16x3 kernel
 * (default 0, measured slower); "gg_groups": reserved.  Results never depend on them beyond the stated tolerances.
 * set: 0 / PV2_EINVAL; get: the value, -1 for an unknown name. */
int pv2_set_option(const char* name, int value);
int pv2_get_option(const char* name);

/* ------------------------------------------------------------------------------------------
 * Rulebook (neighbour-map) build.  Replaces spconv's indice-pair generation that the first
 * SubMConv3d / SparseConv3d of every indice_key triggers
 * (spconv_unet_v1m1_base.py:47-66,111-119 keys stem/subm0..4; :135-142 keys spconv1..4).
 * coords: [n,4] int32 rows (batch, c0, c1, c2), all >= 0, c_a < spatial_shape[a].
 * ------------------------------------------------------------------------------------------ */

/* bytes of workspace for the coordinate hash table used by both rulebook builders */
size_t pv2_rulebook_workspace_bytes(int64_t n);

/* Submanifold map.  nbr: [ksize^3, n] int32, nbr[k][j] = row i whose coord equals
 * coord[j] + (k_a - ksize/2) per axis, k = (k0*ksize + k1)*ksize + k2, or -1.
 * Duplicate coords resolve to the smallest row index.  pair_count (optional, may be NULL):
 * device int64[1] receiving the number of non-negative entries. */
int pv2_rulebook_subm(const int32_t* coords, int64_t n, const int32_t* spatial_shape_host,
                      int ksize, int32_t* nbr, int64_t* pair_count,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Strided (kernel 2, stride 2, pad 0) map, stage 1.  Output voxels are the distinct
 * (batch, c>>1) rows, numbered in order of the first input row that maps to them.
 *   out_coords [n,4] (capacity n; first *n_out rows valid), in2out [n], koff [n] with
 *   koff = ((c0&1)*2 + (c1&1))*2 + (c2&1); n_out: device int32[1]. */
int pv2_rulebook_down(const int32_t* coords, int64_t n, const int32_t* spatial_shape_host,
                      int32_t* out_coords, int32_t* in2out, int32_t* koff, int32_t* n_out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Stage 2 (after the caller read n_out): gather maps for both directions.
 *   nbr_down [8, n_out]: nbr_down[k][j] = fine row i with in2out[i]==j && koff[i]==k, else -1
 *                        (SparseConv3d fwd, SparseInverseConv3d dgrad)
 *   nbr_up   [8, n]    : nbr_up[k][i] = in2out[i] if koff[i]==k else -1
 *                        (SparseInverseConv3d fwd, SparseConv3d dgrad) */
int pv2_rulebook_down_maps(const int32_t* in2out, const int32_t* koff, int64_t n, int64_t n_out,
                           int32_t* nbr_down, int32_t* nbr_up, void* stream);

/* Tile order for the sparse-conv kernels: order[pos] = row, rows sorted (stably) by their neighbour-presence mask
 * mask[j] = OR_k (nbr[k][j] >= 0) << k, so that a 128-row tile touches few distinct offsets and the (tile, offset)
 * pairs without any neighbour can be skipped (what spconv's implicit-GEMM mask sort does).  For 32 < kvol <= 128 the
 * key is the mask folded modulo 32; kvol > 128 returns PV2_EUNSUPPORTED.  Convolution results never depend on the order.
 * Optional outputs (may be NULL): nbr_sorted [kvol, n] = the map in tile order, nbr_sorted[k][pos] = nbr[k][order[pos]]
 * (what the conv kernels take together with `order`); blk_active [kvol, ceil(n/32)] = 1 iff any of the 32 consecutive
 * tile-order rows of a block has a neighbour at offset k (the weight-gradient kernel skips the other blocks). */
size_t pv2_rulebook_row_order_workspace_bytes(int64_t n);
int pv2_rulebook_row_order(const int32_t* nbr, int64_t n, int kvol, int32_t* order, int32_t* nbr_sorted,
                           uint8_t* blk_active, void* workspace, size_t workspace_bytes, void* stream);

/* batch ids from cumulative offsets (ponder/models/utils.py:11-26 offset2batch) fused with the
 * [n,4] int32 (batch, c0, c1, c2) assembly of spconv_unet_v1m1_base.py:247-256.
 * grid_coord: [n,3] int64, offset: [b] int64 cumulative. */
int pv2_make_indices(const int64_t* grid_coord, const int64_t* offset, int64_t n, int batch,
                     int32_t* indices, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sparse convolution arithmetic (gather -> GEMM -> output-stationary accumulate).
 * Replaces spconv's SubMConv3d / SparseConv3d / SparseInverseConv3d forward, dgrad and wgrad.
 *   y[j, :] = bias + sum_k W_k * x[nbr[k][j], :]          (rows with nbr == -1 contribute 0)
 * Weight element (co, k, ci) is read at w[co*w_stride_co + k*w_stride_k + ci] so the same
 * entry point serves forward (spconv layout [Cout,K,Cin]: strides K*Cin, Cin) and dgrad
 * (the transposed/k-flipped copy the host shim prepares).
 * ------------------------------------------------------------------------------------------ */
/* row_order (optional, may be NULL): [n_out] permutation from pv2_rulebook_row_order grouping the output rows into
 * 128-row tiles.  When it is given, `nbr` must be the map IN TILE ORDER (pv2_rulebook_row_order's nbr_sorted):
 * nbr[k][pos] is the input row feeding output row row_order[pos]. */
int pv2_spconv_gather_gemm(const void* x, const void* w, int64_t w_stride_co, int64_t w_stride_k,
                           const float* bias, const int32_t* nbr, const int32_t* row_order, void* y,
                           int64_t n_in, int64_t n_out, int cin, int cout, int kvol,
                           int dtype, void* workspace, size_t workspace_bytes, void* stream);
/* Scratch of the gather-GEMM: 0 (operands are gathered raw and split into TF32 halves on chip); kept so callers size
 * workspaces uniformly. */
size_t pv2_spconv_workspace_bytes(int64_t n_in, int cin, int cout, int kvol, int dtype);

/* Weights for the data gradient: out[ci][k][co] = w[co][flip ? K-1-k : k][ci]  (w is [cout, kvol, cin], out
 * [cin, kvol, cout]); with flip = 1 a submanifold conv's dgrad is pv2_spconv_gather_gemm over the forward map. */
int pv2_spconv_dgrad_weights(const void* w, void* out, int cout, int kvol, int cin, int flip, int dtype, void* stream);

/* dw[co, k, ci] (+)= sum_j dy[j, co] * x[nbr[k][j], ci];  dw is float32 [Cout, K, Cin], must be
 * zeroed by the caller (accumulated with atomics across row chunks). */
/* row_order / nbr as for pv2_spconv_gather_gemm; blk_active (optional) from pv2_rulebook_row_order. */
int pv2_spconv_wgrad(const void* x, const void* dy, const int32_t* nbr, const int32_t* row_order,
                     const uint8_t* blk_active, float* dw,
                     int64_t n_in, int64_t n_out, int cin, int cout, int kvol,
                     int dtype, void* workspace, size_t workspace_bytes, void* stream);
/* scratch of the fp32 tensor-core weight-gradient kernel (dy^T in tile order); without it the exact-fp32 SIMT
 * kernel runs */
size_t pv2_wgrad_workspace_bytes(int64_t n_in, int64_t n_out, int cin, int cout);

/* ------------------------------------------------------------------------------------------
 * BatchNorm1d with batch statistics (+ residual add, + ReLU) over [n, c] voxel features, c % 4 == 0:
 * the conv -> bn -> relu / conv -> bn -> (+ residual) -> relu chains of spconv_unet_v1m1_base.py:70-83,111-180.
 *   y = [relu]((x - mean) * invstd * gamma + beta [+ res]);  mean / invstd [c] are outputs (saved for backward);
 *   running_mean / running_var (optional) are updated with torch's semantics (momentum, unbiased variance).
 * Backward: dz = dy * (y > 0 if relu); dgamma = sum dz * xhat; dbeta = sum dz;
 *   dx = gamma * invstd * (dz - dbeta / n - xhat * dgamma / n);  dres (optional) = dz.
 * Deterministic (per-block partial sums combined in double; no atomics).
 * ------------------------------------------------------------------------------------------ */
size_t pv2_bn_workspace_bytes(int64_t n, int c);
int pv2_bn_act_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, float momentum, float eps, int relu, int64_t n, int c, float* y, float* mean,
                   float* invstd, void* workspace, size_t workspace_bytes, void* stream);
int pv2_bn_act_bwd(const float* x, const float* dy, const float* y, const float* gamma, const float* mean,
                   const float* invstd, int relu, int64_t n, int c, float* dx, float* dres, float* dgamma, float* dbeta,
                   void* workspace, size_t workspace_bytes, void* stream);
/* The same two calls with the [n, c] feature matrices (x, res, y / dy, dx, dres) stored in `dtype` (PV2_F32 or
 * PV2_BF16); statistics, gamma / beta and their gradients are always float32.  Replaces nn.BatchNorm1d under the
 * reference's autocast (engines/train.py:183-196): half-precision activations, fp32 statistics. */
int pv2_bn_act_fwd_t(const void* x, const void* res, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, int relu, int64_t n, int c, void* y, float* mean,
                     float* invstd, int dtype, void* workspace, size_t workspace_bytes, void* stream);
/* accumulate != 0: dgamma / dbeta are ADDED to (slices of the caller's flat gradient buffer) instead of overwritten */
int pv2_bn_act_bwd_t(const void* x, const void* dy, const void* y, const float* gamma, const float* mean,
                     const float* invstd, int relu, int64_t n, int c, void* dx, void* dres, float* dgamma, float* dbeta,
                     int accumulate, int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * SDF decoder of the outdoor configuration (hidden 16, feature width 32, L <= 8 linears, softplus beta = 100):
 * `SDFDecoder.forward` (render_utils/decoders.py:6-36) fused with the two vector-Jacobian products u = d sdf / d f
 * [P,32] and v = d sdf / d p [P,3] (direct fc_p path) that `autograd.grad(sdf, points, create_graph=True)`
 * (fields/sdf_field.py:226-238) needs; u / v may be NULL (no-grad coarse pass).  params: packed
 * [Wp 16x3 | bp 16 | Fc_l 16x32 (x L) | bc_l 16 (x L) | W_l 16x16 (x L-1), W_last O x 16 | b_l 16 (x L-1), b_last O]
 * (pv2_sdf_mlp_param_count floats).  bwd: given dL/dsdf, dL/du, dL/dv (any may be NULL) writes fbar = dL/df [P,32] and the
 * per-layer adjoint vectors A, C [L-1][P][16], Z, D [L][P][16] whose contractions over P are the parameter gradients
 * (host: ponderv2_b200/render/mlp.py).  F != 32 or H != 16 -> PV2_EUNSUPPORTED.
 * ------------------------------------------------------------------------------------------ */
int64_t pv2_sdf_mlp_param_count(int L, int O);
int pv2_sdf_mlp_fwd(const float* f, const float* pts, const float* params, int L, int F, int H, int O, float points_factor,
                    int64_t P, float* sdf, float* u, float* v, void* stream);
int pv2_sdf_mlp_bwd(const float* f, const float* pts, const float* params, int L, int F, int H, int O, float points_factor,
                    int64_t P, const float* g_sdf, const float* g_u, const float* g_v, float* fbar, float* A, float* C,
                    float* Z, float* D, void* stream);

/* ------------------------------------------------------------------------------------------
 * Densify: voxel features -> dense channels-last volume, scatter-mean
 * (ponder_indoor_base.py:177-216,332-342; ponder_outdoor_base.py:178-210).
 * cell: [n] int64 flattened cell id in the OUTPUT memory order, or -1 to drop the row.
 * volume: [cells, c] float32 (zero-filled here), count: [cells] int32 (zero-filled here).
 * ------------------------------------------------------------------------------------------ */
int pv2_densify_fwd(const float* feat, const int64_t* cell, int64_t n, int c, int64_t cells,
                    float* volume, int32_t* count, void* stream);
/* dfeat[i,:] = dvolume[cell[i],:] / count[cell[i]]  (0 for cell[i] outside [0, cells), as the forward drops them) */
int pv2_densify_bwd(const float* dvolume, const int64_t* cell, const int32_t* count, int64_t n,
                    int c, int64_t cells, float* dfeat, void* stream);

/* ------------------------------------------------------------------------------------------
 * Trilinear sampler with first and second derivatives (B2).  Same argument meaning as
 * smooth_sampler._C.forward / backward / backward_backward (smooth_sampler.cpp:36-97):
 *   input (N,C,D,H,W) contiguous, grid (N,P,3) with P = Do*Ho*Wo, output (N,C,P).
 *   padding_mode 0 zeros / 1 border / 2 reflection.
 * dtype PV2_F32 or PV2_F64.
 * ------------------------------------------------------------------------------------------ */
int pv2_trilinear_fwd(const void* input, const void* grid, void* output,
                      int64_t N, int64_t C, int64_t D, int64_t H, int64_t W, int64_t P,
                      int padding_mode, int align_corners, int apply_smoothstep,
                      int dtype, void* stream);
/* grad_input may be NULL (input does not require grad); otherwise zero-filled by the caller. */
int pv2_trilinear_bwd(const void* grad_output, const void* input, const void* grid,
                      void* grad_input, void* grad_grid,
                      int64_t N, int64_t C, int64_t D, int64_t H, int64_t W, int64_t P,
                      int padding_mode, int align_corners, int apply_smoothstep,
                      int dtype, void* stream);
/* grad_out_input may be NULL; grad_input and grad_grad_out are zero-filled by the caller. */
int pv2_trilinear_bwd_bwd(const void* grad_out_input, const void* grad_out_grid,
                          const void* input, const void* grid, const void* grad_output,
                          void* grad_input, void* grad_grid, void* grad_grad_out,
                          int64_t N, int64_t C, int64_t D, int64_t H, int64_t W, int64_t P,
                          int padding_mode, int align_corners, int apply_smoothstep,
                          int dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Render MLP building block: dense per-row linear layer on the tensor cores (fp32 storage, 3xTF32).
 * Replaces the nn.Linear chains of ponder/models/ponder/render_utils/decoders.py:6-109.
 *   y[j, 0:cout] = act(x[j, 0:cin] . w^T + bias);  w is [cout, cin] row-major.
 * x: plain [rows, cin] (x_presplit = 0, x_row = cin) or split-precision (hi at x[j*x_row + c], lo at +x_lo_off).
 * y (and y2 when act = 1): row stride y_row; y_split = 1 writes TF32 hi/lo halves (lo at +y_lo_off).
 * act: 0 none; 1 y = softplus(beta=100, threshold=20), y2 = sigmoid(100 v) (the softplus derivative);
 *      backward epilogues, y2 [rows, y2_row] an INPUT s: 2 y = v*100*s*(1-s); 3 y = y + v*s; 4 y = y + v.
 * ------------------------------------------------------------------------------------------ */
size_t pv2_linear_workspace_bytes(int64_t rows, int cin, int cout, int x_presplit);
int pv2_linear(const float* x, int64_t x_row, int64_t x_lo_off, int x_presplit, const float* w, const float* bias,
               float* y, int64_t y_row, int64_t y_lo_off, int y_split, int act, float* y2, int64_t y2_row,
               int64_t y2_lo_off, int64_t rows, int cin, int cout, void* workspace, size_t workspace_bytes,
               void* stream);
/* dw[co, ci] += sum_j dy[j, co] * x[j, ci]; operands plain (lo_off = 0) or split-precision (value = hi + lo). */
int pv2_dense_wgrad(const float* x, int64_t x_row, int64_t x_lo_off, const float* dy, int64_t dy_row,
                    int64_t dy_lo_off, int64_t rows, int cin, int cout, float* dw, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Renderer field kernels on a channels-last volume [Z][Y][X][C] (fields/sdf_field.py:148-257 and the
 * sampler kernels smooth_sampler_kernel.cu:39-619 they drive).  pts are normalised coordinates in [0,1]^3
 * (grid = 2p-1, zeros padding, align_corners=True).
 * ------------------------------------------------------------------------------------------ */
/* trilinear fetch of channels [0,c_use): [0,ca) -> out_a split-precision, [ca,c_use) -> out_b plain */
int pv2_field_sample_fwd(const float* vol, const float* pts, int64_t P, int Z, int Y, int X, int C, int c_use, int ca,
                         float* out_a, int64_t a_row, int64_t a_lo, float* out_b, int64_t b_row, void* stream);
/* grad = d sdf/d p = J^T u;  rgb = sigmoid(Mr [grad|f_r|geo|dir] + cr);  Mr is [3,134] */
int pv2_field_post_fwd(const float* vol, const float* pts, const float* dirs, int samples_per_ray, const float* u,
                       const float* f_r, const float* out_geo, int64_t geo_row, const float* Mr, const float* cr,
                       int64_t P, int Z, int Y, int X, int C, float* grad, float* rgb, void* stream);
/* backward of the above; doutbar is [P, 72] (d sdf | d geo[64] | zeros: rows padded to a multiple of 8 channels) */
int pv2_field_post_bwd(const float* vol, const float* pts, const float* dirs, int samples_per_ray, const float* f_r,
                       const float* out_geo, int64_t geo_row, const float* grad, const float* rgb, const float* Mr,
                       const float* g_rgb, const float* g_grad, const float* g_sdf, int64_t P, int Z, int Y, int X,
                       int C, float* gbar, float* dF, int64_t dF_row, float* doutbar, float* ubar, float* dMr,
                       float* dcr, float* ubar_sum /*[64] accumulated, optional*/,
                       float* dout_sum /*[68] accumulated, optional*/, void* stream);
/* dvol[corner,c] += w dF[p,c] + [c<cs] (dw.gbar) u[p,c]   (u may be NULL) */
int pv2_field_sample_bwd(const float* pts, const float* dF, int64_t dF_row, const float* u, const float* gbar,
                         int64_t P, int Z, int Y, int X, int C, int cs, float* dvol, void* stream);

/* ------------------------------------------------------------------------------------------
 * Per-ray kernels of the NeuS renderer (one warp per ray).  Replace the torch op chains of
 * scene_colliders.py:38-99, ray_samplers.py:55-107,227-463, rays.py:83-153, sdf_field.py:122-146,
 * renderers.py:5-75 and base_surface_model.py:102-211.  R rays, S0 coarse + Si importance samples, S = S0 + Si.
 * ------------------------------------------------------------------------------------------ */
/* AABB collider + stratified spacing bins [R,S0+1] + coarse points [R,S0,3].  noise: [R,noise_cols] uniform numbers
 * (noise_cols = S0+1, or 1 for single_jitter) or NULL (eval: no jitter).  bbox_host: 6 floats (min xyz, max xyz). */
int pv2_ray_setup(const float* origins, const float* dirs, const float* noise, int noise_cols, int64_t R, int S0,
                  const float* bbox_host, float near_plane, float* nears, float* fars, float* bins, float* pts,
                  void* stream);
/* NeuSSampler with one upsample step: fixed-inv_s alphas of the coarse sdf [R,S0] -> weights (init_weights [R,S0],
 * optional) -> PDF resampling (noise [R,noise_cols], noise_cols = Si+1 or 1, NULL = bin centres) -> merge ->
 * starts/deltas [R,S], sample points [R,S,3] (normalised by 1+norm_padding+1e-3 when norm_pts), new_bins [R,Si]
 * (optional), minmax[2] = float bits of the global min / max of starts.  S0 <= 128, Si <= 63. */
int pv2_ray_resample(const float* origins, const float* dirs, const float* nears, const float* fars, const float* bins,
                     const float* sdf, const float* noise, int noise_cols, int64_t R, int S0, int Si, float inv_s,
                     int norm_pts, float norm_padding, float* starts, float* deltas, float* pts_norm, float* init_weights,
                     float* new_bins, int32_t* minmax, void* stream);
/* NeuS alpha -> transmittance -> weights [R,S] -> rgb [R,3] (rgbs/rgb may both be NULL), depth [R] (clipped to the
 * global start range), normal [R,3].  variance: device float[1] (inv_s = clip(exp(10 v), 1e-6, 1e6)).  S <= 256. */
int pv2_ray_composite_fwd(const float* sdf, const float* grad, const float* rgbs, const float* starts,
                          const float* deltas, const float* dirs, const float* variance, const int32_t* minmax,
                          float cos_anneal, int64_t R, int S, int clamp_rgb, float* weights, float* rgb, float* depth,
                          float* normal, void* stream);
/* g_rgb [R,3], g_depth [R], g_normal [R,3], g_weights [R,S] may each be NULL; g_variance [1] is accumulated (zero it). */
int pv2_ray_composite_bwd(const float* sdf, const float* grad, const float* rgbs, const float* starts,
                          const float* deltas, const float* dirs, const float* variance, const int32_t* minmax,
                          float cos_anneal, int64_t R, int S, const float* g_rgb, const float* g_depth,
                          const float* g_normal, const float* g_weights, float* g_sdf, float* g_grad, float* g_rgbs,
                          float* g_variance, void* stream);
/* Loss partial sums (zeroed here): sums[0..4] = numerators of depth-L1, rgb-L1, free-space, sdf, eikonal;
 * sums[5..9] = counts n_valid, 0, n_front, n_sdf_mask, 0; sums[10] = sum of squared rgb error. */
int pv2_ray_loss_fwd(const float* depth_pred, const float* rgb_pred, const float* depth_gt, const float* rgb_gt,
                     const float* sdf, const float* z, const float* grad, int64_t R, int S, float trunc, float* sums,
                     void* stream);
/* gradients of sum_k coef[k] * numerator_k (coef: device float[5]) w.r.t. depth_pred, rgb_pred, sdf, grad */
int pv2_ray_loss_bwd(const float* depth_pred, const float* rgb_pred, const float* depth_gt, const float* rgb_gt,
                     const float* sdf, const float* z, const float* grad, int64_t R, int S, float trunc,
                     const float* coef, float* g_depth, float* g_rgb, float* g_sdf, float* g_grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PV2_B200_H_ */
