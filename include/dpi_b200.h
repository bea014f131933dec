/*
 * dpi_b200.h - C ABI of libdpi_b200.so: the sm_100a implementation of the DPI
 * variational-posterior training step (RealNVP flow run in reverse with log|det J|,
 * measurement operators, data/prior/entropy losses, their backward pass, clip + Adam).
 *
 * The reference (HeSunPU/DPI) is pure Python/PyTorch and has no FFI layer of its own; the
 * Python callables listed below are what this library stands behind. Every entry point
 *   - takes raw device pointers into caller-owned (torch) memory, explicit sizes, a
 *     caller-supplied workspace and a cudaStream_t;
 *   - is asynchronous on that stream: no internal device synchronisation, no host read of
 *     device data (unlike the reference's ActNorm `.item()`, realnvpfc_model.py:33,52);
 *   - is safe to call from any host thread (PyTorch calls backward from its autograd thread);
 *     the device is taken from the handle, not from thread-local state;
 *   - returns 0 on success or a negative code; `dpi_last_error()` (thread-local) explains.
 * Handles own only immutable device tables built at create time.
 *
 * Paths are relative to /root/reference/DPItorch.
 */
#ifndef DPI_B200_H
#define DPI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dpi_stream_t; /* cudaStream_t */

#define DPI_OK 0
#define DPI_ERR_INVALID (-1)
#define DPI_ERR_CUDA (-2)
#define DPI_ERR_UNSUPPORTED (-3)
#define DPI_ERR_WORKSPACE (-4)

const char* dpi_last_error(void);
int dpi_version(void);
/* Number of kernel launches issued by this library in this process (for bench.py's gpu_launches). */
long long dpi_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * RealNVP flow: generative_model/realnvpfc_model.py  RealNVP :562-603, Flow :427-474,
 * AffineCoupling :143-257, ZeroFC :128-141, ActNorm :8-66.
 * All trainable tensors live in ONE flat fp32 buffer (same order as nn.Module.parameters(),
 * each tensor start padded to 32 floats); BatchNorm running statistics in a second flat buffer.
 * ------------------------------------------------------------------------------------------ */
typedef struct dpi_flow_s* dpi_flow_t;

/* tensor ids for dpi_flow_param_offset: per flow block */
enum {
  DPI_T_ACT1_LOC = 0, DPI_T_ACT1_LSI = 1, DPI_T_ACT2_LOC = 2, DPI_T_ACT2_LSI = 3,
  /* per coupling (add DPI_T_COUPLING_STRIDE for coupling2) */
  DPI_T_W1 = 4, DPI_T_B1 = 5, DPI_T_G1 = 6, DPI_T_BE1 = 7, DPI_T_W2 = 8, DPI_T_B2 = 9,
  DPI_T_G2 = 10, DPI_T_BE2 = 11, DPI_T_SCALE = 12, DPI_T_W3 = 13, DPI_T_B3 = 14,
  DPI_T_COUPLING_STRIDE = 11
};
/* running-stat ids for dpi_flow_bn_offset: per coupling */
enum { DPI_R_MEAN1 = 0, DPI_R_VAR1 = 1, DPI_R_MEAN2 = 2, DPI_R_VAR2 = 3 };

/* orders: host int32 [n_flow * ndim], orders[i] = the permutation applied after block i in
 * forward (RealNVP.__init__ :568-577). hidden = int(ndim / (2*seqfrac)). */
/* Device-free layout query (same layout dpi_flow_create uses): param_offsets [n_flow][26] indexed by
 * the tensor ids above (-1 where a tensor does not exist), bn_offsets [n_flow][2][4]. */
int dpi_flow_layout(int ndim, int n_flow, int hidden, int batch_norm, int64_t* param_offsets,
                    int64_t* bn_offsets, int64_t* n_params, int64_t* n_bn);
int dpi_flow_create(dpi_flow_t* out, int device, int ndim, int n_flow, int hidden, int batch_norm,
                    const int32_t* orders);
int dpi_flow_destroy(dpi_flow_t h);
int64_t dpi_flow_param_count(dpi_flow_t h); /* floats in the flat parameter buffer */
int64_t dpi_flow_bn_count(dpi_flow_t h);    /* floats in the flat running-stat buffer */
int64_t dpi_flow_param_offset(dpi_flow_t h, int flow, int tensor_id);
int64_t dpi_flow_bn_offset(dpi_flow_t h, int flow, int coupling /*0|1*/, int stat_id);
/* the fused gather map of stage s (reverse direction), int32[ndim] copied to host `out` */
int dpi_flow_gather_map(dpi_flow_t h, int direction, int stage, int32_t* out);

size_t dpi_flow_workspace_bytes(dpi_flow_t h, int batch);
/* Byte offset inside the workspace of a saved activation of execution stage `stage`. The saved matrices
 * are feature-major [features][ld] with ld >= batch samples per row:
 * what 0: stage output y [ndim][ld]; 1/2: LeakyReLU outputs [hidden][ld]; 3: ZeroFC output
 * [2*(ndim/2)][ld]; 4: batch statistics (mean1, rstd1, mean2, rstd2) [4, hidden];
 * what 5 returns the row stride ld (floats) instead of an offset. */
int64_t dpi_flow_ws_offset(dpi_flow_t h, int batch, int what, int stage);

/* RealNVP.reverse (direction 0, :594-603) / RealNVP.forward (direction 1, :583-592).
 * in/out [batch, ndim], logdet [batch]. train!=0: BatchNorm batch statistics, running stats in
 * `bn_running` updated (momentum 0.1, unbiased variance); train==0: running stats used.
 * The workspace keeps every activation dpi_flow_backward needs. */
int dpi_flow_run(dpi_flow_t h, int direction, int train, int batch, const float* params,
                 float* bn_running, const float* in, float* out, float* logdet, void* workspace,
                 size_t workspace_bytes, dpi_stream_t stream);

/* Backward of a train-mode dpi_flow_run(direction=0): given dL/dout [batch, ndim] and
 * dL/dlogdet [batch], writes dL/din [batch, ndim] (may be NULL) and OVERWRITES the flat
 * gradient buffer `grads` (same layout as params; padding untouched). */
int dpi_flow_backward(dpi_flow_t h, int batch, const float* params, float* grads, const float* in,
                      const float* g_out, const float* g_logdet, float* g_in, void* workspace,
                      size_t workspace_bytes, dpi_stream_t stream);

/* The same backward restricted to the execution stages [stage_lo, stage_hi) (stage S - 1 = the flow block applied last
 * in RealNVP.reverse runs first; S = 2 n_flow). Calling it for consecutive ranges from [.., S) down to [0, ..) equals
 * one dpi_flow_backward; after a range returns (stream order) the gradients of its stages - contiguous in `grads`
 * because the backward visits the flow blocks in parameter order - are final, so a data-parallel caller can
 * all-reduce them while the next range runs. Only the tcgen05 path (dpi_flow_tc_active != 0) supports partial ranges. */
int dpi_flow_backward_range(dpi_flow_t h, int batch, const float* params, float* grads, const float* in,
                            const float* g_out, const float* g_logdet, float* g_in, void* workspace,
                            size_t workspace_bytes, int stage_lo, int stage_hi, dpi_stream_t stream);
/* Can dpi_flow_backward_range split the backward at this batch size? 1: tcgen05 path, 2: small-batch kernel (one persistent
 * launch per range), 0: no (sample-resident kernels, tiled program). */
int dpi_flow_range_ok(dpi_flow_t h, int batch);
/* 1 if a pass at this batch size runs on the tcgen05 path (one launch per operation), 0 if one kernel per pass. */
int dpi_flow_tc_active(dpi_flow_t h, int batch);
/* 1 when `batch` runs RealNVP.reverse (realnvpfc_model.py:594-603) as the cluster-resident tcgen05 pass (flow_cl.cu:
 * ndim 1024, hidden 128, batch a multiple of 4 up to 64 - the reference's default interferometry flow,
 * DPI_interferometry.py:127); bit 1 is set when the backward pass runs there too. Opt-in: DPI_B200_CL=1 (it is
 * parity-green but not yet faster than the fp32 small-batch kernel). */
int dpi_flow_cl_active(dpi_flow_t h, int batch);
/* 1 when `batch` runs RealNVP.reverse as the sample-resident narrow-flow pass (flow_nar.cu: ndim <= 32, hidden <= 160 and a
 * multiple of 4, batch > 64 and a multiple of 4 - the alpha-DPI parameter flow DPIx_interferometry.py:250 and the 2-D toy
 * flow DPI_2Dtoydist.py:52); bit 1 is set when the backward pass runs there too; 4 = only eval-mode passes run there (batches
 * the training plan cannot take, e.g. posterior sampling at 8192). DPI_B200_NAR=0 disables it. */
int dpi_flow_nar_active(dpi_flow_t h, int batch);
/* 1 when batch `batch` runs the sample-resident small-batch kernels (flow_sr.cu: batch <= 64, weights streamed from L2). */
int dpi_flow_sr_active(dpi_flow_t h, int batch);

/* Glow building block (generative_model/glow_model.py:260-268, ZeroConv2d :238-251), forward only: a k x k
 * convolution (k = 1, or k = 3 with one pixel of padding filled with pad_value: 0 for nn.Conv2d(padding=1), 1 for
 * ZeroConv2d) over a FEATURE-major image stack x [c_in][ldx] (position = b * h * w + y * w + x), weight
 * [c_out][c_in][k][k] (the nn.Conv2d layout), y[c][pos] = (conv + bias[c]) * exp(3 * scale[c]); bias and scale may be
 * NULL. One tcgen05 GEMM (3xTF32) with the window gather in the operand pack. */
size_t dpi_conv2d_fm_workspace_bytes(int batch, int h, int w, int c_in, int c_out, int ksize);
int dpi_conv2d_fm_forward(int batch, int h, int w, int c_in, int c_out, int ksize, float pad_value, const float* x,
                          int64_t ldx, const float* weight, const float* bias, const float* scale, float* y, int64_t ldy,
                          void* workspace, size_t workspace_bytes, dpi_stream_t stream);

/* Same convolution with bias + LeakyReLU(slope) (glow_model.py:261-265), and the 1 x 1 convolution with a per-channel
 * y = conv * mul[c] - sub[c] epilogue (InvConv2dLU.reverse followed by ActNorm.reverse, :226-234, :132-148). */
int dpi_conv2d_fm_lrelu(int batch, int h, int w, int c_in, int c_out, int ksize, float pad_value, const float* x, int64_t ldx,
                        const float* weight, const float* bias, float slope, float* y, int64_t ldy, void* workspace,
                        size_t workspace_bytes, dpi_stream_t stream);
int dpi_conv2d_fm_affine(int batch, int h, int w, int c_in, int c_out, const float* x, int64_t ldx, const float* weight,
                         const float* mul, const float* sub, float* y, int64_t ldy, void* workspace, size_t workspace_bytes,
                         dpi_stream_t stream);

/* Weight gradient of the same convolution: g_weight [c_out][c_in][k][k] = sum over positions of g_out[c_out][pos] times the
 * (padded) input window - one tcgen05 GEMM with K = positions and the window gather in the B-operand pack. The data
 * gradient is the forward entry point applied to g_out with the flipped, transposed weight. */
size_t dpi_conv2d_fm_wgrad_workspace_bytes(int batch, int h, int w, int c_in, int c_out, int ksize);
int dpi_conv2d_fm_wgrad(int batch, int h, int w, int c_in, int c_out, int ksize, float pad_value, const float* x, int64_t ldx,
                        const float* g_out, int64_t ldg, float* g_weight, void* workspace, size_t workspace_bytes,
                        dpi_stream_t stream);

/* The non-GEMM pieces of Glow.reverse on feature-major image stacks [channel][b * h * w] (glow.cu), forward only.
 * dpi_glow_bn_rows: nn.BatchNorm2d over each channel row (train != 0: batch statistics, running statistics updated
 *   with momentum 0.1 and the unbiased variance; else running statistics), out may alias x.
 * dpi_glow_lu_inverse: winv [C][C] = (w_p (w_l * l_mask + l_eye) (w_u * u_mask + diag(s_sign * exp(w_s))))^-1, C <= 64.
 * dpi_glow_affine_reverse: rows [C/2, C) of x <- x / exp(tanh(net_out[j])) - net_out[C/2 + j]; logdet[b] -= sum tanh.
 * dpi_glow_prior_sample: z[j] = mean_logsd[j] + exp(mean_logsd[C + j]) * eps[j]; logdet[b] += sum mean_logsd[C + j].
 * dpi_glow_unsqueeze: [C][b, h, w] -> [C/4][b, 2h, 2w], Block.reverse :453-459.
 * dpi_glow_logdet_scalar: logdet[b] += hw * (coef_ws * sum w_s + coef_si * sum log|scale_inv|) (either array may be NULL). */
int dpi_glow_bn_rows(int channels, int64_t positions, int64_t ld, const float* x, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, int train, float eps, float* out, dpi_stream_t stream);
int dpi_glow_lu_inverse(int channels, const float* w_p, const float* w_l, const float* w_s, const float* w_u,
                        const float* l_mask, const float* u_mask, const float* s_sign, const float* l_eye, float* winv,
                        dpi_stream_t stream);
int dpi_glow_affine_reverse(int batch, int hw, int channels, const float* net_out, int64_t ld_net, float* x, int64_t ldx,
                            float* logdet, dpi_stream_t stream);
int dpi_glow_prior_sample(int batch, int hw, int channels, const float* mean_logsd, int64_t ld_ml, const float* eps,
                          int64_t ld_eps, float* z, int64_t ld_z, float* logdet, dpi_stream_t stream);
int dpi_glow_unsqueeze(int batch, int h, int w, int channels, const float* x, float* out, dpi_stream_t stream);
int dpi_glow_logdet_scalar(int batch, int hw, int channels, const float* w_s, float coef_ws, const float* scale_inv,
                           float coef_si, float* logdet, dpi_stream_t stream);

/* The density direction (Glow.forward, :508-524), forward only. dpi_glow_lu_weight: the 1x1 convolution matrix itself;
 * dpi_glow_affine_forward: rows [C/2, C) of x <- (x + t) * exp(tanh(log_s0)), logdet[b] += sum tanh; dpi_glow_prior_forward:
 * z_eps = (z - mean) * exp(-log_sd), logdet[b] -= sum log_sd, log_p[b] += sum(-0.5 log 2 pi - 0.5 z_eps^2);
 * dpi_glow_squeeze: [C][b, 2h, 2w] -> [4C][b, h, w] (:395-398); dpi_glow_row_affine: out = (x + add[c]) / div[c]. */
int dpi_glow_lu_weight(int channels, const float* w_p, const float* w_l, const float* w_s, const float* w_u,
                       const float* l_mask, const float* u_mask, const float* s_sign, const float* l_eye, float* weight,
                       dpi_stream_t stream);
int dpi_glow_affine_forward(int batch, int hw, int channels, const float* net_out, int64_t ld_net, float* x, int64_t ldx,
                            float* logdet, dpi_stream_t stream);
int dpi_glow_prior_forward(int batch, int hw, int channels, const float* mean_logsd, int64_t ld_ml, const float* z,
                           int64_t ld_z, float* z_eps, int64_t ld_eps, float* logdet, float* log_p, dpi_stream_t stream);
int dpi_glow_squeeze(int batch, int h_out, int w_out, int channels_in, const float* x, float* out, dpi_stream_t stream);
int dpi_glow_row_affine(int channels, int64_t positions, const float* x, const float* add, const float* div, float* out,
                        dpi_stream_t stream);

/* Training backward of the sampling direction (Glow.reverse, glow_model.py:526-539), op by op; stacks feature-major
 * [channel][b * hw + p], per-channel reductions in a fixed order. hw_gld points at ONE device float = h w sum_b g_logdet[b]
 * (dpi_glow_scaled_sum makes it).
 *  actnorm_reverse_bwd : x = v si - loc (:132-148) given y = x: g_v = g si, g_loc = -sum g, g_si = sum g v + hw_gld / si
 *  affine_reverse_bwd  : x_b = b / exp(tanh ls0) - t (:299-320); net_out = [ls0 ; t], x_mid = the flow's state after the
 *                        coupling; g rows [C/2, C) become the gradient of b IN PLACE, g_out = gradient of net_out
 *  zeroconv_bwd        : out = zc exp(3 scale) (:246-251); g <- g exp(3 scale) in place, g_scale = 3 sum g out, g_bias = sum
 *  bn_lrelu_bwd        : Conv -> LeakyReLU -> BatchNorm2d in training mode (:260-268); lrelu_out = the BatchNorm input;
 *                        g <- gradient of the conv output in place; g_gamma, g_beta, g_bias (the conv bias)
 *  lu_bwd              : InvConv2dLU.reverse x = W^-1 u (:215-234) given g_winv = sum_m g[c][m] u[c'][m]: gradients of w_l,
 *                        w_u, w_s (incl. the -h w sum w_s logdet term); scratch: 3 C^2 floats; C <= 64
 *  prior_sample_bwd    : z = mean + exp(lsd) eps, logdet += sum lsd (:436-452): g_ml = gradient of [mean ; lsd], g_eps */
int dpi_glow_actnorm_reverse_bwd(int channels, int64_t positions, const float* y, const float* loc, const float* scale_inv,
                                 const float* g, const float* hw_gld, float* g_v, float* g_loc, float* g_scale_inv,
                                 dpi_stream_t stream);
int dpi_glow_affine_reverse_bwd(int batch, int hw, int channels, const float* net_out, int64_t ld_net, const float* x_mid,
                                int64_t ldx, float* g, int64_t ldg, const float* g_logdet, float* g_out, int64_t ldo,
                                dpi_stream_t stream);
int dpi_glow_zeroconv_bwd(int channels, int64_t positions, float* g, const float* out, const float* scale, float* g_scale,
                          float* g_bias, dpi_stream_t stream);
int dpi_glow_bn_lrelu_bwd(int channels, int64_t positions, float* g, const float* lrelu_out, const float* gamma, float eps,
                          float slope, float* g_gamma, float* g_beta, float* g_bias, dpi_stream_t stream);
int dpi_glow_row_sum(int channels, int64_t positions, const float* a, float coef, float* out, dpi_stream_t stream);
int dpi_glow_lu_bwd(int channels, const float* w_p, const float* w_l, const float* w_s, const float* w_u, const float* l_mask,
                    const float* u_mask, const float* s_sign, const float* l_eye, const float* w_inv, const float* g_winv,
                    const float* hw_gld, float* scratch, float* g_wl, float* g_wu, float* g_ws, dpi_stream_t stream);
int dpi_glow_prior_sample_bwd(int batch, int hw, int channels, const float* mean_logsd, int64_t ld_ml, const float* eps,
                              int64_t ld_eps, const float* g_z, int64_t ld_gz, const float* g_logdet, float* g_ml, int64_t ld_gml,
                              float* g_eps, int64_t ld_geps, dpi_stream_t stream);
int dpi_glow_scaled_sum(int n, const float* v, float coef, float* out, dpi_stream_t stream);

/* Host-only: the tile plan the tcgen05 path (batch >= 96, flow_tc.cu) would use for one coupling-net GEMM
 * C[M, N] = A[M, K] B[N, K]^T on a GPU with n_sm SMs. out[0..9] = column tile TN, k-tile KT, accumulators per CTA
 * MH, 128-row tiles, column tiles, k tiles, K splits, k tiles per split, pipeline stages, dynamic shared-memory
 * bytes. No CUDA call is made (usable without a device; the CPU test suite checks the planner's invariants). */
int dpi_flow_tc_plan(int M, int N, int K, int n_sm, int32_t* out);

/* Debug/profiling aid (synchronises; never on the hot path): when enabled, CTA 0 stamps %globaltimer at
 * the start and at the end of its work in every phase of the next launches. dpi_flow_debug_read returns
 * the number of phases of the last launch and fills stamps[2*i], stamps[2*i+1] (ns) and ops[i]
 * (op*100 + layer*10 + has_second_subop). */
int dpi_flow_debug_timing(dpi_flow_t h, int enable);
int dpi_flow_debug_read(dpi_flow_t h, uint64_t* stamps, int32_t* ops, int max_phases);
/* 0: one launch per phase; 1: one persistent cooperative launch with grid barriers */
int dpi_flow_set_mode(dpi_flow_t h, int persistent, int ctas_per_sm);

/* ------------------------------------------------------------------------------------------
 * Image head: positivity + scale (DPI_interferometry.py:205-211, DPI_MRI.py:174-179) fused
 * with the image priors (interferometry_helpers.py Loss_l1 :349, Loss_TSV :354, Loss_TV :358,
 * Loss_flux :363, Loss_center :369-382, Loss_cross_entropy :384-387; MRI_helpers.py:28-38).
 * ------------------------------------------------------------------------------------------ */
enum { DPI_PRIOR_L1 = 0, DPI_PRIOR_TSV = 1, DPI_PRIOR_TV = 2, DPI_PRIOR_FLUX = 3,
       DPI_PRIOR_CENTER = 4, DPI_PRIOR_MEM = 5, DPI_PRIOR_COUNT = 6 };

/* img = softplus(x) * exp(*log_scale); det[b] = sum(x - softplus(x)) + *log_scale * npix^2 */
int dpi_positivity_forward(int batch, int npix, const float* x, const float* log_scale, float* img,
                           float* det, dpi_stream_t stream);
/* g_x = (g_img + g_img2) * exp(ls) * sigmoid(x) + g_det[b] * (1 - sigmoid(x));
 * g_log_scale_partial[b] = sum_p (g_img + g_img2)*img + npix^2 * g_det[b]  (caller sums over b).
 * g_img, g_img2, g_det, g_log_scale_partial may each be NULL. */
int dpi_positivity_backward(int batch, int npix, const float* x, const float* log_scale,
                            const float* img, const float* g_img, const float* g_img2,
                            const float* g_det, float* g_x, float* g_log_scale_partial,
                            dpi_stream_t stream);

/* All six priors of each image in one pass: values[b, DPI_PRIOR_COUNT]. `prior_img` (may be NULL
 * unless MEM is wanted) is the y_true of Loss_cross_entropy; `flux`/`center` as in the factories.
 * If g_img != NULL also writes g_img[b,:] = sum_t gw[b,t] * d values[b,t] / d img (gw [batch, 6]). */
int dpi_image_priors(int batch, int npix, const float* img, const float* prior_img, float flux,
                     float center, float* values, const float* gw, float* g_img, dpi_stream_t stream);
/* Fused positivity + priors: reads x, writes img and det, then as dpi_image_priors. */
int dpi_image_head(int batch, int npix, const float* x, const float* log_scale, float* img, float* det,
                   const float* prior_img, float flux, float center, float* values, const float* gw,
                   float* g_img, dpi_stream_t stream);
/* Loss assembly (DPI_interferometry.py:224-229, DPI_MRI.py:186-193), batch means on the device.
 * data64 / data32: per-sample data terms (either may be NULL) with weights w64 / w32; priors [batch, 6]
 * with host weights prior_weights[6] (both may be NULL); total logdet = logdet_flow + det.
 * out[0] = loss, [1] = mean data64, [2] = mean data32, [3] = mean weighted prior, [4] = mean logdet,
 * [5..10] = mean of each prior value. */
int dpi_step_terms(int batch, const double* data64, float w64, const float* data32, float w32,
                   const float* priors, const float* prior_weights /*host*/, const float* logdet_flow,
                   const float* det, float w_logdet, float* out /*[16]*/, dpi_stream_t stream);
/* out[0] = sum(in[0..n)); sq_out[0] = out[0]^2 (either may be NULL). Deterministic, one CTA. */
int dpi_reduce_sum(int n, const float* in, float* out, double* sq_out, dpi_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Interferometry operator: interferometry_helpers.py eht_observation_pytorch :235-292
 * (ttype='direct': torch_complex_matmul :36-39) and the chi^2 terms :299-347.
 * ------------------------------------------------------------------------------------------ */
typedef struct dpi_eht_s* dpi_eht_t;
/* dft_mat host fp32 [npix*npix, n_vis, 2]; index/sign tables host, as Obs_params_torch returns. */
int dpi_eht_create(dpi_eht_t* out, int device, int npix, int n_vis, const float* dft_mat, int n_cp,
                   const int64_t* cp_ind /*[3*n_cp]*/, const double* cp_sign /*[3*n_cp]*/, int n_ca,
                   const int64_t* ca_ind /*[4*n_ca]*/);
int dpi_eht_destroy(dpi_eht_t h);
size_t dpi_eht_workspace_bytes(dpi_eht_t h, int batch);
/* vis [B,2,n_vis] fp32, visamp [B,n_vis] fp32, cphase [B,n_cp] fp64 (degrees), logcamp [B,n_ca] */
int dpi_eht_forward(dpi_eht_t h, int batch, const float* img, float* vis, float* visamp,
                    double* cphase, float* logcamp, void* workspace, size_t workspace_bytes,
                    dpi_stream_t stream);
/* any upstream gradient may be NULL; g_img [B, npix*npix] is overwritten */
int dpi_eht_backward(dpi_eht_t h, int batch, const float* vis, const float* g_vis,
                     const float* g_visamp, const double* g_cphase, const float* g_logcamp,
                     float* g_img, void* workspace, size_t workspace_bytes, dpi_stream_t stream);

/* Fused data term of DPI_interferometry.py:221-229. Given vis from dpi_eht_forward:
 * loss_cp[b] (fp64) = Loss_angle_diff, loss_ca[b] = Loss_logca_diff2, and
 * g_img = d(w_cp * sum_b loss_cp[b] + w_ca * sum_b loss_ca[b]) / d img. */
int dpi_eht_data_grad(dpi_eht_t h, int batch, const float* vis, const float* cphase_true,
                      const float* sigma_cp, const float* logcamp_true, const float* sigma_ca,
                      float w_cp, float w_ca, double* loss_cp, float* loss_ca, float* g_img,
                      void* workspace, size_t workspace_bytes, dpi_stream_t stream);

/* chi^2 data terms. mode: */
enum { DPI_CHI2_ANGLE = 0,   /* Loss_angle_diff :299 (y in degrees, pred fp64) */
       DPI_CHI2_SQ = 1,      /* Loss_logca_diff2 :318 / Loss_visamp_diff :342 */
       DPI_CHI2_LOGSQ = 2,   /* Loss_logamp_diff :334 / Loss_logca_diff :310 */
       DPI_CHI2_VIS = 3 };   /* Loss_vis_diff :326 (true [2,n], pred [B,2,n]) */
/* loss[b] (fp64 for ANGLE, fp32 otherwise - matching the reference's promotion).
 * If g_loss != NULL also writes g_pred = g_loss[b] * d loss[b] / d pred. */
int dpi_chi2(int mode, int batch, int n, const float* y_true, const void* y_pred, const float* sigma,
             void* loss, const void* g_loss, void* g_pred, dpi_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * alpha-DPI (DPIx_interferometry.py): geometric crescent renderer geometric_model.py
 * SimpleCrescentNuisanceFloor_Param2Img :237-389 (flux_flag=False: 6 + 6 g parameters), sigmoid squashing
 * :339,344 and the alpha-divergence objective :350-380.
 * ------------------------------------------------------------------------------------------ */
/* params [batch, 6 + 6*n_gaussian] in (0, 1) -> img [batch, npix, npix] (unit flux); forward() :333-381 */
int dpi_crescent_forward(int batch, int npix, int n_gaussian, float fov, const float* params, float* img,
                         dpi_stream_t stream);
/* g_params[b, :] = gscale[b] * sum_p g_img[b, p] * d img[b, p] / d params[b, :]  (gscale may be NULL = 1) */
int dpi_crescent_backward(int batch, int npix, int n_gaussian, float fov, const float* params,
                          const float* g_img, const float* gscale, float* g_params, dpi_stream_t stream);
/* params = sigmoid(ps); det[b] = sum_j(-ps - 2 softplus(-ps))   (DPIx_interferometry.py:339,344) */
int dpi_sigmoid_forward(int batch, int n, const float* ps, float* params, float* det, dpi_stream_t stream);
/* g_ps = g_params * s (1 - s) + g_det[b] * (1 - 2 s); g_params / g_det may be NULL */
int dpi_sigmoid_backward(int batch, int n, const float* ps, const float* g_params, const float* g_det,
                         float* g_ps, dpi_stream_t stream);
/* The objective of DPIx_interferometry.py:350-380 over the whole batch (data_product 'cphase_logcamp'):
 *   loss_i = data_weight * 0.5 (w_ca loss_ca[i] + w_cp loss_cp[i]) / scale_factor - (logdet_flow[i] + det[i]) - 0.5 |z_i|^2
 *   weights w_i = 1 / global_batch (alpha == 1: KL) or softmax_i(-(1 - alpha) loss_i), detached
 * outputs: terms[0] = sum_i w_i scale_factor loss_i, [1] = loss_orig, [2] = mean loss_cp, [3] = mean loss_ca,
 * [4] = mean logdet, [5] = mean loss_i; gscale[i] = 0.5 data_weight w_i (the factor of
 * d(w_ca loss_ca + w_cp loss_cp)/d img); g_logdet[i] = -w_i scale_factor; loss_i (may be NULL).
 * mode 0: softmax over this batch. Data parallel, softmax over the global batch: mode 1 writes stats[0] = max_i
 * t_i (t = -(1 - alpha) loss); mode 2 reads the all-reduced max and writes stats[1] = sum_i exp(t_i - max);
 * mode 3 reads the all-reduced (max, sum) and finishes. */
int dpi_alpha_objective(int batch, int nparams, const float* z, const double* loss_cp, const float* loss_ca,
                        const float* logdet_flow, const float* det, float w_cp, float w_ca, float scale_factor,
                        float data_weight, float alpha, int mode, int global_batch, float* loss_i,
                        float* gscale, float* g_logdet, float* terms /*[8]*/, float* stats /*[2]*/,
                        dpi_stream_t stream);

/* y[r][i] = s[r] x[r][i] (s NULL: 1; y NULL: skipped) and dot[r] = coef sum_i x[r][i] b[r][i] (dot NULL: skipped) over rows of
 * n floats: the total-flux factor of the crescent renderer with flux_flag=True (geometric_model.py:373-376) and its gradient. */
int dpi_row_scale_dot(int rows, int n, const float* x, const float* s, float* y, const float* b, float coef, float* dot,
                      dpi_stream_t stream);

/* interferometry_helpers.py torch_complex_mul :30-34. x, out: [outer][2][n] (real row, imaginary row), y: [2][n];
 * out = x * y, or x * conj(y) when conj != 0 (the gradient wrt x of the former). */
int dpi_complex_mul(long long outer, int n, int conj, const float* x, const float* y, float* out, dpi_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * MRI operator: DPI_MRI.py fft2c_torch :70-74 + MRI_helpers.py Loss_kspace_diff2 :23-26.
 * ------------------------------------------------------------------------------------------ */
/* kspace[b, ky, kx, 2] = ortho FFT2 of the real image */
int dpi_fft2c_forward(int batch, int npix, const float* img, float* kspace, dpi_stream_t stream);
/* g_img = adjoint applied to g_kspace (real part) */
int dpi_fft2c_backward(int batch, int npix, const float* g_kspace, float* g_img, dpi_stream_t stream);
/* MRI_helpers.py Loss_kspace_diff :18-21 (power 1) / Loss_kspace_diff2 :23-26 (power 2):
 * loss[b] = mean_i |y_pred[b][i] - y_true[i]|^power / sigma^power over the n values of a sample (y_true broadcast over the
 * batch). With g_loss [batch] and g_pred [batch][n] also the gradient wrt y_pred; loss may then be NULL. */
int dpi_kspace_loss(int power, int batch, int n, float sigma, const float* y_true, const float* y_pred, float* loss,
                    const float* g_loss, float* g_pred, dpi_stream_t stream);
/* Fused data term of DPI_MRI.py:181-182: loss[b] = mean((fft2c(img)*mask - kspace_true)^2)/sigma^2/mean_mask
 * and, if g_img != NULL, g_img[b,:] = gw[b] * d loss[b]/d img. mask, kspace_true [npix,npix,2]. */
int dpi_mri_data_loss(int batch, int npix, const float* img, const float* mask, const float* kspace_true,
                      float sigma, float mean_mask, float* loss, const float* gw, float* g_img,
                      dpi_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser: nn.utils.clip_grad_norm_ + optim.Adam (DPI_interferometry.py:172,233-235).
 * ------------------------------------------------------------------------------------------ */
size_t dpi_clip_adam_workspace_bytes(int64_t n);
/* sq_out[0] (device fp64) = sum g^2, deterministic two-level reduction in one launch. */
int dpi_grad_sqnorm(int64_t n, const float* g, double* sq_out, void* workspace, size_t workspace_bytes,
                    dpi_stream_t stream);
/* One fused pass: total norm = sqrt(sum sq_norms[0..n_sq)) * grad_scale,
 * coef = min(1, clip / (norm + 1e-6)) (clip <= 0 disables), g' = coef * grad_scale * g, Adam update
 * with the step number read from device memory (graph-replay safe).
 * norm_out (may be NULL): [0] = norm before clipping, [1] = coef. */
int dpi_clip_adam(int64_t n, float* p, const float* g, float* m, float* v, const double* sq_norms,
                  int n_sq, float clip, float lr, float beta1, float beta2, float eps,
                  const int64_t* step_dev, float grad_scale, float* norm_out, dpi_stream_t stream);
int dpi_step_increment(int64_t* step_dev, dpi_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DPI_B200_H */
