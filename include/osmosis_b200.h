/* osmosis_b200.h - C ABI of the B200-native Osmosis guided-sampling path.
 *
 * The reference (osmosis-diffusion/osmosis-diffusion-code) is pure Python/PyTorch and has no FFI; its
 * "plugin API" is the Python registry/factory surface (create_model / create_sampler / get_operator /
 * get_conditioning_method).  Each entry point below is what that surface binds underneath in this
 * repo; the comment on each names the reference call it replaces (paths relative to the reference
 * root).  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on failure; osm_last_error_string()
 *     returns the message of the last failure on the calling thread.  Nothing throws.
 *   - all data pointers are DEVICE pointers owned by the caller unless the name says `host`.
 *   - tensors crossing the boundary are NCHW fp32, contiguous (the reference's layout).
 *   - every launch goes to the given cudaStream_t (passed as void*); no hidden host sync, no
 *     allocation after osm_unet_create / osm_unet_load_param / osm_unet_bind.
 *   - B > 1 means B independent B=1 reference runs (per-image loss norm / aux means).
 */
#ifndef OSMOSIS_B200_H_
#define OSMOSIS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSM_OK 0
#define OSM_ERR_INVALID (-1)
#define OSM_ERR_CUDA (-2)
#define OSM_ERR_STATE (-3)

const char* osm_last_error_string(void);
/* ABI version of this header; bumped on any signature change. */
int osm_abi_version(void);

/* ------------------------------------------------------------------ UNet ------------------------------
 * replaces guided_diffusion/unet.py: create_model :27-98, UNetModel.__init__ :503-695,
 * UNetModel.forward :713-742 and (for the input gradient) the autograd graph that
 * condition_methods.py:186-191 back-propagates through.                                              */
typedef struct osm_unet_config {
  int in_channels;        /* 4  (RGBD)                                   utils.py:265-288            */
  int out_channels;       /* 8  (eps | learned-range variance)                                       */
  int model_channels;     /* num_channels                                unet.py:29                  */
  int num_res_blocks;     /*                                             unet.py:30                  */
  int num_levels;         /* len(channel_mult)                                                       */
  int channel_mult[8];    /*                                             unet.py:47-59               */
  int num_attention_ds;   /* number of entries in attention_ds                                       */
  int attention_ds[8];    /* downsample rates with attention             unet.py:61-66               */
  int num_heads;          /* used when num_head_channels == -1                                       */
  int num_head_channels;  /* 64 in every shipped config                                              */
  int conv_mode;          /* 0: tcgen05 TF32 tensor-core convs (product path), 1: fp32 CUDA-core convs
                             (exact mode used to separate precision from logic in parity tests)       */
} osm_unet_config;

typedef struct osm_unet* osm_unet_t;

int osm_unet_create(const osm_unet_config* cfg, osm_unet_t* out);
int osm_unet_destroy(osm_unet_t h);
/* Parameter table in guided-diffusion state_dict order/names (OIHW fp32).  Host-only, no GPU needed. */
int osm_unet_param_count(osm_unet_t h);
int osm_unet_param_info(osm_unet_t h, int index, const char** name, int* ndim, int64_t shape[4]);
/* Copies + repacks one state_dict tensor (host fp32, contiguous) into the engine's device layout.
 * replaces model.load_state_dict (unet.py:94-97).                                                    */
int osm_unet_load_param(osm_unet_t h, const char* name, const float* host_data, int64_t numel, void* stream);
/* Bytes of caller-provided device workspace needed for batch B at HxW (activations kept for the input
 * VJP, gradients, scratch).  Plans the engine for that shape: an existing binding is dropped and
 * osm_unet_bind must be called again before the next forward.                                         */
int64_t osm_unet_workspace_bytes(osm_unet_t h, int B, int H, int W);
/* Builds the launch plan (buffer offsets, TMA descriptors) for (B,H,W) on `workspace`.               */
int osm_unet_bind(osm_unet_t h, int B, int H, int W, void* workspace, int64_t workspace_bytes);
/* out[B,out_ch,H,W] = UNet(x[B,in_ch,H,W], t[B]); t are the (possibly fractional) model timesteps as
 * fp32 (the reference's timesteps.float(), nn.py:116).  Keeps what the VJP needs in the workspace.   */
int osm_unet_forward(osm_unet_t h, const float* x, const float* t, float* out, void* stream);
/* grad_x[B,in_ch,H,W] = d<out, grad_out>/dx for the most recent forward (input gradient only - the
 * reference requests no weight gradients: condition_methods.py:188-191).                             */
int osm_unet_vjp_input(osm_unet_t h, const float* grad_out, float* grad_x, void* stream);
/* Number of kernel launches one forward / one vjp issues (bench.py's gpu_launches).                  */
int osm_unet_launch_count(osm_unet_t h, int which /*0 fwd, 1 vjp*/);
/* Algorithmic FLOPs (2*MAC) of one forward for the bound shape (convs + attention + linear).         */
double osm_unet_forward_flops(osm_unet_t h);
/* Measurement hook: runs program `which` (0 forward ops, 1 input-VJP ops; on the buffers of the last
 * forward) with a CUDA-event pair around every op on `stream`, synchronises, and returns the number of
 * ops n (<= cap) with per-op device time ms[n], kind[n] (0 conv, 1 gn_stats, 2 gn_apply, 3 gn_bwd,
 * 4 attention fwd, 5 attention bwd, 6 linear), algorithmic flops[n] / HBM bytes[n], dims6[6n]
 * (conv: H,W,Cin,Cout,taps,is_dgrad; norm: H,W,C; attention: L,C,heads).  Host arrays.                */
int osm_unet_profile_ops(osm_unet_t h, int which, void* stream, int cap, float* ms, int* kinds, double* flops,
                         double* bytes, int* dims6);

/* ------------------------------------------------------------- sampler elementwise ---------------------
 * coef: [T][OSM_COEF_COLS] fp32 rows = {sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1,
 * posterior_mean_coef2, log(beta), posterior_log_variance_clipped, alphas_cumprod, alphas_cumprod_prev,
 * log(posterior_variance), log(append(posterior_variance[1], betas[1:])), 1 / coef1, coef2 / coef1}, each the fp32 rounding of the
 * reference's float64 table entry (posterior_mean_variance.py:265-269).  t_idx[B]: int32 respaced index. */

/* replaces EpsilonXMeanProcessor.get_mean_and_xstart (posterior_mean_variance.py:104-136) and
 * LearnedRangeVarianceProcessor.get_variance (:227-258), as called from p_mean_variance
 * (gaussian_diffusion.py:345-365).  model_out [B,2C,HW]; x, x0, mean, logvar [B,C,HW].                */
int osm_posterior_fwd(const float* coef, const int32_t* t_idx, const float* x, const float* model_out, float* x0,
                      float* mean, float* logvar, int B, int C, int HW, void* stream);
/* VJP of the above: any of g_x0 / g_mean / g_logvar may be NULL (treated as zero).
 * g_x [B,C,HW] (direct path through x), g_model_out [B,2C,HW].                                        */
int osm_posterior_vjp(const float* coef, const int32_t* t_idx, const float* g_x0, const float* g_mean,
                      const float* g_logvar, float* g_x, float* g_model_out, int B, int C, int HW, void* stream);
/* The same two for every mean / variance processor of the reference's registries and `clip_denoised`
 * (posterior_mean_variance.py:41-50 process_xstart, :54-136 mean processors, :173-258 variance processors).
 * flags = OSM_POST_CLIP | OSM_POST_MEAN_* | OSM_POST_VAR_*  (0 = epsilon + learned_range, no clamp).  With the clamp, x0 is
 * clamped to [-1, 1] before the posterior mean; the VJP recomputes the unclamped x0 from x / model_out and passes the
 * gradient only where it lay inside [-1, 1], as torch's clamp backward does.                                        */
#define OSM_COEF_COLS 12
#define OSM_POST_CLIP 0x1
#define OSM_POST_MEAN_MASK 0xF0
#define OSM_POST_MEAN_EPSILON 0x00   /* :104-136 */
#define OSM_POST_MEAN_STARTX 0x10    /* :76-101  x0 = model_out                          */
#define OSM_POST_MEAN_PREVX 0x20     /* :54-73   mean = model_out, x0 = mean / coef1 - coef2 / coef1 x */
#define OSM_POST_VAR_MASK 0xF00
#define OSM_POST_VAR_LEARNED_RANGE 0x000 /* :227-258 */
#define OSM_POST_VAR_LEARNED 0x100       /* :217-224 log variance = model_out[:, C:]     */
#define OSM_POST_VAR_FIXED_SMALL 0x200   /* :173-191 log(posterior_variance)             */
#define OSM_POST_VAR_FIXED_LARGE 0x300   /* :194-214 log(append(posterior_variance[1], betas[1:])) */
int osm_posterior_fwd_ex(const float* coef, const int32_t* t_idx, const float* x, const float* model_out, float* x0,
                         float* mean, float* logvar, int B, int C, int HW, int flags, void* stream);
int osm_posterior_vjp_ex(const float* coef, const int32_t* t_idx, const float* g_x0, const float* g_mean,
                         const float* g_logvar, float* g_x, float* g_model_out, int B, int C, int HW, const float* x,
                         const float* model_out, int flags, void* stream);
/* replaces condition_methods.py:211-223 + gaussian_diffusion.py:266-271:
 *   x_out = mean - scale[c] * clamp(g_a + g_b, +-clip) + [t_idx != 0] * exp(0.5*logvar) * noise
 * g_b may be NULL; clip < 0 disables clamping.  The clamped-before-scale raw gradient g_a+g_b is written to
 * grad_out if non-NULL (the `gradients` the reference returns, condition_methods.py:224).            */
int osm_sampler_update(const float* mean, const float* g_a, const float* g_b, const float* scale4, float clip,
                       const float* logvar, const float* noise, const int32_t* t_idx, float* x_out, float* grad_out,
                       int B, int C, int HW, void* stream);
/* osm_sampler_update with the order of the rgb_guidance branch (noise_first != 0): DDPM.p_sample adds the noise
 * (gaussian_diffusion.py:499-501), then the `ps` conditioning subtracts scale * gradient (condition_methods.py:248):
 *   x_out = (mean + [t_idx != 0] * exp(0.5*logvar) * noise) - clamp(g_a + g_b) * scale[c]                          */
int osm_sampler_update_ex(const float* mean, const float* g_a, const float* g_b, const float* scale4, float clip,
                          const float* logvar, const float* noise, const int32_t* t_idx, float* x_out, float* grad_out,
                          int B, int C, int HW, int noise_first, void* stream);
/* replaces DDIM.p_sample after p_mean_variance (gaussian_diffusion.py:506-535): eps from (x, pred_xstart x0), sigma from
 * eta and the alphas_cumprod columns of coef, out = x0 sqrt(abar_prev) + sqrt(1 - abar_prev - sigma^2) eps + [t != 0] sigma z */
int osm_ddim_sample(const float* coef, const int32_t* t_idx, const float* x, const float* x0, const float* noise, float eta,
                    float* out, int B, int C, int HW, void* stream);
/* replaces ConditioningMethod.grad_and_value (gaussian branch, condition_methods.py:36-40) for the identity
 * `rgb_guidance` operator (measurements.py:80-97), per image: losses[b] = || y_b - x0_b[:3] ||_2 and
 * g_x0 [B,C,HW] = d losses[b] / d x0 (zero for channels >= 3).  y [B,3,HW].                                         */
int osm_ps_guidance(const float* x0, const float* y, float* g_x0, float* losses, int B, int C, int HW, void* stream);
/* replaces osmosis_utils/diffusion.py:122 (GaussianDiffusion.inverse, unguided ancestral update):
 *   x = (x - c_eps * eps) * c_x + c_z * z   with eps = model_out[:, :C]                               */
int osm_ddpm_uncond_update(float* x, const float* model_out, const float* z, float c_x, float c_eps, float c_z, int B,
                           int C, int C_model_out, int HW, void* stream);

/* ------------------------------------------------------- measurement operators + guidance --------------*/
#define OSM_OP_UNDERWATER_REVISED 0 /* measurements.py:211-329  phi_a, phi_b, phi_inf  [B,3] each      */
#define OSM_OP_UNDERWATER 1         /* measurements.py:332-433  phi_ab, phi_inf                         */
#define OSM_OP_HAZE 2               /* measurements.py:107-208  scalar phi_ab, phi_inf                  */
#define OSM_DEPTH_ORIGINAL 0        /* utils.py:560-561  0.5*(d+1)                                      */
#define OSM_DEPTH_GAMMA 1           /* utils.py:557-558  ((d+v0)*v1)^v2                                 */
#define OSM_DEPTH_MOVE 2            /* utils.py:554-555  d+v0                                           */

#define OSM_OPT_SGD 0
#define OSM_OPT_ADAM 1
#define OSM_LOSS_NORM 0             /* condition_methods.py:127-130  ||r||_2                            */
#define OSM_LOSS_MSE 1              /* condition_methods.py:133-138  mean(r^2) per image                */

typedef struct osm_guidance_params {
  int op_kind;           /* OSM_OP_*                                                                    */
  int depth_kind;        /* OSM_DEPTH_* of the operator                                                 */
  float depth_val[3];
  int weight_kind;       /* 0: none, 1: depth (condition_methods.py:121-125, utils.py:674-700)          */
  int weight_depth_kind; /* OSM_DEPTH_* of weight_function                                              */
  float weight_val[3];
  float eta[3];          /* SGD step per phi group in get_variable_list() order (0 if learn flag off)   */
  int n_iter;            /* inner iterations when not frozen (sample_pattern.n_iter)                    */
  float gamma_avrg;      /* aux_loss.avrg_loss weight, 0 if absent (losses.py:29-45)                    */
  float gamma_val;       /* aux_loss.val_loss weight, 0 if absent  (losses.py:51-62)                    */
  int loss_kind;         /* OSM_LOSS_NORM / OSM_LOSS_MSE: loss_function (condition_methods.py:127-138)  */
  int optimizer;         /* OSM_OPT_SGD (sgd / GD, measurements.py:279-303) or OSM_OPT_ADAM (utils.py:499-500,
                            torch.optim.Adam defaults: betas 0.9 / 0.999, eps 1e-8, lr = eta per group)             */
  float* opt_state;      /* OSM_OPT_ADAM: device [B][19] = {exp_avg[9], exp_avg_sq[9], step}, zero-initialised by the
                            caller and kept across calls like the reference's optimizer object; NULL for SGD        */
  int phi_batch;         /* rows of the caller's phi [phi_batch][9] (and opt_state) buffers: the call fails unless it equals B
                            (0 = not stated, unchecked).  The reference builds phi with data.batch_size rows
                            (measurements.py:225-232) independently of the batch it is later run on                  */
} osm_guidance_params;

/* replaces Operator.forward (measurements.py:138-151, 251-264, 363-376): out[B,3,HW] = A_phi(x[B,4,HW]).
 * phi [B,9] = {a0,a1,a2, b0,b1,b2, inf0,inf1,inf2}; tied operators read a* for both, haze reads a0.   */
int osm_operator_forward(int op_kind, int depth_kind, const float depth_val[3], const float* x, const float* phi,
                         float* out, int B, int HW, void* stream);
/* replaces the whole inner loop of PosteriorSamplingOsmosis.conditioning (condition_methods.py:146-209):
 * grad_and_value (:109-144) + AuxiliaryLoss.forward (losses.py:67-83) + backward w.r.t. phi and x_0_hat +
 * operator.optimize (measurements.py:266-303), 1 evaluation if freeze else n_iter, SGD after every
 * evaluation, x-gradient taken at the last evaluation.  freeze_flag: device int32[1] (non-zero = frozen),
 * so the same captured launch serves both phases.
 *   x0 [B,4,HW], y [B,3,HW] -> g_x0 [B,4,HW] (d total_loss / d x_0_hat), phi [B,9] updated in place,
 *   losses [B,4] = {norm loss (sep_loss), avrg term, val term, total}.                                 */
int osm_guidance_phi_loop(const osm_guidance_params* p, const float* x0, const float* y, float* phi,
                          const int32_t* freeze_flag, float* g_x0, float* losses, int B, int HW, void* stream);

/* ------------------------------------------------------------ post-processing of finished samples --------
 * Replaces the per-image CPU block after p_sample_loop in osmosis_sampling.py:207-292 (batched, on the device).
 * x0 [B,4,HW] (pred_xstart), y [B,3,HW] in [-1,1] (the measurement), phi [B,9] as in osm_guidance_phi_loop.
 *   rgb_clip   clamp(0.5 (x0_rgb + 1), 0, 1)                                   (:214-215)
 *   degraded   2 A_phi(x0) - 1, norm_loss[b] = || degraded - y ||_2            (:245-251, :283-291)
 *   recon      exp(phi_a d) (0.5 (y + 1) - backscatter)                        (:254-256, :286-287)          */
int osm_postprocess(int op_kind, int depth_kind, const float depth_val[3], const float* x0, const float* y, const float* phi,
                    float* rgb_clip, float* degraded, float* recon, float* norm_loss, int B, int HW, void* stream);
/* min_max_norm_range (q_low = 0, q_high = 1) / min_max_norm_range_percentile of each of the B planes of n elements
 * (osmosis_utils/utils.py:46-114): clip to torch.quantile(img, q) (linear interpolation), map [min, max] to [vmin, vmax]. */
int osm_minmax_percentile(const float* img, float* out, int B, int n, float q_low, float q_high, float vmin, float vmax,
                          void* stream);
/* depth_tensor_to_color_image (utils.py:748-763): matplotlib-style lookup of img in [0,1] in a [256][3] table -> [B,3,n] */
int osm_colormap(const float* img, const float* lut, float* out, int B, int n, void* stream);

/* ------------------------------------------------------------ input pipeline ------------------------------
 * Replaces the torchvision transform chain of osmosis_sampling.py:46-49 for one decoded image already in device memory:
 * ToTensor -> Resize(size) [bilinear, antialias, short side -> size] -> CenterCrop([size, size]) -> Normalize(0.5, 0.5), and
 * (degamma != 0) the de-gamma of osmosis_sampling.py:170-175, y <- 2 (0.5 (y + 1))^2.2 - 1.
 *   src u8 [H][W][channels] with row_pitch_bytes between rows, channels 3 (RGB) or 1 (grey, replicated as .convert("RGB") does);
 *   scratch >= 3 * H * size floats; out f32 [3][size][size] in [-1, 1].                                              */
int osm_preprocess_image(const uint8_t* src, int H, int W, int channels, int row_pitch_bytes, float* scratch, float* out,
                         int size, int degamma, void* stream);
/* the de-gamma alone on n floats (in place allowed) */
int osm_degamma(const float* y, float* out, long n, void* stream);

/* ------------------------------------------------------------ layer-level entry points -----------------
 * Used by the kernel parity tests (tests/test_kernels_gpu.py); NHWC fp32 with an explicit pixel stride
 * `ld` (elements).  Not part of the sampling API.                                                     */
int osm_dbg_conv(int conv_mode, const float* x, int ldx, const float* w_packed, const float* bias, const float* res,
                 int ldr, int res_mode, float* out, int ldo, int accumulate, int B, int H, int W, int Cin, int Cout,
                 int taps, void* stream);
/* tcgen05 conv whose epilogue also reduces GroupNorm statistics of its output (mode 1: forward mean / rstd; mode 2: the two
 * backward means for the GroupNorm with input gn_x and dy = out).  scratch_partial: >= B * tiles * 256 floats,
 * scratch_coef: B * Cout * 4 floats, stats_out: [B][32][2].  *fused = 0 (nothing run) if the plan cannot fuse. */
int osm_dbg_conv_stats(const float* x, int ldx, const float* w_packed, const float* bias, float* out, int ldo, int B, int H,
                       int W, int Cin, int Cout, int taps, int mode, const float* gn_x, int gn_ldx, const float* gamma,
                       const float* beta, const float* scale_shift, int ld_ss, int silu, const float* fwd_stats,
                       float* scratch_partial, float* scratch_coef, float* stats_out, int* fused, void* stream);
/* the same through the fp16-operand halo kernel (w_packed_f16 from osm_dbg_pack_conv_weight_f16; 3x3, Cin % 64 == 0, Cout % 256 == 0) */
int osm_dbg_conv_stats_f16(const float* x, int ldx, const void* w_packed_f16, const float* bias, float* out, int ldo, int B, int H,
                           int W, int Cin, int Cout, int taps, int mode, const float* gn_x, int gn_ldx, const float* gamma,
                           const float* beta, const float* scale_shift, int ld_ss, int silu, const float* fwd_stats,
                           float* scratch_partial, float* scratch_coef, float* stats_out, int* fused, void* stream);
/* Halo-tile CTA-pair tcgen05 conv (3x3, Cout % 256 == 0, H % 16 == 0, W % 8 == 0).  coef != NULL: [B][Cin] float2 (a, b) and the
 * conv runs on tf32(SiLU?(x a + b)) computed in shared memory from the raw x (GroupNorm + SiLU fused into the operand load,
 * nn.py:17-19 + unet.py:315-335).  tile_n: 128 / 256 = output channels per CTA-pair tile, 0 = chosen by the plan. */
int osm_dbg_conv_halo(const float* x, int ldx, const float* w_packed, const float* bias, const float* coef, int silu, const float* res,
                      int ldr, int res_mode, float* out, int ldo, int B, int H, int W, int Cin, int Cout, int tile_n, void* stream);
/* The fp16-operand variant of the halo kernel (tcgen05 kind::f16, fp32 accumulation): x stays fp32 in memory and is converted
 * (after the optional x a + b / SiLU transform, coef as above) to fp16 in shared memory; w_packed_f16 is the fp16 pack
 * [9][Cin/64][Cout][64] written by osm_dbg_pack_conv_weight_f16.  Cin % 64 == 0. */
int osm_dbg_conv_halo16(const float* x, int ldx, const void* w_packed_f16, const float* bias, const float* coef, int silu,
                        const float* res, int ldr, int res_mode, float* out, int ldo, int accumulate, int B, int H, int W, int Cin,
                        int Cout, void* stream);
/* The tile-per-CTA / persistent / CTA-pair tcgen05 kernels on fp16 operands read straight from memory (kind::f16): x_f16 is an NHWC
 * fp16 tensor (ldx in elements), w_packed_f16 the fp16 pack [taps][Cin/64][Cout][64]; 3x3 or 1x1, Cin % 64 == 0, Cout % 32 == 0. */
int osm_dbg_conv_f16(const void* x_f16, int ldx, const void* w_packed_f16, const float* bias, const float* res, int ldr, int res_mode,
                     float* out, int ldo, int accumulate, int B, int H, int W, int Cin, int Cout, int taps, void* stream);
int osm_dbg_pack_conv_weight_f16(const float* w_oihw, void* w_fwd, void* w_dgrad, int Cout, int Cin, int Cout_p, int Cin_p, int taps,
                                 void* stream);
int osm_dbg_pack_conv_weight(const float* w_oihw, float* w_fwd, float* w_dgrad, int Cout, int Cin, int Cout_p, int Cin_p,
                             int taps, int round_tf32, void* stream);
int osm_dbg_gn_forward(const float* x, int ldx, const float* gamma, const float* beta, const float* scale_shift,
                       int ld_ss, int silu, int resample, float* stats, float* y, int B, int H, int W, int C, void* stream);
int osm_dbg_gn_backward(const float* x, int ldx, const float* gamma, const float* beta, const float* scale_shift,
                        int ld_ss, int silu, int resample, const float* stats, const float* dy, const float* addend,
                        int ld_add, int add_mode, float* dx, int ld_dx, int accumulate, int B, int H, int W, int C,
                        void* stream);
/* the same two with the output written as fp16 elements (the operand of an fp16 tensor-core conv; backward: accumulate must be 0) */
int osm_dbg_gn_forward_f16(const float* x, int ldx, const float* gamma, const float* beta, const float* scale_shift,
                           int ld_ss, int silu, int resample, float* stats, void* y_f16, int B, int H, int W, int C, void* stream);
int osm_dbg_gn_backward_f16(const float* x, int ldx, const float* gamma, const float* beta, const float* scale_shift,
                            int ld_ss, int silu, int resample, const float* stats, const float* dy, const float* addend,
                            int ld_add, int add_mode, void* dx_f16, int ld_dx, int accumulate, int B, int H, int W, int C,
                            void* stream);
int osm_dbg_attention(const float* qkv, float* out, float* scratch_P, int B, int L, int C, int heads, void* stream);
int osm_dbg_attention_bwd(const float* qkv, const float* g_out, float* g_qkv, float* scratch_P, float* scratch_D, int B,
                          int L, int C, int heads, void* stream);
/* Fused tcgen05 flash attention (product mode).  qkvT [B,3C,L], lse/Dv [B,heads,L], g_outT [B,C,L] are caller scratch; the
 * backward expects qkvT/out/lse as left by the forward on the same qkv. */
int osm_dbg_attention_flash(const float* qkv, float* qkvT, float* out, float* lse, int B, int L, int C, int heads, void* stream);
int osm_dbg_attention_flash_bwd(const float* qkv, float* qkvT, float* out, float* lse, float* Dv, const float* g_out,
                                float* g_outT, float* g_qkv, int B, int L, int C, int heads, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSMOSIS_B200_H_ */
