// Shared declarations for the sm_100a kernels and the native UNet engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/osmosis_b200.h"

namespace osm {

// ---- error plumbing (thread-local message; C ABI returns codes, never throws) ----
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define OSM_CUDA_CHECK(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return osm::cuda_fail(_e, #expr);   \
  } while (0)
#define OSM_LAUNCH_CHECK(name)                                 \
  do {                                                         \
    cudaError_t _e = cudaGetLastError();                       \
    if (_e != cudaSuccess) return osm::cuda_fail(_e, name);    \
  } while (0)

// All kernels ask for the same (maximum-shared) L1/shared carveout as the 197 KB tensor-core conv kernel, so the SMs are
// not re-partitioned every time a small-shared-memory kernel follows a conv (hundreds of alternations per step).
template <typename K>
inline void prefer_max_smem_once(K kernel, bool* done) {
  if (!*done) {
    cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributePreferredSharedMemoryCarveout,
                         (int)cudaSharedmemCarveoutMaxShared);
    *done = true;
  }
}
#define OSM_PREFER_SMEM(kernel)                        \
  do {                                                 \
    static bool _osm_done = false;                     \
    osm::prefer_max_smem_once(kernel, &_osm_done);     \
  } while (0)

// ---- programmatic dependent launch (PDL) ----
// Every kernel of the UNet programs is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with
//     pdl_launch_dependents();  <per-CTA prologue: barriers, TMEM allocation, constants>  pdl_wait();
// so the NEXT kernel's CTAs are scheduled (and run their prologue) while this one drains, instead of paying a full launch
// + drain gap at each of the ~800 kernel boundaries of a step.  pdl_wait() blocks until every prerequisite grid has
// completed and its memory is visible, so data dependences are exactly those of plain stream order; nothing before it
// touches global memory.  Without the attribute (the default, see pdl_enabled) both instructions are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define OSM_LAUNCH_PDL(name, kernel, grid, block, smem, stream, ...)                                  \
  do {                                                                                                \
    cudaError_t _e = osm::launch_pdl(kernel, grid, block, smem, stream, __VA_ARGS__);                 \
    if (_e != cudaSuccess) return osm::cuda_fail(_e, name);                                           \
  } while (0)

// NHWC activation view: pixel-major, `ld` floats between consecutive pixels, C valid channels.
struct View {
  float* p = nullptr;
  int C = 0;
  int ld = 0;
  int H = 0, W = 0;
};

enum ResMode { RES_NONE = 0, RES_SAME = 1, RES_AVGPOOL = 2, RES_NEAREST_UP = 3 };
// resample applied between GroupNorm(+SiLU) and the conv that follows it (ResBlock h_upd, unet.py:318-320)
enum Resample { RS_NONE = 0, RS_DOWN = 1, RS_UP = 2 };
// how an addend at another resolution is folded into a gradient: same res / avg-pool backward / nearest-up backward
enum AddMode { ADD_NONE = 0, ADD_SAME = 1, ADD_FROM_COARSE_QUARTER = 2, ADD_SUM4_FINE = 3 };

// ---------------- conv / GEMM ----------------
struct ConvArgs {
  const float* x; int ldx;        // input NHWC view [B,H,W,Cin_p]
  const float* w;                 // packed [taps][Cin_p/32][Cout_p][32]  (K-block-major, see pack_conv_weight_kernel)
  const float* bias;              // [Cout_p] or null
  const float* res; int ldr; int res_mode;  // residual source (see ResMode); for AVGPOOL the source is [B,2H,2W], for UP [B,H/2,W/2]
  float* out; int ldo;            // output NHWC view [B,H,W,Cout_p]
  int accumulate;                 // out += result
  int B, H, W, Cin_p, Cout_p, taps;
  // fused GroupNorm statistics in the epilogue (see EpiArgs in conv_epilogue.cuh); 0 = off
  int stat_mode, stat_cpg; float* stat_partial; const float* stat_x; int stat_ldx; const void* stat_coef; int stat_silu;
  // halo kernel (3x3, conv_tc_halo_2sm_kernel): halo = 0 off, 1 = on (pair-tile width chosen by the plan), 128 / 256 = on with that
  // tile width.  xf_coef != null: the A operand is tf32(SiLU?(x a + b)) computed in shared memory from the
  // raw input x, with (a, b) = xf_coef[b][ci] (float2 per (image, input channel), as written by gn_coef_fwd_kernel).
  int halo; const void* xf_coef; int xf_silu;
  // fp16 operands (conv_tc_halo16_2sm_kernel; needs halo, Cin_p % 64 == 0): `w` then points to the fp16 pack
  // [taps][Cin_p/64][Cout_p][64] (pack_conv_weight_f16_launch) and x is converted (after the optional transform) in shared memory
  int f16;
  // cluster split-K (conv_tc_kernel): scratch for the fp32 partial tiles, one per CTA of the launch (the engine passes a slice of its
  // own workspace, so two engines on two streams never share it; null: a process-wide scratch the debug entry points use)
  float* splitk_ws; size_t splitk_ws_bytes;
};
bool conv_tc_halo_ok(int B, int H, int W, int Cin_p, int Cout_p, int taps);   // shapes the halo kernel takes
bool conv_tc_halo16_ok(int B, int H, int W, int Cin_p, int Cout_p, int taps); // shapes its fp16-operand variant takes
int conv_check(const ConvArgs& a);
int conv_simt_launch(const ConvArgs& a, cudaStream_t s);

// tcgen05 path: the plan owns the TMA descriptors (built once per bound shape)
struct ConvTcPlan {
  alignas(64) unsigned char tmA[128];
  alignas(64) unsigned char tmB[128];
  ConvArgs a;
  int BN, stages, split;
  int m256;                       // 256-pixel x 256-channel persistent tiles (conv_tc_persist_m256_kernel)
  int two_sm;                     // CTA-pair tcgen05.mma.cta_group::2 kernel (conv_tc_persist_2sm_kernel)
  int halo;                       // halo-tile CTA-pair kernel (conv_tc_halo_2sm_kernel); BN is then 256 or 128
  int f16;                        // fp16-operand halo kernel (conv_tc_halo16_2sm_kernel)
  int tw, th, tn, tiles_w, tiles_h, tiles_b;
  size_t smem_bytes;
};
int conv_tc_plan(const ConvArgs& a, ConvTcPlan* plan);          // host only (driver entry point for tensor maps)
int conv_tc_launch(const ConvTcPlan& plan, cudaStream_t s);
bool conv_tc_stats_capable(const ConvTcPlan& plan);            // can this plan's kernel reduce GroupNorm statistics?
int conv_tc_stat_slots(const ConvTcPlan& plan);                // partial slots per image it writes

// fp16 K-block-major packs for the fp16-operand kernel: Wf[tap][ci/64][co][ci%64], Wd[tap'][co/64][ci][co%64] (RN, saturating)
int pack_conv_weight_f16_launch(const float* w_oihw, void* w_fwd, void* w_dgrad, int Cout, int Cin, int Cout_p, int Cin_p, int taps,
                                cudaStream_t s);
int pack_conv_weight_launch(const float* w_oihw, float* w_fwd, float* w_dgrad, int Cout, int Cin, int Cout_p, int Cin_p,
                            int taps, int round_tf32, cudaStream_t s);

// ---------------- GroupNorm (32 groups, eps 1e-5) ----------------
struct GnArgs {
  const float* x; int ldx;        // input view [B,H,W,C]
  const float* gamma; const float* beta;
  const float* scale_shift; int ld_ss;   // [B][ld_ss]: scale at [c], shift at [C + c]; null = no modulation
  int silu; int resample;         // Resample applied AFTER the activation
  float* stats;                   // [B][32][2] (mean, rstd)
  double* partial; unsigned int* counter;  // scratch: [B][chunks][32][2] doubles, [B] counters (zero-initialised, self-resetting)
  int B, H, W, C;
  int round_tf32;                 // round the written activation to TF32 (RN) - it only feeds tensor-core convs
  int out_f16;                    // y holds fp16 elements (same element indices): the operand of an fp16 tensor-core conv
};
int gn_stats_launch(const GnArgs& a, cudaStream_t s);
int gn_apply_launch(const GnArgs& a, float* y /*dense [B,H',W',C]*/, cudaStream_t s);
struct GnBwdArgs {
  GnArgs f;                       // forward description (x, stats, affine, modulation, silu, resample)
  const float* dy;                // dense [B,H',W',C] gradient w.r.t. the (resampled) activation output
  const float* addend; int ld_add; int add_mode;   // extra gradient added to dx (skip path), see AddMode
  float* dx; int ld_dx; int accumulate;
  float* bstats;                  // [B][32][2] scratch (m1, m2)
  int dx_f16;                     // dx holds fp16 elements (same element indices; needs accumulate == 0): operand of an fp16 dgrad conv
};
int gn_bwd_launch(const GnBwdArgs& a, cudaStream_t s);  // reduce + apply (2 kernels)
int gn_chunks(int H, int W, int C);                     // number of partial chunks per image for gn_stats
// small tensors (<= 32768 elements per (image, group); a 2x resample only in the register-cached form): statistics + apply / both
// backward passes in ONE launch
bool gn_small_capable(const GnArgs& a);
int gn_small_fwd_launch(const GnArgs& a, float* y, cudaStream_t s);
int gn_small_bwd_launch(const GnBwdArgs& a, cudaStream_t s);
int gn_bwd_apply_launch(const GnBwdArgs& a, cudaStream_t s);   // apply only: bstats already hold the two means
// Fused-statistics helpers: fold the conv epilogue's partials into [B][32][2]; per-channel coefficients for mode 2.
//   mode 1: out = (mean, rstd)      mode 2: out = (mean d, mean d xhat) given the forward stats
// coef_gn / coef != null (mode 1): the same launch also writes the forward operand-transform coefficients (gn_coef_fwd) of that GroupNorm
int gn_fused_finalize_launch(const float* partial, int slots_per_image, const float* fwd_stats, float* out, int B, int HW, int C,
                             int mode, cudaStream_t s, const GnArgs* coef_gn = nullptr, float* coef = nullptr);
// one launch for every backward-statistics coefficient set of an input-VJP (device table of descriptors)
struct GnCoefDesc {
  const float* stats; const float* gamma; const float* beta; const float* ss;
  float4* coef;
  int ld_ss, C;
  int pad[4];
};
int gn_coef_batch_launch(const GnCoefDesc* table, int n, int B, cudaStream_t s);
// (mean, rstd) of a concat [a | b] of two equal channel halves from the halves' own (mean, rstd) [B][32][2]
int gn_combine_stats_launch(const float* stats_a, const float* stats_b, float* out, int B, cudaStream_t s);
int gn_coef_launch(const GnArgs& a, float* coef /*[B][C][4]*/, cudaStream_t s);
// forward operand-transform coefficients of GroupNorm(+modulation) `a`: coef[b][c] = (A, Bc) with pre-activation = x A + Bc
int gn_coef_fwd_launch(const GnArgs& a, float* coef /*[B][C][2]*/, cudaStream_t s);

// ---------------- attention (QKVAttentionLegacy, fp32) ----------------
// qkv [B,L,3C] token-major, head h owns channels [h*3*ch, (h+1)*3*ch) as (q,k,v); out [B,L,C]
int attention_fwd_launch(const float* qkv, float* out, float* P /*[B*heads,L,L]*/, int B, int L, int C, int heads, cudaStream_t s);
int attention_bwd_launch(const float* qkv, const float* g_out, float* g_qkv, float* P, float* D, int B, int L, int C, int heads,
                         cudaStream_t s);
int attention_launches(int which);
// L = 64 tokens with 64 channels per head (the 8x8 level): attention_fwd/bwd_launch are ONE fused fp32 launch each (P / D unused)
bool attention_small_ok(int L, int C, int heads);

// Fused tcgen05 flash attention (attention_flash.cu).  The plan owns the TMA descriptors of one attention block's buffers.
struct AttnFlashPlan {
  alignas(64) unsigned char tm[6][128];  // qkv row tile, qkv chunk, qkvT chunk, dO row tile, dO chunk, dOT chunk
  const float* qkv; float* qkvT;         // [B,L,3C] token-major and its channel-major copy [B,3C,L] (written by the forward)
  float* O; float* lse; float* Dv;       // attention output [B,L,C] (kept for the backward), log2-domain LSE / rowsum(dO o O) [B,heads,L]
  const float* dO; float* dOT;           // output gradient [B,L,C] and scratch for its channel-major copy (null: forward-only plan)
  float* g_qkv;                          // [B,L,3C]
  int B, L, C, heads, R;
};
bool attn_flash_supported(int L, int C, int heads);
int attn_flash_plan(AttnFlashPlan* pl, const float* qkv, float* qkvT, float* O, float* lse, float* Dv, const float* dO, float* dOT,
                    float* g_qkv, int B, int L, int C, int heads);
int attn_flash_fwd_launch(const AttnFlashPlan& pl, cudaStream_t s);   // 2 launches: transpose + fused forward
int attn_flash_bwd_launch(const AttnFlashPlan& pl, cudaStream_t s);   // 3 launches: transpose + dQ + dK/dV

// ---------------- small ops ----------------
int timestep_embedding_launch(const float* t, float* out, int B, int dim, cudaStream_t s);
// out[b,n] = bias[n] + sum_k act(in[b,k]) * W[n,k];  act = SiLU if silu_in
int linear_launch(const float* in, int ld_in, const float* W, const float* bias, float* out, int ld_out, int B, int K, int N,
                  int silu_in, cudaStream_t s);
// scale_bits != null: per-image power-of-two scaling (amax of image b * scale in [2^texp, 2^(texp+1))) applied on the way in / undone
// on the way out; amax_bits_launch zeroes bits[B] and fills them with the bit pattern of each image's max |x|
int nchw_to_nhwc_pad_launch(const float* src, float* dst, int B, int C, int HW, int Cp, cudaStream_t s,
                            const unsigned int* scale_bits = nullptr, int texp = 0);
int nhwc_to_nchw_launch(const float* src, int ld, float* dst, int B, int C, int HW, cudaStream_t s,
                        const unsigned int* scale_bits = nullptr, int texp = 0);
int amax_bits_launch(const float* src, unsigned int* bits, int B, size_t n_per_image, cudaStream_t s);

// ---------------- sampler / guidance ----------------
int posterior_fwd_launch(const float* coef, const int32_t* t_idx, const float* x, const float* mo, float* x0, float* mean,
                         float* logvar, int B, int C, int HW, int flags, cudaStream_t s);
int posterior_vjp_launch(const float* coef, const int32_t* t_idx, const float* g_x0, const float* g_mean, const float* g_logvar,
                         float* g_x, float* g_mo, int B, int C, int HW, const float* x_clip, const float* mo_clip, int flags,
                         cudaStream_t s);
int ddim_sample_launch(const float* coef, const int32_t* t_idx, const float* x, const float* x0, const float* noise, float eta,
                       float* out, int B, int C, int HW, cudaStream_t s);
int ps_guidance_launch(const float* x0, const float* y, float* g_x0, float* losses, int B, int C, int HW, cudaStream_t s);
int sampler_update_launch(const float* mean, const float* g_a, const float* g_b, const float* scale4, float clip,
                          const float* logvar, const float* noise, const int32_t* t_idx, float* x_out, float* grad_out, int B,
                          int C, int HW, int noise_first, cudaStream_t s);
int ddpm_uncond_launch(float* x, const float* mo, const float* z, float c_x, float c_eps, float c_z, int B, int C, int Cmo,
                       int HW, cudaStream_t s);
int operator_fwd_launch(int op_kind, int depth_kind, const float* dv, const float* x, const float* phi, float* out, int B,
                        int HW, cudaStream_t s);
int guidance_phi_loop_launch(const osm_guidance_params* p, const float* x0, const float* y, float* phi,
                             const int32_t* freeze_flag, float* g_x0, float* losses, int B, int HW, cudaStream_t s);

// ---------------- post-processing of finished samples (postprocess.cu) ----------------
int postprocess_launch(int op_kind, int depth_kind, const float* dv, const float* x0, const float* y, const float* phi, float* rgb_clip,
                       float* degraded, float* recon, float* norm_out, int B, int HW, cudaStream_t s);
int minmax_quantile_launch(const float* img, float* out, int B, int n, float q_lo, float q_hi, float vmin, float vmax, cudaStream_t s);
int colormap_launch(const float* img, const float* lut, float* out, int B, int n, cudaStream_t s);
int preprocess_launch(const uint8_t* src, int H, int W, int Cs, int pitch, float* tmp, float* out, int S, int degamma, cudaStream_t s);
int degamma_launch(const float* y, float* out, size_t n, cudaStream_t s);

}  // namespace osm
