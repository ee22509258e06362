// Sampler-side kernels: posterior mean/variance (+VJP), guidance update, unguided DDPM update,
// measurement operators and the fused phi-optimisation / guidance-gradient loop.
//
// All tensors here are NCHW planes [B,C,HW] fp32 (the layout at the reference boundary), so every
// access is pixel-contiguous and coalesced.  These kernels are HBM/L2-bound; arithmetic is written
// with explicit __fmul_rn/__fadd_rn where the reference rounds op-by-op so that results track the
// fp32 oracle to the last bit where possible (no FMA contraction across reference ops).
#include "common.cuh"

namespace osm {

// ------------------------------------------------------------------------------------------------
// posterior forward:  x0 = c1*x - c2*eps ; mean = m1*x0 + m2*x ; logvar = frac*maxlog + (1-frac)*minlog
// (posterior_mean_variance.py:127-136, 246-258)
// ------------------------------------------------------------------------------------------------
__global__ void posterior_fwd_kernel(const float* __restrict__ coef, const int32_t* __restrict__ t_idx,
                                     const float* __restrict__ x, const float* __restrict__ mo, float* __restrict__ x0,
                                     float* __restrict__ mean, float* __restrict__ logvar, int C, int HW, int flags) {
  const int b = blockIdx.y;
  const float* cf = coef + OSM_COEF_COLS * (size_t)t_idx[b];
  const float c1 = cf[0], c2 = cf[1], m1 = cf[2], m2 = cf[3], maxlog = cf[4], minlog = cf[5];
  const float fsmall = cf[8], flarge = cf[9], r1 = cf[10], r2 = cf[11];
  const bool clip = (flags & OSM_POST_CLIP) != 0;
  const int mean_kind = flags & OSM_POST_MEAN_MASK, var_kind = flags & OSM_POST_VAR_MASK;
  const size_t n4 = (size_t)C * HW / 4;
  const float4* xv = reinterpret_cast<const float4*>(x + (size_t)b * C * HW);
  const float4* ev = reinterpret_cast<const float4*>(mo + (size_t)b * 2 * C * HW);
  const float4* vv = reinterpret_cast<const float4*>(mo + (size_t)b * 2 * C * HW + (size_t)C * HW);
  float4* x0v = reinterpret_cast<float4*>(x0 + (size_t)b * C * HW);
  float4* mv = reinterpret_cast<float4*>(mean + (size_t)b * C * HW);
  float4* lv = reinterpret_cast<float4*>(logvar + (size_t)b * C * HW);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 X = xv[i], E = ev[i], V = vv[i], X0, M, L;
    // mean processors (posterior_mean_variance.py): epsilon :104-136, start_x :76-101, previous_x :54-73
    // variance processors: learned_range :227-258, learned :217-224, fixed_small :173-191, fixed_large :194-214
#define OSM_POST1(f)                                                                                         \
  {                                                                                                          \
    float p0;                                                                                                \
    if (mean_kind == OSM_POST_MEAN_STARTX) p0 = E.f;                                                         \
    else if (mean_kind == OSM_POST_MEAN_PREVX) p0 = __fsub_rn(__fmul_rn(r1, E.f), __fmul_rn(r2, X.f));       \
    else p0 = __fsub_rn(__fmul_rn(c1, X.f), __fmul_rn(c2, E.f));                                             \
    if (clip) p0 = fminf(fmaxf(p0, -1.0f), 1.0f); /* process_xstart: clamp(-1, 1) */                         \
    X0.f = p0;                                                                                               \
    M.f = (mean_kind == OSM_POST_MEAN_PREVX) ? E.f : __fadd_rn(__fmul_rn(m1, p0), __fmul_rn(m2, X.f));       \
    if (var_kind == OSM_POST_VAR_LEARNED) L.f = V.f;                                                         \
    else if (var_kind == OSM_POST_VAR_FIXED_SMALL) L.f = fsmall;                                             \
    else if (var_kind == OSM_POST_VAR_FIXED_LARGE) L.f = flarge;                                             \
    else {                                                                                                   \
      float fr = __fdiv_rn(__fadd_rn(V.f, 1.0f), 2.0f);                                                      \
      L.f = __fadd_rn(__fmul_rn(fr, maxlog), __fmul_rn(__fsub_rn(1.0f, fr), minlog));                        \
    }                                                                                                        \
  }
    OSM_POST1(x) OSM_POST1(y) OSM_POST1(z) OSM_POST1(w)
#undef OSM_POST1
    x0v[i] = X0; mv[i] = M; lv[i] = L;
  }
}

int posterior_fwd_launch(const float* coef, const int32_t* t_idx, const float* x, const float* mo, float* x0, float* mean,
                         float* logvar, int B, int C, int HW, int flags, cudaStream_t s) {
  if ((C * HW) % 4) return fail(OSM_ERR_INVALID, "posterior_fwd: C*HW must be a multiple of 4");
  int blocks = (int)(((size_t)C * HW / 4 + 255) / 256);
  if (blocks > 1184) blocks = 1184;  // 8 x 148 SMs, grid-stride
  posterior_fwd_kernel<<<dim3(blocks, B), 256, 0, s>>>(coef, t_idx, x, mo, x0, mean, logvar, C, HW, flags);
  OSM_LAUNCH_CHECK("posterior_fwd_kernel");
  return OSM_OK;
}

// VJP.  epsilon:    G = mask (g_x0 + m1 g_mean);  g_eps = -c2 G;  g_x = c1 G + m2 g_mean
//       start_x:    G = mask (g_x0 + m1 g_mean);  g_mo  = G;      g_x = m2 g_mean
//       previous_x: G = mask g_x0;                g_mo  = g_mean + r1 G;  g_x = -r2 G
// mask = 1 where the unclamped x0 lies in [-1, 1] (clip_denoised), else everywhere.
// g_v = 0.5 (maxlog - minlog) g_logvar (learned_range), g_logvar (learned), 0 (fixed_*).
__global__ void posterior_vjp_kernel(const float* __restrict__ coef, const int32_t* __restrict__ t_idx,
                                     const float* __restrict__ g_x0, const float* __restrict__ g_mean,
                                     const float* __restrict__ g_logvar, float* __restrict__ g_x,
                                     float* __restrict__ g_mo, int C, int HW, const float* __restrict__ x_clip,
                                     const float* __restrict__ mo_clip, int flags) {
  const int b = blockIdx.y;
  const float* cf = coef + OSM_COEF_COLS * (size_t)t_idx[b];
  const float c1 = cf[0], c2 = cf[1], m1 = cf[2], m2 = cf[3], dl = 0.5f * (cf[4] - cf[5]), r1 = cf[10], r2 = cf[11];
  const int mean_kind = flags & OSM_POST_MEAN_MASK, var_kind = flags & OSM_POST_VAR_MASK;
  const size_t n = (size_t)C * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t o = (size_t)b * n + i;
    const float gm = g_mean ? g_mean[o] : 0.f;
    float g0 = g_x0 ? g_x0[o] : 0.f;
    if (mean_kind != OSM_POST_MEAN_PREVX) g0 += m1 * gm;
    if (x_clip) {  // clip_denoised: the clamp passes the gradient only where the unclamped x0 lies in [-1, 1]
      const float xe = x_clip[o], ee = mo_clip[(size_t)b * 2 * n + i];
      float p0;
      if (mean_kind == OSM_POST_MEAN_STARTX) p0 = ee;
      else if (mean_kind == OSM_POST_MEAN_PREVX) p0 = __fsub_rn(__fmul_rn(r1, ee), __fmul_rn(r2, xe));
      else p0 = __fsub_rn(__fmul_rn(c1, xe), __fmul_rn(c2, ee));
      if (!(p0 >= -1.0f && p0 <= 1.0f)) g0 = 0.f;
    }
    float gx, ge;
    if (mean_kind == OSM_POST_MEAN_STARTX) { ge = g0; gx = m2 * gm; }
    else if (mean_kind == OSM_POST_MEAN_PREVX) { ge = gm + r1 * g0; gx = -r2 * g0; }
    else { ge = -c2 * g0; gx = c1 * g0 + m2 * gm; }
    g_x[o] = gx;
    g_mo[(size_t)b * 2 * n + i] = ge;
    float gv = 0.f;
    if (g_logvar) {
      if (var_kind == OSM_POST_VAR_LEARNED_RANGE) gv = dl * g_logvar[o];
      else if (var_kind == OSM_POST_VAR_LEARNED) gv = g_logvar[o];
    }
    g_mo[(size_t)b * 2 * n + n + i] = gv;
  }
}

int posterior_vjp_launch(const float* coef, const int32_t* t_idx, const float* g_x0, const float* g_mean, const float* g_logvar,
                         float* g_x, float* g_mo, int B, int C, int HW, const float* x_clip, const float* mo_clip, int flags,
                         cudaStream_t s) {
  if ((x_clip == nullptr) != (mo_clip == nullptr)) return fail(OSM_ERR_INVALID, "posterior_vjp: clip needs both x and model_out");
  int blocks = (int)(((size_t)C * HW + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  posterior_vjp_kernel<<<dim3(blocks, B), 256, 0, s>>>(coef, t_idx, g_x0, g_mean, g_logvar, g_x, g_mo, C, HW, x_clip, mo_clip, flags);
  OSM_LAUNCH_CHECK("posterior_vjp_kernel");
  return OSM_OK;
}

// ------------------------------------------------------------------------------------------------
// x_out = mean - scale_c * clamp(g, +-clip) + [t != 0] * exp(0.5*logvar) * noise
// (condition_methods.py:211-223, gaussian_diffusion.py:266-268)
// ------------------------------------------------------------------------------------------------
__global__ void sampler_update_kernel(const float* __restrict__ mean, const float* __restrict__ g_a,
                                      const float* __restrict__ g_b, const float* __restrict__ scale4, float clip,
                                      const float* __restrict__ logvar, const float* __restrict__ noise,
                                      const int32_t* __restrict__ t_idx, float* __restrict__ x_out,
                                      float* __restrict__ grad_out, int C, int HW, int noise_first) {
  const int b = blockIdx.y;
  const bool add_noise = t_idx[b] != 0;
  const size_t n = (size_t)C * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t o = (size_t)b * n + i;
    const int c = (int)(i / HW);
    float g = g_a[o];
    if (g_b) g = __fadd_rn(g, g_b[o]);
    if (grad_out) grad_out[o] = g;
    float gc = g;
    if (clip >= 0.f) gc = fminf(fmaxf(g, -clip), clip);
    float v;
    if (noise_first) {  // p_sample adds the noise, then the `ps` conditioning subtracts the gradient (:499-501, :248)
      v = mean[o];
      if (add_noise) v = __fadd_rn(v, __fmul_rn(expf(__fmul_rn(0.5f, logvar[o])), noise[o]));
      v = __fsub_rn(v, __fmul_rn(gc, scale4[c]));
    } else {
      v = __fsub_rn(mean[o], __fmul_rn(scale4[c], gc));
      if (add_noise) v = __fadd_rn(v, __fmul_rn(expf(__fmul_rn(0.5f, logvar[o])), noise[o]));
    }
    x_out[o] = v;
  }
}

int sampler_update_launch(const float* mean, const float* g_a, const float* g_b, const float* scale4, float clip,
                          const float* logvar, const float* noise, const int32_t* t_idx, float* x_out, float* grad_out, int B,
                          int C, int HW, int noise_first, cudaStream_t s) {
  int blocks = (int)(((size_t)C * HW + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  sampler_update_kernel<<<dim3(blocks, B), 256, 0, s>>>(mean, g_a, g_b, scale4, clip, logvar, noise, t_idx, x_out, grad_out,
                                                        C, HW, noise_first);
  OSM_LAUNCH_CHECK("sampler_update_kernel");
  return OSM_OK;
}

// x = c_x * (x - c_eps * eps) + c_z * z      (osmosis_utils/diffusion.py:122)
__global__ void ddpm_uncond_kernel(float* __restrict__ x, const float* __restrict__ mo, const float* __restrict__ z, float c_x,
                                   float c_eps, float c_z, int C, int Cmo, int HW) {
  const int b = blockIdx.y;
  const size_t n = (size_t)C * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t o = (size_t)b * n + i;
    float e = mo[(size_t)b * Cmo * HW + i];
    float v = __fmul_rn(c_x, __fsub_rn(x[o], __fmul_rn(c_eps, e)));
    x[o] = __fadd_rn(v, __fmul_rn(c_z, z[o]));
  }
}

int ddpm_uncond_launch(float* x, const float* mo, const float* z, float c_x, float c_eps, float c_z, int B, int C, int Cmo,
                       int HW, cudaStream_t s) {
  int blocks = (int)(((size_t)C * HW + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  ddpm_uncond_kernel<<<dim3(blocks, B), 256, 0, s>>>(x, mo, z, c_x, c_eps, c_z, C, Cmo, HW);
  OSM_LAUNCH_CHECK("ddpm_uncond_kernel");
  return OSM_OK;
}

// ------------------------------------------------------------------------------------------------
// DDIM.p_sample (gaussian_diffusion.py:506-535), op by op in fp32 like the reference's tensor expression:
//   eps = (c1 x - x0) / c2 ;  sigma = eta sqrt((1 - abar_prev) / (1 - abar)) sqrt(1 - abar / abar_prev)
//   mean = x0 sqrt(abar_prev) + sqrt(1 - abar_prev - sigma^2) eps ;  [t != 0] mean += sigma z
// Writes the result into `mean_out` and log(sigma^2) is not needed: the guidance update runs with add_noise off.
// ------------------------------------------------------------------------------------------------
__global__ void ddim_sample_kernel(const float* __restrict__ coef, const int32_t* __restrict__ t_idx, const float* __restrict__ x,
                                   const float* __restrict__ x0, const float* __restrict__ noise, float eta,
                                   float* __restrict__ out, int C, int HW) {
  const int b = blockIdx.y;
  const float* cf = coef + OSM_COEF_COLS * (size_t)t_idx[b];
  const float c1 = cf[0], c2 = cf[1], ab = cf[6], abp = cf[7];
  const float sigma = __fmul_rn(__fmul_rn(eta, sqrtf(__fdiv_rn(__fsub_rn(1.0f, abp), __fsub_rn(1.0f, ab)))),
                                sqrtf(__fsub_rn(1.0f, __fdiv_rn(ab, abp))));
  const float sx0 = sqrtf(abp);
  const float se = sqrtf(__fsub_rn(__fsub_rn(1.0f, abp), __fmul_rn(sigma, sigma)));
  const bool add_noise = t_idx[b] != 0;
  const size_t n = (size_t)C * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t o = (size_t)b * n + i;
    const float p0 = x0[o];
    const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(c1, x[o]), p0), c2);
    float v = __fadd_rn(__fmul_rn(p0, sx0), __fmul_rn(se, eps));
    if (add_noise) v = __fadd_rn(v, __fmul_rn(sigma, noise[o]));
    out[o] = v;
  }
}

int ddim_sample_launch(const float* coef, const int32_t* t_idx, const float* x, const float* x0, const float* noise, float eta,
                       float* out, int B, int C, int HW, cudaStream_t s) {
  int blocks = (int)(((size_t)C * HW + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  ddim_sample_kernel<<<dim3(blocks, B), 256, 0, s>>>(coef, t_idx, x, x0, noise, eta, out, C, HW);
  OSM_LAUNCH_CHECK("ddim_sample_kernel");
  return OSM_OK;
}

// ------------------------------------------------------------------------------------------------
// `ps` conditioning with the identity (rgb_guidance) operator (condition_methods.py:36-40, 234-251):
//   L_b = || y_b - x0_b[:3] ||_2 ;  g_x0[b, c<3] = -(y - x0) / L_b ;  g_x0[b, 3:] = 0
// One CTA per image (the 3 planes are L2-resident between the two passes); fp64 fixed-order reduction.
// ------------------------------------------------------------------------------------------------
constexpr int PS_THREADS = 1024;
__global__ void __launch_bounds__(PS_THREADS)
ps_guidance_kernel(const float* __restrict__ x0, const float* __restrict__ y, float* __restrict__ g_x0, float* __restrict__ losses,
                   int C, int HW) {
  __shared__ double red[33];
  const int b = blockIdx.x;
  const float* xb = x0 + (size_t)b * C * HW;
  const float* yb = y + (size_t)b * 3 * HW;
  float* gb = g_x0 + (size_t)b * C * HW;
  const int n3 = 3 * HW;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n3; i += PS_THREADS) {
    const float d = __fsub_rn(yb[i], xb[i]);
    acc += (double)d * (double)d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < PS_THREADS / 32; ++w) t += red[w];
    red[32] = t;
  }
  __syncthreads();
  const float L = (float)sqrt(red[32]);
  const float invL = 1.0f / L;
  for (int i = threadIdx.x; i < n3; i += PS_THREADS) gb[i] = -__fsub_rn(yb[i], xb[i]) * invL;
  for (int i = n3 + threadIdx.x; i < C * HW; i += PS_THREADS) gb[i] = 0.f;
  if (threadIdx.x == 0) losses[b] = L;
}

int ps_guidance_launch(const float* x0, const float* y, float* g_x0, float* losses, int B, int C, int HW, cudaStream_t s) {
  if (C < 3) return fail(OSM_ERR_INVALID, "ps guidance: needs at least 3 channels");
  ps_guidance_kernel<<<B, PS_THREADS, 0, s>>>(x0, y, g_x0, losses, C, HW);
  OSM_LAUNCH_CHECK("ps_guidance_kernel");
  return OSM_OK;
}

// ------------------------------------------------------------------------------------------------
// measurement operator  uw_c = J_c e^{-phi_a,c d} + phi_inf,c (1 - e^{-phi_b,c d})
// ------------------------------------------------------------------------------------------------
struct DepthFn {
  int kind;
  float v0, v1, v2;
};

__device__ __forceinline__ float depth_convert(const DepthFn& f, float d) {
  if (f.kind == OSM_DEPTH_GAMMA) {
    float base = __fmul_rn(__fadd_rn(d, f.v0), f.v1);
    return (f.v2 == 1.0f) ? base : powf(base, f.v2);
  }
  if (f.kind == OSM_DEPTH_MOVE) return __fadd_rn(d, f.v0);
  return __fmul_rn(0.5f, __fadd_rn(d, 1.0f));
}
// d(depth_convert)/d(d)
__device__ __forceinline__ float depth_slope(const DepthFn& f, float d) {
  if (f.kind == OSM_DEPTH_GAMMA) {
    if (f.v2 == 1.0f) return f.v1;
    float base = (d + f.v0) * f.v1;
    return f.v2 * f.v1 * powf(base, f.v2 - 1.0f);
  }
  if (f.kind == OSM_DEPTH_MOVE) return 1.0f;
  return 0.5f;
}

struct Phi9 {
  float a[3], b[3], inf[3];
};
__device__ __forceinline__ Phi9 load_phi(int op_kind, const float* __restrict__ phi) {
  Phi9 p;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    p.a[c] = (op_kind == OSM_OP_HAZE) ? phi[0] : phi[c];
    p.b[c] = (op_kind == OSM_OP_UNDERWATER_REVISED) ? phi[3 + c] : p.a[c];
    p.inf[c] = phi[6 + c];
  }
  return p;
}

__global__ void operator_fwd_kernel(int op_kind, DepthFn df, const float* __restrict__ x, const float* __restrict__ phi,
                                    float* __restrict__ out, int HW) {
  const int b = blockIdx.y;
  const Phi9 p = load_phi(op_kind, phi + 9 * b);
  const float* xb = x + (size_t)b * 4 * HW;
  float* ob = out + (size_t)b * 3 * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const float d = depth_convert(df, xb[3 * (size_t)HW + i]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float J = __fmul_rn(0.5f, __fadd_rn(xb[c * (size_t)HW + i], 1.0f));
      const float ea = expf(__fmul_rn(-p.a[c], d));
      const float eb = expf(__fmul_rn(-p.b[c], d));
      ob[c * (size_t)HW + i] = __fadd_rn(__fmul_rn(J, ea), __fmul_rn(p.inf[c], __fsub_rn(1.0f, eb)));
    }
  }
}

int operator_fwd_launch(int op_kind, int depth_kind, const float* dv, const float* x, const float* phi, float* out, int B,
                        int HW, cudaStream_t s) {
  DepthFn df{depth_kind, dv[0], dv[1], dv[2]};
  int blocks = (HW + 255) / 256;
  if (blocks > 592) blocks = 592;
  operator_fwd_kernel<<<dim3(blocks, B), 256, 0, s>>>(op_kind, df, x, phi, out, HW);
  OSM_LAUNCH_CHECK("operator_fwd_kernel");
  return OSM_OK;
}

// ------------------------------------------------------------------------------------------------
// Fused guidance / phi-optimisation loop.  One thread-block CLUSTER of GUID_CLUSTER CTAs per image (512 threads each, a
// pixel slice per CTA); the image's 7 planes (x0 RGBD + y RGB, 1.8 MB at 256x256) stay L2-resident across the n_iter
// passes.  The 10 per-evaluation sums are block-reduced, exchanged through distributed shared memory and added in rank
// order by every CTA, so all CTAs of an image take identical phi steps (deterministic, no atomics, no host round trip).
//
// Per evaluation:
//   r_c = (y_c - (2 uw_c - 1)) w ;  L = ||r||_2 ;  dL/dphi = (1/L) sum_pix r_c * d r_c/d phi
//   phi <- phi - eta * dL/dphi     (aux losses do not depend on phi)
// Last evaluation additionally writes d(L + aux)/d x0 using phi BEFORE that evaluation's update.
// ------------------------------------------------------------------------------------------------
constexpr int GUID_THREADS = 512;
constexpr int GUID_NRED = 10;
constexpr int GUID_CLUSTER = 8;        // portable cluster size: the default
constexpr int GUID_CLUSTER_MAX = 16;   // non-portable (one cluster per GPC): small batches, where 8 SMs per image leave the GPU idle

__device__ __forceinline__ void guid_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double guid_ld_dsmem_f64(const double* local, uint32_t cta) {
  const uint32_t la = (uint32_t)__cvta_generic_to_shared(local);
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(cta));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of NV doubles per thread; result broadcast to all threads via smem `red` [NV]
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red /*[32*NV + NV]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = warp_sum(v[k]);
    if (lane == 0) red[warp * NV + k] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = (lane < (int)(blockDim.x >> 5)) ? red[lane * NV + k] : 0.0;
      s = warp_sum(s);
      if (lane == 0) red[32 * NV + k] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = red[32 * NV + k];
  __syncthreads();
}

// Sum of the block totals over the CTAs of the cluster (rank order).  `slot` is this CTA's exchange buffer [2][NV]; the
// parity alternates per call so one cluster barrier per call suffices.
template <int NV>
__device__ __forceinline__ void cluster_sum(double (&v)[NV], double* slot, double* tot, int& parity, int csize) {
  if (csize == 1) return;
  __shared__ double peer[GUID_CLUSTER_MAX * GUID_NRED];
  if (threadIdx.x < NV) slot[parity * NV + threadIdx.x] = v[threadIdx.x];
  guid_cluster_sync();
  // one thread per (peer, value): the csize x NV remote loads are in flight together; the sum stays in rank order
  if ((int)threadIdx.x < csize * NV) {
    const int q = threadIdx.x / NV, k = threadIdx.x - q * NV;
    peer[q * NV + k] = guid_ld_dsmem_f64(&slot[parity * NV + k], (uint32_t)q);
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double t = 0.0;
    for (int q = 0; q < csize; ++q) t += peer[q * NV + threadIdx.x];
    tot[threadIdx.x] = t;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = tot[k];
  __syncthreads();
  parity ^= 1;
}

__global__ void __launch_bounds__(GUID_THREADS, 1)
guidance_phi_loop_kernel(osm_guidance_params P, const float* __restrict__ x0, const float* __restrict__ y,
                         float* __restrict__ phi_io, const int32_t* __restrict__ freeze_flag, float* __restrict__ g_x0,
                         float* __restrict__ losses, int HW, int csize) {
  __shared__ double red[32 * GUID_NRED + GUID_NRED];
  __shared__ double xslot[2 * GUID_NRED], xtot[GUID_NRED];
  __shared__ float sphi[9];
  __shared__ float s_m[9], s_v[9], s_step;   // Adam state of this image (every CTA of the cluster keeps an identical copy)
  const int b = blockIdx.x / csize, rank = blockIdx.x % csize;
  int parity = 0;
  const int per = (HW + csize - 1) / csize;
  const int i0 = rank * per, i1 = min(HW, i0 + per);
  const float* xb = x0 + (size_t)b * 4 * HW;
  const float* yb = y + (size_t)b * 3 * HW;
  float* gb = g_x0 + (size_t)b * 4 * HW;
  const DepthFn df{P.depth_kind, P.depth_val[0], P.depth_val[1], P.depth_val[2]};
  const DepthFn wf{P.weight_depth_kind, P.weight_val[0], P.weight_val[1], P.weight_val[2]};
  const bool freeze = freeze_flag[0] != 0;
  const int n_eval = freeze ? 1 : P.n_iter;
  if (threadIdx.x < 9) sphi[threadIdx.x] = phi_io[9 * b + threadIdx.x];
  const bool adam = P.optimizer == OSM_OPT_ADAM;
  if (adam) {
    if (threadIdx.x < 9) { s_m[threadIdx.x] = P.opt_state[19 * b + threadIdx.x]; s_v[threadIdx.x] = P.opt_state[19 * b + 9 + threadIdx.x]; }
    if (threadIdx.x == 0) s_step = P.opt_state[19 * b + 18];
  }

  // phi-independent reductions: channel means (avrg_loss) and the val_loss sum
  double pre[4] = {0, 0, 0, 0};
  for (int i = i0 + threadIdx.x; i < i1; i += GUID_THREADS) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = xb[c * (size_t)HW + i];
      pre[c] += v;
      const float e = fmaxf(fabsf(v) - 0.7f, 0.f);
      pre[3] += (double)(e * e);
    }
  }
  block_sum<4>(pre, red);
  cluster_sum<4>(pre, xslot, xtot, parity, csize);
  float mean_c[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) mean_c[c] = (float)(pre[c] / HW);
  const float avrg_term = fabsf(mean_c[0]) + fabsf(mean_c[1]) + fabsf(mean_c[2]);
  const float val_term = (float)(pre[3] / (3.0 * HW));

  float Lnorm = 0.f;
  for (int it = 0; it < n_eval; ++it) {
    const bool last = (it == n_eval - 1);
    const Phi9 p = load_phi(P.op_kind, sphi);
    double acc[GUID_NRED];
#pragma unroll
    for (int k = 0; k < GUID_NRED; ++k) acc[k] = 0.0;
    for (int i = i0 + threadIdx.x; i < i1; i += GUID_THREADS) {
      const float xd = xb[3 * (size_t)HW + i];
      const float d = depth_convert(df, xd);
      const float w = P.weight_kind ? depth_convert(wf, xd) : 1.0f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float J = __fmul_rn(0.5f, __fadd_rn(xb[c * (size_t)HW + i], 1.0f));
        const float ea = expf(__fmul_rn(-p.a[c], d));
        const float eb = expf(__fmul_rn(-p.b[c], d));
        const float uw = __fadd_rn(__fmul_rn(J, ea), __fmul_rn(p.inf[c], __fsub_rn(1.0f, eb)));
        const float deg = __fsub_rn(__fmul_rn(2.0f, uw), 1.0f);
        const float r = __fmul_rn(__fsub_rn(yb[c * (size_t)HW + i], deg), w);
        acc[0] += (double)r * (double)r;
        // un-normalised dL/duw = -2 w r ;  duw/dphi_a = -d J ea ; duw/dphi_b = phi_inf d eb ; duw/dphi_inf = 1 - eb
        const float gu = -2.0f * w * r;
        acc[1 + c] += (double)(gu * (-d * J * ea));
        acc[4 + c] += (double)(gu * (p.inf[c] * d * eb));
        acc[7 + c] += (double)(gu * (1.0f - eb));
      }
    }
    block_sum<GUID_NRED>(acc, red);
    cluster_sum<GUID_NRED>(acc, xslot, xtot, parity, csize);
    // loss_function (condition_methods.py:127-138): 'norm' L = sqrt(sum r^2), dL/dr = r / L;
    //                                               'mse'  L = mean r^2 over [3,H,W], dL/dr = 2 r / (3 HW)
    float invL;
    if (P.loss_kind == OSM_LOSS_MSE) {
      Lnorm = (float)(acc[0] / (3.0 * HW));
      invL = 2.0f / (3.0f * (float)HW);
    } else {
      Lnorm = (float)sqrt(acc[0]);
      invL = 1.0f / Lnorm;
    }

    if (last) {
      // d total / d x0 at the current phi (before this evaluation's SGD step)
      const float ka = P.gamma_avrg / (float)HW;
      const float kv = P.gamma_val * 2.0f / (3.0f * (float)HW);
      for (int i = i0 + threadIdx.x; i < i1; i += GUID_THREADS) {
        const float xd = xb[3 * (size_t)HW + i];
        const float d = depth_convert(df, xd);
        const float w = P.weight_kind ? depth_convert(wf, xd) : 1.0f;
        float gd = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float xc = xb[c * (size_t)HW + i];
          const float J = __fmul_rn(0.5f, __fadd_rn(xc, 1.0f));
          const float ea = expf(__fmul_rn(-p.a[c], d));
          const float eb = expf(__fmul_rn(-p.b[c], d));
          const float uw = __fadd_rn(__fmul_rn(J, ea), __fmul_rn(p.inf[c], __fsub_rn(1.0f, eb)));
          const float deg = __fsub_rn(__fmul_rn(2.0f, uw), 1.0f);
          const float r = __fmul_rn(__fsub_rn(yb[c * (size_t)HW + i], deg), w);
          const float guw = -2.0f * w * r * invL;
          float g = 0.5f * guw * ea;
          const float sgn_m = (mean_c[c] > 0.f) ? 1.f : ((mean_c[c] < 0.f) ? -1.f : 0.f);
          g += ka * sgn_m;
          const float ex = fmaxf(fabsf(xc) - 0.7f, 0.f);
          g += kv * ex * ((xc > 0.f) ? 1.f : ((xc < 0.f) ? -1.f : 0.f));
          gb[c * (size_t)HW + i] = g;
          gd += guw * (-p.a[c] * J * ea + p.inf[c] * p.b[c] * eb);
        }
        gb[3 * (size_t)HW + i] = gd * depth_slope(df, xd);
      }
    }
    if (!freeze) {
      __syncthreads();
      if (threadIdx.x == 0) {
        // gradients per phi element in get_variable_list() order (tied operators sum the a and b contributions)
        float g[9], lr[9];
        int idx[9], n = 0;
        if (P.op_kind == OSM_OP_UNDERWATER_REVISED) {
          for (int c = 0; c < 3; ++c) { g[n] = (float)(acc[1 + c] * invL); lr[n] = P.eta[0]; idx[n++] = c; }
          for (int c = 0; c < 3; ++c) { g[n] = (float)(acc[4 + c] * invL); lr[n] = P.eta[1]; idx[n++] = 3 + c; }
          for (int c = 0; c < 3; ++c) { g[n] = (float)(acc[7 + c] * invL); lr[n] = P.eta[2]; idx[n++] = 6 + c; }
        } else if (P.op_kind == OSM_OP_UNDERWATER) {
          for (int c = 0; c < 3; ++c) { g[n] = (float)((acc[1 + c] + acc[4 + c]) * invL); lr[n] = P.eta[0]; idx[n++] = c; }
          for (int c = 0; c < 3; ++c) { g[n] = (float)(acc[7 + c] * invL); lr[n] = P.eta[1]; idx[n++] = 6 + c; }
        } else {
          double sab = 0;
          for (int c = 0; c < 3; ++c) sab += acc[1 + c] + acc[4 + c];
          g[n] = (float)(sab * invL); lr[n] = P.eta[0]; idx[n++] = 0;
          for (int c = 0; c < 3; ++c) { g[n] = (float)(acc[7 + c] * invL); lr[n] = P.eta[1]; idx[n++] = 6 + c; }
        }
        if (!adam) {
          // SGD per group (measurements.py:266-303 / torch.optim.SGD, no momentum)
          for (int k = 0; k < n; ++k) sphi[idx[k]] -= lr[k] * g[k];
        } else {
          // torch.optim.Adam (single-tensor path, defaults): step += 1; m = lerp(m, g, 1 - b1); v = b2 v + (1 - b2) g^2;
          // p -= (lr / (1 - b1^step)) * m / (sqrt(v) / sqrt(1 - b2^step) + eps) - bias corrections in double like the host code
          s_step += 1.0f;
          const double bc1 = 1.0 - pow(0.9, (double)s_step), bc2 = 1.0 - pow(0.999, (double)s_step);
          const float bc2_sqrt = (float)sqrt(bc2);
          for (int k = 0; k < n; ++k) {
            const int i = idx[k];
            s_m[i] = __fadd_rn(s_m[i], __fmul_rn(__fsub_rn(g[k], s_m[i]), (float)(1.0 - 0.9)));
            s_v[i] = __fadd_rn(__fmul_rn(s_v[i], 0.999f), __fmul_rn(__fmul_rn(g[k], g[k]), (float)(1.0 - 0.999)));
            const float denom = __fadd_rn(__fdiv_rn(sqrtf(s_v[i]), bc2_sqrt), 1e-8f);
            sphi[i] = __fadd_rn(sphi[i], __fmul_rn((float)(-(double)lr[k] / bc1), __fdiv_rn(s_m[i], denom)));
          }
        }
      }
      __syncthreads();
    }
  }
  if (csize > 1) guid_cluster_sync();  // no CTA may exit while a peer can still read its exchange slots
  if (rank != 0) return;
  if (threadIdx.x < 9) phi_io[9 * b + threadIdx.x] = sphi[threadIdx.x];
  if (adam && !freeze) {
    if (threadIdx.x < 9) { P.opt_state[19 * b + threadIdx.x] = s_m[threadIdx.x]; P.opt_state[19 * b + 9 + threadIdx.x] = s_v[threadIdx.x]; }
    if (threadIdx.x == 0) P.opt_state[19 * b + 18] = s_step;
  }
  if (threadIdx.x == 0) {
    losses[4 * b + 0] = Lnorm;
    losses[4 * b + 1] = avrg_term;
    losses[4 * b + 2] = val_term;
    losses[4 * b + 3] = Lnorm + P.gamma_avrg * avrg_term + P.gamma_val * val_term;
  }
}

int guidance_phi_loop_launch(const osm_guidance_params* p, const float* x0, const float* y, float* phi,
                             const int32_t* freeze_flag, float* g_x0, float* losses, int B, int HW, cudaStream_t s) {
  if (p->op_kind < 0 || p->op_kind > 2) return fail(OSM_ERR_INVALID, "guidance: unknown operator kind");
  if (p->n_iter < 1) return fail(OSM_ERR_INVALID, "guidance: n_iter must be >= 1");
  if (p->loss_kind != OSM_LOSS_NORM && p->loss_kind != OSM_LOSS_MSE) return fail(OSM_ERR_INVALID, "guidance: unknown loss kind");
  if (p->optimizer != OSM_OPT_SGD && p->optimizer != OSM_OPT_ADAM) return fail(OSM_ERR_INVALID, "guidance: unknown optimizer");
  if (p->optimizer == OSM_OPT_ADAM && !p->opt_state) return fail(OSM_ERR_INVALID, "guidance: Adam needs its state buffer");
  int csize = (HW >= GUID_CLUSTER * GUID_THREADS) ? GUID_CLUSTER : 1;
  // Up to 8 images: 16-CTA clusters (a phi iteration is instruction-bound on the image's SMs: 17 us on 8, 20 iterations per step).
  // A GPU holds about one such cluster per GPC, so larger batches keep 8 (18 clusters at once).  OSM_GUID_CLUSTER16=0: never.
  static const int allow16 = [] {
    const char* e = getenv("OSM_GUID_CLUSTER16");
    if (e && atoi(e) == 0) return 0;
    if (cudaFuncSetAttribute(guidance_phi_loop_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t q{};
    q.gridDim = dim3(GUID_CLUSTER_MAX); q.blockDim = dim3(GUID_THREADS);
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension;
    a[0].val.clusterDim.x = GUID_CLUSTER_MAX; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    q.attrs = a; q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, guidance_phi_loop_kernel, &q) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
  }();
  if (csize == GUID_CLUSTER && HW >= GUID_CLUSTER_MAX * GUID_THREADS && B <= allow16 && B <= 8) csize = GUID_CLUSTER_MAX;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(B * csize));
  cfg.blockDim = dim3(GUID_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  OSM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, guidance_phi_loop_kernel, *p, x0, y, phi, freeze_flag, g_x0, losses, HW, csize));
  return OSM_OK;
}

}  // namespace osm
