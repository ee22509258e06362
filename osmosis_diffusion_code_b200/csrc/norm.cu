// GroupNorm32 (+ scale-shift modulation + SiLU + 2x resample) forward and input-gradient kernels on
// NHWC fp32 views.   Reference: nn.py:17-19, 93-100 (GroupNorm32), unet.py:315-335 (how ResBlock
// composes it with SiLU, the scale-shift modulation and AvgPool2d / nearest-up), unet.py:378-384.
//
// HBM-bound: every kernel streams whole pixel rows with 128-bit loads; a thread owns a fixed
// 4-channel slot so per-group partial sums live in registers (fp64 - the FP64 pipe on B200 is far
// from limiting at HBM speed, and it keeps E[x^2]-E[x]^2 accurate).  Reductions are fixed-order
// (no float atomics) so results are bit-reproducible and independent of how a batch is sharded.
#include "common.cuh"

namespace osm {

constexpr int GN_GROUPS = 32;
constexpr float GN_EPS = 1e-5f;
constexpr int GN_MAX_CHUNKS = 256;

static inline int gn_tpb(int C) {
  const int C4 = C / 4;
  if (C4 <= 256) return C4 * (256 / C4);
  return C4;
}

int gn_chunks(int H, int W, int C) {
  const int C4 = C / 4, tpb = gn_tpb(C), ppi = tpb / C4;
  const int HW = H * W;
  int iters = (HW + ppi - 1) / ppi;
  int chunks = (iters + 7) / 8;
  if (chunks > GN_MAX_CHUNKS) chunks = GN_MAX_CHUNKS;
  if (chunks < 1) chunks = 1;
  return chunks;
}

static int gn_check(const GnArgs& a) {
  if (a.C % (4 * GN_GROUPS) != 0 && !(a.C % GN_GROUPS == 0 && (a.C / GN_GROUPS) % 4 == 0))
    return fail(OSM_ERR_INVALID, "GroupNorm: channels per group must be a multiple of 4");
  if (a.C / 4 > 1024) return fail(OSM_ERR_INVALID, "GroupNorm: C too large");
  if (a.ldx % 4) return fail(OSM_ERR_INVALID, "GroupNorm: ld must be a multiple of 4");
  if (a.resample == RS_DOWN && ((a.H | a.W) & 1)) return fail(OSM_ERR_INVALID, "GroupNorm: odd size with downsample");
  return OSM_OK;
}

// ---- fixed-order block reduce of per-thread (s0, s1) into 32 groups, then cross-block finalize ----
// mode 0: forward stats  -> out[b][g] = (mean, rstd)
// mode 1: backward stats -> out[b][g] = (sum0/N, sum1/N)
__device__ __forceinline__ void gn_group_reduce_and_finalize(double s0, double s1, int C4, int chunks, double* partial,
                                                             unsigned int* counter, float* out, double N, int mode) {
  extern __shared__ double sred[];  // [blockDim][2]
  __shared__ int s_last;
  const int tid = threadIdx.x, b = blockIdx.y, chunk = blockIdx.x;
  sred[2 * tid] = s0;
  sred[2 * tid + 1] = s1;
  __syncthreads();
  const int vpg = C4 / GN_GROUPS, ppi = blockDim.x / C4;
  if (tid < GN_GROUPS) {
    double S0 = 0, S1 = 0;
    for (int r = 0; r < ppi; ++r)
      for (int j = 0; j < vpg; ++j) {
        const int idx = r * C4 + tid * vpg + j;
        S0 += sred[2 * idx];
        S1 += sred[2 * idx + 1];
      }
    double* dst = partial + (((size_t)b * chunks + chunk) * GN_GROUPS + tid) * 2;
    __stcg(dst, S0);
    __stcg(dst + 1, S1);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int ticket = atomicAdd(&counter[b], 1u);
    s_last = (ticket == (unsigned int)(chunks - 1));
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    if (tid < GN_GROUPS) {
      double S0 = 0, S1 = 0;
      for (int c = 0; c < chunks; ++c) {
        const double* src = partial + (((size_t)b * chunks + c) * GN_GROUPS + tid) * 2;
        S0 += __ldcg(src);
        S1 += __ldcg(src + 1);
      }
      float* o = out + ((size_t)b * GN_GROUPS + tid) * 2;
      if (mode == 0) {
        const double mean = S0 / N;
        double var = S1 / N - mean * mean;
        if (var < 0) var = 0;
        o[0] = (float)mean;
        o[1] = (float)(1.0 / sqrt(var + (double)GN_EPS));
      } else {
        o[0] = (float)(S0 / N);
        o[1] = (float)(S1 / N);
      }
    }
    if (tid == 0) counter[b] = 0;  // self-reset for the next launch on this stream
  }
}

__global__ void gn_stats_kernel(const float* __restrict__ x, int ldx, int C4, int HW, int pix_chunk, int chunks,
                                double* partial, unsigned int* counter, float* stats) {
  const int tid = threadIdx.x, b = blockIdx.y;
  const int c4 = tid % C4, prow = tid / C4, ppi = blockDim.x / C4;
  const int p0 = blockIdx.x * pix_chunk;
  const int p1 = min(HW, p0 + pix_chunk);
  double s = 0, ss = 0;
  const float* xb = x + (size_t)b * HW * ldx + 4 * c4;
  for (int p = p0 + prow; p < p1; p += ppi) {
    const float4 v = *reinterpret_cast<const float4*>(xb + (size_t)p * ldx);
    s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
    ss += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  gn_group_reduce_and_finalize(s, ss, C4, chunks, partial, counter, stats, (double)HW * (4.0 * C4 / GN_GROUPS), 0);
}

int gn_stats_launch(const GnArgs& a, cudaStream_t s) {
  if (int e = gn_check(a)) return e;
  const int C4 = a.C / 4, tpb = gn_tpb(a.C), ppi = tpb / C4, HW = a.H * a.W;
  const int chunks = gn_chunks(a.H, a.W, a.C);
  int pix_chunk = (HW + chunks - 1) / chunks;
  pix_chunk = (pix_chunk + ppi - 1) / ppi * ppi;
  gn_stats_kernel<<<dim3(chunks, a.B), tpb, tpb * 2 * sizeof(double), s>>>(a.x, a.ldx, C4, HW, pix_chunk, chunks, a.partial,
                                                                           a.counter, a.stats);
  OSM_LAUNCH_CHECK("gn_stats_kernel");
  return OSM_OK;
}

// ---- shared per-vector math ----
struct GnVec {
  float4 xhat;  // normalised input
  float4 v;     // pre-activation (after affine + modulation)
};

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }
__device__ __forceinline__ float silu_grad_f(float v) {
  const float sg = 1.0f / (1.0f + expf(-v));
  return sg * (1.0f + v * (1.0f - sg));
}
__device__ __forceinline__ float round_tf32_f(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

struct GnChan {  // per-(image, 4-channel slot) constants
  float mean, rstd;
  float4 gamma, beta, sc1;  // sc1 = 1 + scale (or 1)
  float4 shift;
};

__device__ __forceinline__ GnChan gn_load_chan(const float* stats, const float* gamma, const float* beta, const float* ss,
                                               int ld_ss, int b, int c4, int C) {
  GnChan k;
  const int cpg = C / GN_GROUPS, g = (4 * c4) / cpg;
  k.mean = stats[((size_t)b * GN_GROUPS + g) * 2];
  k.rstd = stats[((size_t)b * GN_GROUPS + g) * 2 + 1];
  k.gamma = *reinterpret_cast<const float4*>(gamma + 4 * c4);
  k.beta = *reinterpret_cast<const float4*>(beta + 4 * c4);
  if (ss) {
    const float4 sc = *reinterpret_cast<const float4*>(ss + (size_t)b * ld_ss + 4 * c4);
    k.sc1 = make_float4(1.0f + sc.x, 1.0f + sc.y, 1.0f + sc.z, 1.0f + sc.w);
    k.shift = *reinterpret_cast<const float4*>(ss + (size_t)b * ld_ss + C + 4 * c4);
  } else {
    k.sc1 = make_float4(1.f, 1.f, 1.f, 1.f);
    k.shift = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return k;
}

__device__ __forceinline__ GnVec gn_eval(const GnChan& k, const float4 x) {
  GnVec r;
  r.xhat = make_float4((x.x - k.mean) * k.rstd, (x.y - k.mean) * k.rstd, (x.z - k.mean) * k.rstd, (x.w - k.mean) * k.rstd);
  r.v.x = (r.xhat.x * k.gamma.x + k.beta.x) * k.sc1.x + k.shift.x;
  r.v.y = (r.xhat.y * k.gamma.y + k.beta.y) * k.sc1.y + k.shift.y;
  r.v.z = (r.xhat.z * k.gamma.z + k.beta.z) * k.sc1.z + k.shift.z;
  r.v.w = (r.xhat.w * k.gamma.w + k.beta.w) * k.sc1.w + k.shift.w;
  return r;
}

__device__ __forceinline__ float4 gn_act(const GnChan& k, const float4 x, int silu) {
  float4 v = gn_eval(k, x).v;
  if (silu) v = make_float4(silu_f(v.x), silu_f(v.y), silu_f(v.z), silu_f(v.w));
  return v;
}

// y dense [B,Ho,Wo,C];  resample: none (Ho=H), down (Ho=H/2, average of the 4 activations), up (Ho=2H, nearest)
__global__ void gn_apply_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ ss, int ld_ss,
                                const float* __restrict__ stats, int silu, int resample, int round_tf32, float* __restrict__ y,
                                int B, int H, int W, int C) {
  const int C4 = C / 4;
  const int Ho = resample == RS_DOWN ? H / 2 : (resample == RS_UP ? H * 2 : H);
  const int Wo = resample == RS_DOWN ? W / 2 : (resample == RS_UP ? W * 2 : W);
  const size_t total = (size_t)B * Ho * Wo * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    size_t pix = i / C4;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho);
    const int b = (int)(pix / Ho);
    const GnChan k = gn_load_chan(stats, gamma, beta, ss, ld_ss, b, c4, C);
    const float* xb = x + (size_t)b * H * W * ldx + 4 * c4;
    float4 o;
    if (resample == RS_DOWN) {
      const float4 a0 = gn_act(k, *reinterpret_cast<const float4*>(xb + ((size_t)(2 * ho) * W + 2 * wo) * ldx), silu);
      const float4 a1 = gn_act(k, *reinterpret_cast<const float4*>(xb + ((size_t)(2 * ho) * W + 2 * wo + 1) * ldx), silu);
      const float4 a2 = gn_act(k, *reinterpret_cast<const float4*>(xb + ((size_t)(2 * ho + 1) * W + 2 * wo) * ldx), silu);
      const float4 a3 = gn_act(k, *reinterpret_cast<const float4*>(xb + ((size_t)(2 * ho + 1) * W + 2 * wo + 1) * ldx), silu);
      o = make_float4((a0.x + a1.x + a2.x + a3.x) * 0.25f, (a0.y + a1.y + a2.y + a3.y) * 0.25f,
                      (a0.z + a1.z + a2.z + a3.z) * 0.25f, (a0.w + a1.w + a2.w + a3.w) * 0.25f);
    } else {
      const int h = resample == RS_UP ? ho / 2 : ho, w = resample == RS_UP ? wo / 2 : wo;
      o = gn_act(k, *reinterpret_cast<const float4*>(xb + ((size_t)h * W + w) * ldx), silu);
    }
    if (round_tf32) o = make_float4(round_tf32_f(o.x), round_tf32_f(o.y), round_tf32_f(o.z), round_tf32_f(o.w));
    reinterpret_cast<float4*>(y)[i] = o;
  }
}

int gn_apply_launch(const GnArgs& a, float* y, cudaStream_t s) {
  if (int e = gn_check(a)) return e;
  const int Ho = a.resample == RS_DOWN ? a.H / 2 : (a.resample == RS_UP ? a.H * 2 : a.H);
  const int Wo = a.resample == RS_DOWN ? a.W / 2 : (a.resample == RS_UP ? a.W * 2 : a.W);
  const size_t total = (size_t)a.B * Ho * Wo * (a.C / 4);
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gn_apply_kernel<<<(unsigned)blocks, 256, 0, s>>>(a.x, a.ldx, a.gamma, a.beta, a.scale_shift, a.ld_ss, a.stats, a.silu,
                                                   a.resample, a.round_tf32, y, a.B, a.H, a.W, a.C);
  OSM_LAUNCH_CHECK("gn_apply_kernel");
  return OSM_OK;
}

// ---- backward ----
// gradient arriving at input-resolution pixel (h,w) from the dense dy at the resampled resolution
__device__ __forceinline__ float4 gn_fetch_dy(const float* __restrict__ dy, int resample, int b, int h, int w, int H, int W,
                                              int C, int c4) {
  if (resample == RS_NONE) return *reinterpret_cast<const float4*>(dy + (((size_t)b * H + h) * W + w) * C + 4 * c4);
  if (resample == RS_DOWN) {  // forward averaged 2x2 -> each input gets a quarter of the coarse gradient
    const int Hc = H / 2, Wc = W / 2;
    float4 g = *reinterpret_cast<const float4*>(dy + (((size_t)b * Hc + h / 2) * Wc + w / 2) * C + 4 * c4);
    return make_float4(0.25f * g.x, 0.25f * g.y, 0.25f * g.z, 0.25f * g.w);
  }
  const int Hf = H * 2, Wf = W * 2;  // forward replicated -> sum of the 4 fine gradients
  const float* base = dy + (((size_t)b * Hf + 2 * h) * Wf + 2 * w) * C + 4 * c4;
  const float4 g0 = *reinterpret_cast<const float4*>(base);
  const float4 g1 = *reinterpret_cast<const float4*>(base + C);
  const float4 g2 = *reinterpret_cast<const float4*>(base + (size_t)Wf * C);
  const float4 g3 = *reinterpret_cast<const float4*>(base + (size_t)Wf * C + C);
  return make_float4(g0.x + g1.x + g2.x + g3.x, g0.y + g1.y + g2.y + g3.y, g0.z + g1.z + g2.z + g3.z, g0.w + g1.w + g2.w + g3.w);
}

__device__ __forceinline__ float4 gn_dxhat(const GnChan& k, const GnVec& e, float4 g, int silu) {
  if (silu) {
    g.x *= silu_grad_f(e.v.x); g.y *= silu_grad_f(e.v.y); g.z *= silu_grad_f(e.v.z); g.w *= silu_grad_f(e.v.w);
  }
  return make_float4(g.x * k.sc1.x * k.gamma.x, g.y * k.sc1.y * k.gamma.y, g.z * k.sc1.z * k.gamma.z, g.w * k.sc1.w * k.gamma.w);
}

__global__ void gn_bwd_reduce_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, const float* __restrict__ ss, int ld_ss,
                                     const float* __restrict__ stats, int silu, int resample, const float* __restrict__ dy,
                                     int C4, int H, int W, int pix_chunk, int chunks, double* partial, unsigned int* counter,
                                     float* bstats) {
  const int tid = threadIdx.x, b = blockIdx.y, C = 4 * C4, HW = H * W;
  const int c4 = tid % C4, prow = tid / C4, ppi = blockDim.x / C4;
  const int p0 = blockIdx.x * pix_chunk;
  const int p1 = min(HW, p0 + pix_chunk);
  const GnChan k = gn_load_chan(stats, gamma, beta, ss, ld_ss, b, c4, C);
  const float* xb = x + (size_t)b * HW * ldx + 4 * c4;
  double s0 = 0, s1 = 0;
  for (int p = p0 + prow; p < p1; p += ppi) {
    const GnVec e = gn_eval(k, *reinterpret_cast<const float4*>(xb + (size_t)p * ldx));
    const float4 g = gn_fetch_dy(dy, resample, b, p / W, p % W, H, W, C, c4);
    const float4 d = gn_dxhat(k, e, g, silu);
    s0 += (double)d.x + (double)d.y + (double)d.z + (double)d.w;
    s1 += (double)d.x * e.xhat.x + (double)d.y * e.xhat.y + (double)d.z * e.xhat.z + (double)d.w * e.xhat.w;
  }
  gn_group_reduce_and_finalize(s0, s1, C4, chunks, partial, counter, bstats, (double)HW * (4.0 * C4 / GN_GROUPS), 1);
}

__device__ __forceinline__ float4 gn_fetch_addend(const float* __restrict__ a, int ld, int mode, int b, int h, int w, int H,
                                                  int W, int c4) {
  if (mode == ADD_SAME) return *reinterpret_cast<const float4*>(a + (((size_t)b * H + h) * W + w) * ld + 4 * c4);
  if (mode == ADD_FROM_COARSE_QUARTER) {
    const int Hc = H / 2, Wc = W / 2;
    const float4 g = *reinterpret_cast<const float4*>(a + (((size_t)b * Hc + h / 2) * Wc + w / 2) * ld + 4 * c4);
    return make_float4(0.25f * g.x, 0.25f * g.y, 0.25f * g.z, 0.25f * g.w);
  }
  const int Hf = H * 2, Wf = W * 2;
  const float* base = a + (((size_t)b * Hf + 2 * h) * Wf + 2 * w) * ld + 4 * c4;
  const float4 g0 = *reinterpret_cast<const float4*>(base);
  const float4 g1 = *reinterpret_cast<const float4*>(base + ld);
  const float4 g2 = *reinterpret_cast<const float4*>(base + (size_t)Wf * ld);
  const float4 g3 = *reinterpret_cast<const float4*>(base + (size_t)Wf * ld + ld);
  return make_float4(g0.x + g1.x + g2.x + g3.x, g0.y + g1.y + g2.y + g3.y, g0.z + g1.z + g2.z + g3.z, g0.w + g1.w + g2.w + g3.w);
}

__global__ void gn_bwd_apply_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ ss, int ld_ss,
                                    const float* __restrict__ stats, const float* __restrict__ bstats, int silu, int resample,
                                    const float* __restrict__ dy, const float* __restrict__ addend, int ld_add, int add_mode,
                                    float* __restrict__ dx, int ld_dx, int accumulate, int B, int H, int W, int C) {
  const int C4 = C / 4, cpg = C / GN_GROUPS;
  const size_t total = (size_t)B * H * W * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    size_t pix = i / C4;
    const int w = (int)(pix % W);
    pix /= W;
    const int h = (int)(pix % H);
    const int b = (int)(pix / H);
    const GnChan k = gn_load_chan(stats, gamma, beta, ss, ld_ss, b, c4, C);
    const int g = (4 * c4) / cpg;
    const float m1 = bstats[((size_t)b * GN_GROUPS + g) * 2], m2 = bstats[((size_t)b * GN_GROUPS + g) * 2 + 1];
    const size_t prow = ((size_t)b * H + h) * W + w;
    const GnVec e = gn_eval(k, *reinterpret_cast<const float4*>(x + prow * ldx + 4 * c4));
    const float4 gy = gn_fetch_dy(dy, resample, b, h, w, H, W, C, c4);
    const float4 d = gn_dxhat(k, e, gy, silu);
    float4 o = make_float4(k.rstd * (d.x - m1 - e.xhat.x * m2), k.rstd * (d.y - m1 - e.xhat.y * m2),
                           k.rstd * (d.z - m1 - e.xhat.z * m2), k.rstd * (d.w - m1 - e.xhat.w * m2));
    if (add_mode != ADD_NONE) {
      const float4 a = gn_fetch_addend(addend, ld_add, add_mode, b, h, w, H, W, c4);
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    float4* dst = reinterpret_cast<float4*>(dx + prow * ld_dx + 4 * c4);
    if (accumulate) {
      const float4 p = *dst;
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
    }
    *dst = o;
  }
}

int gn_bwd_launch(const GnBwdArgs& a, cudaStream_t s) {
  const GnArgs& f = a.f;
  if (int e = gn_check(f)) return e;
  if (a.ld_dx % 4 || (a.add_mode != ADD_NONE && a.ld_add % 4)) return fail(OSM_ERR_INVALID, "gn_bwd: ld must be a multiple of 4");
  const int C4 = f.C / 4, tpb = gn_tpb(f.C), ppi = tpb / C4, HW = f.H * f.W;
  const int chunks = gn_chunks(f.H, f.W, f.C);
  int pix_chunk = (HW + chunks - 1) / chunks;
  pix_chunk = (pix_chunk + ppi - 1) / ppi * ppi;
  gn_bwd_reduce_kernel<<<dim3(chunks, f.B), tpb, tpb * 2 * sizeof(double), s>>>(
      f.x, f.ldx, f.gamma, f.beta, f.scale_shift, f.ld_ss, f.stats, f.silu, f.resample, a.dy, C4, f.H, f.W, pix_chunk, chunks,
      f.partial, f.counter, a.bstats);
  OSM_LAUNCH_CHECK("gn_bwd_reduce_kernel");
  const size_t total = (size_t)f.B * HW * C4;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gn_bwd_apply_kernel<<<(unsigned)blocks, 256, 0, s>>>(f.x, f.ldx, f.gamma, f.beta, f.scale_shift, f.ld_ss, f.stats, a.bstats,
                                                       f.silu, f.resample, a.dy, a.addend, a.ld_add, a.add_mode, a.dx, a.ld_dx,
                                                       a.accumulate, f.B, f.H, f.W, f.C);
  OSM_LAUNCH_CHECK("gn_bwd_apply_kernel");
  return OSM_OK;
}

}  // namespace osm
