// GroupNorm32 (+ scale-shift modulation + SiLU + 2x resample) forward and input-gradient kernels on
// NHWC fp32 views.   Reference: nn.py:17-19, 93-100 (GroupNorm32), unet.py:315-335 (how ResBlock
// composes it with SiLU, the scale-shift modulation and AvgPool2d / nearest-up), unet.py:378-384.
//
// HBM-bound.  Common structure: grid = (pixel chunks, images); a thread owns a FIXED 4-channel slot
// (c4 = tid % C4) and walks pixels with stride ppi = blockDim / C4, so
//   * a warp reads/writes whole contiguous pixel rows with 128-bit accesses (coalesced),
//   * all per-channel constants (gamma, beta, 1+scale, shift, the group's mean/rstd) are loaded ONCE per
//     thread and live in registers,
//   * the loop is unrolled 4x with the loads issued first, to keep enough bytes in flight per SM.
// Group reductions accumulate in fp64 per thread (the FP64 pipe on B200 is far from limiting at HBM speed
// and it keeps E[x^2]-E[x]^2 accurate) and are combined in a FIXED order (no float atomics): results are
// bit-reproducible and independent of how a batch is sharded across GPUs.
#include <stdlib.h>

#include "common.cuh"

namespace osm {

constexpr int GN_GROUPS = 32;
constexpr float GN_EPS = 1e-5f;
constexpr int GN_MAX_CHUNKS = 1024;     // pointwise kernels
constexpr int GN_MAX_RED_CHUNKS = 256;  // reduction kernels: the last block folds all chunk partials, so keep them few
// ... but not fewer than the machine needs: 256 blocks of 256 threads on 148 SMs leave 40 SMs with one block and 108 with two
// (ncu: 21 % occupancy, 2.0 TB/s on a 67 MB tensor).  Small batches get 4 blocks per SM in total instead.
static inline int gn_red_max_chunks(int B) { return B >= 3 ? GN_MAX_RED_CHUNKS : 592 / B; }
constexpr int GN_UNROLL = 4;
constexpr int GN_APPLY_UNROLL = 8;   // forward apply: loads in flight per thread
constexpr int GN_RED_UNROLL = 8;

static inline int gn_tpb(int C) {
  const int C4 = C / 4;
  if (C4 <= 256) return C4 * (256 / C4);
  return C4;
}

// number of pixel chunks (= blocks per image) for `npix` pixels walked `ppi` at a time
static inline int chunks_for(int npix, int ppi, int iters_per_thread, int max_chunks = GN_MAX_CHUNKS) {
  const int iters = (npix + ppi - 1) / ppi;
  // small tensors: fewer iterations per thread so that ~128 blocks share the work (latency, not bandwidth, bounds them)
  int cap = iters / 128;
  if (cap < 1) cap = 1;
  if (iters_per_thread > cap) iters_per_thread = cap;
  int chunks = (iters + iters_per_thread - 1) / iters_per_thread;
  if (chunks > max_chunks) chunks = max_chunks;
  return chunks < 1 ? 1 : chunks;
}

int gn_chunks(int H, int W, int C) { return chunks_for(H * W, gn_tpb(C) / (C / 4), 16, GN_MAX_RED_CHUNKS); }

static int gn_check(const GnArgs& a) {
  if (a.C % (4 * GN_GROUPS)) return fail(OSM_ERR_INVALID, "GroupNorm: channels per group must be a multiple of 4");
  if (a.C / 4 > 1024) return fail(OSM_ERR_INVALID, "GroupNorm: C too large");
  if (a.ldx % 4) return fail(OSM_ERR_INVALID, "GroupNorm: ld must be a multiple of 4");
  if (a.resample == RS_DOWN && ((a.H | a.W) & 1)) return fail(OSM_ERR_INVALID, "GroupNorm: odd size with downsample");
  return OSM_OK;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

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
    __threadfence();  // only the 32 writers publish; the block barrier below orders them before thread 0's ticket
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned int ticket = atomicAdd(&counter[b], 1u);
    s_last = (ticket == (unsigned int)(chunks - 1));
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    // 32 groups x J lanes: lane j sums chunks j, j+J, ... then a fixed-order J-way combine (deterministic; J depends on
    // the block size only, i.e. on C, never on the batch or its sharding)
    const int J = blockDim.x >= 256 ? 8 : (blockDim.x >= 128 ? 4 : (blockDim.x >= 64 ? 2 : 1));
    const int g = tid / J, j = tid % J;
    if (tid < GN_GROUPS * J) {
      // fixed order per lane: chunks j, j+J, ...; 8 independent 16-byte loads in flight per round (the partials sit in L2)
      double S0 = 0, S1 = 0;
      const double2* base = reinterpret_cast<const double2*>(partial) + ((size_t)b * chunks) * GN_GROUPS + g;
      int c = j;
      for (; c + 7 * J < chunks; c += 8 * J) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(base + (size_t)(c + u * J) * GN_GROUPS);
#pragma unroll
        for (int u = 0; u < 8; ++u) { S0 += v[u].x; S1 += v[u].y; }
      }
      for (; c < chunks; c += J) {
        const double2 v = __ldcg(base + (size_t)c * GN_GROUPS);
        S0 += v.x;
        S1 += v.y;
      }
      for (int o = J >> 1; o > 0; o >>= 1) {
        S0 += __shfl_down_sync(0xffffffffu, S0, o, J);
        S1 += __shfl_down_sync(0xffffffffu, S1, o, J);
      }
      if (j == 0) {
        float* o = out + ((size_t)b * GN_GROUPS + g) * 2;
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
    }
    if (tid == 0) counter[b] = 0;  // self-reset for the next launch on this stream
  }
}

__global__ void gn_stats_kernel(const float* __restrict__ x, int ldx, int C4, int HW, int pix_chunk, int chunks,
                                double* partial, unsigned int* counter, float* stats) {
  pdl_wait();
  const int tid = threadIdx.x, b = blockIdx.y;
  const int c4 = tid % C4, prow = tid / C4, ppi = blockDim.x / C4;
  const int p0 = blockIdx.x * pix_chunk;
  const int p1 = min(HW, p0 + pix_chunk);
  double s = 0, ss = 0;
  const float* xb = x + (size_t)b * HW * ldx + 4 * c4;
  int p = p0 + prow;
  for (; p + (GN_RED_UNROLL - 1) * ppi < p1; p += GN_RED_UNROLL * ppi) {
    float4 v[GN_RED_UNROLL];
#pragma unroll
    for (int u = 0; u < GN_RED_UNROLL; ++u) v[u] = ldg4(xb + (size_t)(p + u * ppi) * ldx);
#pragma unroll
    for (int u = 0; u < GN_RED_UNROLL; ++u) {
      const float ps = (v[u].x + v[u].y) + (v[u].z + v[u].w);
      const float pq = (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
      s += (double)ps;
      ss += (double)pq;
    }
  }
  for (; p < p1; p += ppi) {
    const float4 v = ldg4(xb + (size_t)p * ldx);
    s += (double)((v.x + v.y) + (v.z + v.w));
    ss += (double)((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
  }
  pdl_launch_dependents();   // main loop done: let the next kernel launch while the reduction tail runs
  gn_group_reduce_and_finalize(s, ss, C4, chunks, partial, counter, stats, (double)HW * (4.0 * C4 / GN_GROUPS), 0);
}

static void chunking(int npix, int C, int iters, int* tpb, int* chunks, int* pix_chunk, int max_chunks = GN_MAX_CHUNKS) {
  const int C4 = C / 4;
  *tpb = gn_tpb(C);
  const int ppi = *tpb / C4;
  *chunks = chunks_for(npix, ppi, iters, max_chunks);
  int pc = (npix + *chunks - 1) / *chunks;
  *pix_chunk = (pc + ppi - 1) / ppi * ppi;
}

int gn_stats_launch(const GnArgs& a, cudaStream_t s) {
  if (int e = gn_check(a)) return e;
  int tpb, chunks, pix_chunk;
  chunking(a.H * a.W, a.C, 16, &tpb, &chunks, &pix_chunk, gn_red_max_chunks(a.B));
  OSM_PREFER_SMEM(gn_stats_kernel);
  OSM_LAUNCH_PDL("gn_stats_kernel", gn_stats_kernel, dim3(chunks, a.B), dim3(tpb), tpb * 2 * sizeof(double), s, a.x, a.ldx, a.C / 4,
                 a.H * a.W, pix_chunk, chunks, a.partial, a.counter, a.stats);
  return OSM_OK;
}

// ---- per-thread channel constants and the pointwise math ----
// sigmoid through ex2.approx / rcp.approx (a few ulp; the SiLU kernels are otherwise issue-bound on expf + IEEE division)
__device__ __forceinline__ float sigmoid_f(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float silu_f(float v) { return v * sigmoid_f(v); }
__device__ __forceinline__ float silu_grad_f(float v) {
  const float sg = sigmoid_f(v);
  return sg * (1.0f + v * (1.0f - sg));
}
__device__ __forceinline__ float round_tf32_f(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

struct GnChan {  // constants of one (image, 4-channel slot)
  float mean, rstd;
  float4 gamma, beta, sc1, shift;  // sc1 = 1 + scale (or 1)
};

__device__ __forceinline__ GnChan gn_load_chan(const float* stats, const float* gamma, const float* beta, const float* ss,
                                               int ld_ss, int b, int c4, int C) {
  GnChan k;
  const int cpg = C / GN_GROUPS, g = (4 * c4) / cpg;
  k.mean = stats[((size_t)b * GN_GROUPS + g) * 2];
  k.rstd = stats[((size_t)b * GN_GROUPS + g) * 2 + 1];
  k.gamma = ldg4(gamma + 4 * c4);
  k.beta = ldg4(beta + 4 * c4);
  if (ss) {
    const float4 sc = ldg4(ss + (size_t)b * ld_ss + 4 * c4);
    k.sc1 = make_float4(1.0f + sc.x, 1.0f + sc.y, 1.0f + sc.z, 1.0f + sc.w);
    k.shift = ldg4(ss + (size_t)b * ld_ss + C + 4 * c4);
  } else {
    k.sc1 = make_float4(1.f, 1.f, 1.f, 1.f);
    k.shift = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return k;
}

__device__ __forceinline__ float4 gn_xhat(const GnChan& k, const float4 x) {
  return make_float4((x.x - k.mean) * k.rstd, (x.y - k.mean) * k.rstd, (x.z - k.mean) * k.rstd, (x.w - k.mean) * k.rstd);
}
__device__ __forceinline__ float4 gn_preact(const GnChan& k, const float4 xh) {
  return make_float4((xh.x * k.gamma.x + k.beta.x) * k.sc1.x + k.shift.x, (xh.y * k.gamma.y + k.beta.y) * k.sc1.y + k.shift.y,
                     (xh.z * k.gamma.z + k.beta.z) * k.sc1.z + k.shift.z, (xh.w * k.gamma.w + k.beta.w) * k.sc1.w + k.shift.w);
}
template <bool SILU, bool RND>
__device__ __forceinline__ float4 gn_act(const GnChan& k, const float4 x) {
  float4 v = gn_preact(k, gn_xhat(k, x));
  if (SILU) v = make_float4(silu_f(v.x), silu_f(v.y), silu_f(v.z), silu_f(v.w));
  if (RND) v = make_float4(round_tf32_f(v.x), round_tf32_f(v.y), round_tf32_f(v.z), round_tf32_f(v.w));
  return v;
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// The same store into a buffer that holds fp16 elements at the same ELEMENT indices (f16 != 0): the operand of an fp16 tensor-core
// conv (round to nearest, saturating).  `base` is the start of the buffer the element index of p is counted from.
__device__ __forceinline__ uint32_t gn_pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void st4x(float* p, float* base, float4 v, int f16) {
  if (f16) {
    uint16_t* hp = reinterpret_cast<uint16_t*>(base) + (p - base);
    *reinterpret_cast<uint2*>(hp) = make_uint2(gn_pack_f16x2(v.x, v.y), gn_pack_f16x2(v.z, v.w));
  } else {
    st4(p, v);
  }
}

// y dense [B,Ho,Wo,C].  RS_NONE / RS_UP walk INPUT pixels (UP writes each result to its 4 outputs); RS_DOWN walks
// OUTPUT pixels (average of the 4 activations, as h_upd(in_rest(x)) in unet.py:317-319).
template <int RS, bool SILU, bool RND>
__global__ void gn_apply_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ ss, int ld_ss,
                                const float* __restrict__ stats, float* __restrict__ y, int H, int W, int C, int pix_chunk, int out_f16) {
  pdl_wait();
  // Blocks walk the tensor BACKWARDS (last image, last pixels first): the producer / statistics pass that ran just before
  // this kernel touched the tail of the tensor last, so that is what the 126 MB L2 still holds.
  const int C4 = C / 4, tid = threadIdx.x, b = gridDim.y - 1 - blockIdx.y, bx = gridDim.x - 1 - blockIdx.x;
  const int c4 = tid % C4, prow = tid / C4, ppi = blockDim.x / C4;
  const GnChan k = gn_load_chan(stats, gamma, beta, ss, ld_ss, b, c4, C);
  const float* xb = x + (size_t)b * H * W * ldx + 4 * c4;
  if (RS == RS_DOWN) {
    const int Ho = H / 2, Wo = W / 2, npix = Ho * Wo;
    float* yb = y + (size_t)b * npix * C + 4 * c4;
    const int p1 = min(npix, (bx + 1) * pix_chunk);
    for (int p = bx * pix_chunk + prow; p < p1; p += ppi) {
      const int ho = p / Wo, wo = p - ho * Wo;
      const float* s0 = xb + ((size_t)(2 * ho) * W + 2 * wo) * ldx;
      const float4 v0 = ldg4(s0), v1 = ldg4(s0 + ldx), v2 = ldg4(s0 + (size_t)W * ldx), v3 = ldg4(s0 + (size_t)W * ldx + ldx);
      const float4 a0 = gn_act<SILU, false>(k, v0), a1 = gn_act<SILU, false>(k, v1), a2 = gn_act<SILU, false>(k, v2),
                   a3 = gn_act<SILU, false>(k, v3);
      float4 o = make_float4((a0.x + a1.x + a2.x + a3.x) * 0.25f, (a0.y + a1.y + a2.y + a3.y) * 0.25f,
                             (a0.z + a1.z + a2.z + a3.z) * 0.25f, (a0.w + a1.w + a2.w + a3.w) * 0.25f);
      if (RND) o = make_float4(round_tf32_f(o.x), round_tf32_f(o.y), round_tf32_f(o.z), round_tf32_f(o.w));
      st4x(yb + (size_t)p * C, y, o, out_f16);
    }
  } else {
    const int npix = H * W;
    const int p1 = min(npix, (bx + 1) * pix_chunk);
    int p = bx * pix_chunk + prow;
    if (RS == RS_NONE) {
      float* yb = y + (size_t)b * npix * C + 4 * c4;
      for (; p + (GN_APPLY_UNROLL - 1) * ppi < p1; p += GN_APPLY_UNROLL * ppi) {
        float4 v[GN_APPLY_UNROLL];
#pragma unroll
        for (int u = 0; u < GN_APPLY_UNROLL; ++u) v[u] = ldg4(xb + (size_t)(p + u * ppi) * ldx);
#pragma unroll
        for (int u = 0; u < GN_APPLY_UNROLL; ++u) st4x(yb + (size_t)(p + u * ppi) * C, y, gn_act<SILU, RND>(k, v[u]), out_f16);
      }
      for (; p < p1; p += ppi) st4x(yb + (size_t)p * C, y, gn_act<SILU, RND>(k, ldg4(xb + (size_t)p * ldx)), out_f16);
    } else {  // RS_UP
      const int Wo = 2 * W;
      float* yb = y + (size_t)b * npix * 4 * C + 4 * c4;
      for (; p < p1; p += ppi) {
        const int h = p / W, w = p - h * W;
        const float4 o = gn_act<SILU, RND>(k, ldg4(xb + (size_t)p * ldx));
        float* d = yb + ((size_t)(2 * h) * Wo + 2 * w) * C;
        st4x(d, y, o, out_f16); st4x(d + C, y, o, out_f16); st4x(d + (size_t)Wo * C, y, o, out_f16); st4x(d + (size_t)Wo * C + C, y, o, out_f16);
      }
    }
  }
  pdl_launch_dependents();   // late trigger: only the launch latency of the next kernel overlaps this one
}

template <int RS>
static void gn_apply_dispatch(const GnArgs& a, float* y, dim3 grid, int tpb, int pix_chunk, cudaStream_t s) {
#define OSM_GN_APPLY(SILU, RND)                                                                                               \
  do {                                                                                                                        \
    OSM_PREFER_SMEM((gn_apply_kernel<RS, SILU, RND>));                                                                        \
    launch_pdl(gn_apply_kernel<RS, SILU, RND>, grid, dim3(tpb), 0, s, a.x, a.ldx, a.gamma, a.beta, a.scale_shift, a.ld_ss,     \
               a.stats, y, a.H, a.W, a.C, pix_chunk, a.out_f16);                                                              \
  } while (0)
  const bool rnd = a.round_tf32 && !a.out_f16;
  if (a.silu) { if (rnd) OSM_GN_APPLY(true, true); else OSM_GN_APPLY(true, false); }
  else        { if (rnd) OSM_GN_APPLY(false, true); else OSM_GN_APPLY(false, false); }
#undef OSM_GN_APPLY
}

int gn_apply_launch(const GnArgs& a, float* y, cudaStream_t s) {
  if (int e = gn_check(a)) return e;
  const int npix = a.resample == RS_DOWN ? (a.H / 2) * (a.W / 2) : a.H * a.W;
  int tpb, chunks, pix_chunk;
  chunking(npix, a.C, a.resample == RS_NONE ? 16 : 8, &tpb, &chunks, &pix_chunk);
  dim3 grid(chunks, a.B);
  if (a.resample == RS_NONE) gn_apply_dispatch<RS_NONE>(a, y, grid, tpb, pix_chunk, s);
  else if (a.resample == RS_DOWN) gn_apply_dispatch<RS_DOWN>(a, y, grid, tpb, pix_chunk, s);
  else gn_apply_dispatch<RS_UP>(a, y, grid, tpb, pix_chunk, s);
  OSM_LAUNCH_CHECK("gn_apply_kernel");
  return OSM_OK;
}

// ---- backward ----
// gradient arriving at input-resolution pixel (h,w) from the dense dy at the resampled resolution
template <int RS>
__device__ __forceinline__ float4 gn_fetch_dy(const float* __restrict__ dyb /* image base + 4*c4 */, int h, int w, int H, int W,
                                              int C) {
  if (RS == RS_NONE) return ldg4(dyb + ((size_t)h * W + w) * C);
  if (RS == RS_DOWN) {  // forward averaged 2x2 -> each input gets a quarter of the coarse gradient
    const float4 g = ldg4(dyb + ((size_t)(h / 2) * (W / 2) + w / 2) * C);
    return make_float4(0.25f * g.x, 0.25f * g.y, 0.25f * g.z, 0.25f * g.w);
  }
  const int Wf = W * 2;  // forward replicated -> sum of the 4 fine gradients
  const float* base = dyb + ((size_t)(2 * h) * Wf + 2 * w) * C;
  const float4 g0 = ldg4(base), g1 = ldg4(base + C), g2 = ldg4(base + (size_t)Wf * C), g3 = ldg4(base + (size_t)Wf * C + C);
  return make_float4(g0.x + g1.x + g2.x + g3.x, g0.y + g1.y + g2.y + g3.y, g0.z + g1.z + g2.z + g3.z, g0.w + g1.w + g2.w + g3.w);
}

template <bool SILU>
__device__ __forceinline__ float4 gn_dxhat(const GnChan& k, const float4 xh, float4 g) {
  if (SILU) {
    const float4 v = gn_preact(k, xh);
    g.x *= silu_grad_f(v.x); g.y *= silu_grad_f(v.y); g.z *= silu_grad_f(v.z); g.w *= silu_grad_f(v.w);
  }
  return make_float4(g.x * k.sc1.x * k.gamma.x, g.y * k.sc1.y * k.gamma.y, g.z * k.sc1.z * k.gamma.z, g.w * k.sc1.w * k.gamma.w);
}

// __launch_bounds__(1024, 1) on the two backward kernels: blocks have up to C / 4 = 1024 threads (C = 2048 concat inputs use
// 512), and the bound caps them at 64 registers - 4 blocks of 256 threads per SM instead of 3 (80 registers): 665 -> 616 us on
// [8,256,256,256] (4.8 -> 5.2 TB/s); 5 blocks (48 registers) spill and fall to 3.9 TB/s.
template <int RS, bool SILU>
__global__ void __launch_bounds__(1024, 1) gn_bwd_reduce_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, const float* __restrict__ ss, int ld_ss,
                                     const float* __restrict__ stats, const float* __restrict__ dy, int C4, int H, int W,
                                     int pix_chunk, int chunks, double* partial, unsigned int* counter, float* bstats) {
  pdl_wait();
  const int tid = threadIdx.x, b = blockIdx.y, C = 4 * C4, HW = H * W;
  const int c4 = tid % C4, prow = tid / C4, ppi = blockDim.x / C4;
  const int p0 = blockIdx.x * pix_chunk;
  const int p1 = min(HW, p0 + pix_chunk);
  const GnChan k = gn_load_chan(stats, gamma, beta, ss, ld_ss, b, c4, C);
  const float* xb = x + (size_t)b * HW * ldx + 4 * c4;
  const size_t ndy = RS == RS_DOWN ? (size_t)HW / 4 : (RS == RS_UP ? (size_t)HW * 4 : (size_t)HW);
  const float* dyb = dy + (size_t)b * ndy * C + 4 * c4;
  double s0 = 0, s1 = 0;
  int p = p0 + prow;
  for (; p + (GN_UNROLL - 1) * ppi < p1; p += GN_UNROLL * ppi) {
    float4 v[GN_UNROLL], g[GN_UNROLL];
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      const int q = p + u * ppi;
      v[u] = ldg4(xb + (size_t)q * ldx);
      g[u] = gn_fetch_dy<RS>(dyb, q / W, q % W, H, W, C);
    }
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u) {
      const float4 xh = gn_xhat(k, v[u]);
      const float4 d = gn_dxhat<SILU>(k, xh, g[u]);
      s0 += (double)((d.x + d.y) + (d.z + d.w));
      s1 += (double)((d.x * xh.x + d.y * xh.y) + (d.z * xh.z + d.w * xh.w));
    }
  }
  for (; p < p1; p += ppi) {
    const float4 xh = gn_xhat(k, ldg4(xb + (size_t)p * ldx));
    const float4 d = gn_dxhat<SILU>(k, xh, gn_fetch_dy<RS>(dyb, p / W, p % W, H, W, C));
    s0 += (double)((d.x + d.y) + (d.z + d.w));
    s1 += (double)((d.x * xh.x + d.y * xh.y) + (d.z * xh.z + d.w * xh.w));
  }
  pdl_launch_dependents();   // main loop done: let the next kernel launch while the reduction tail runs
  gn_group_reduce_and_finalize(s0, s1, C4, chunks, partial, counter, bstats, (double)HW * (4.0 * C4 / GN_GROUPS), 1);
}

__device__ __forceinline__ float4 gn_fetch_addend(const float* __restrict__ ab /* image base + 4*c4 */, int ld, int mode, int h,
                                                  int w, int H, int W) {
  if (mode == ADD_SAME) return ldg4(ab + ((size_t)h * W + w) * ld);
  if (mode == ADD_FROM_COARSE_QUARTER) {
    const float4 g = ldg4(ab + ((size_t)(h / 2) * (W / 2) + w / 2) * ld);
    return make_float4(0.25f * g.x, 0.25f * g.y, 0.25f * g.z, 0.25f * g.w);
  }
  const int Wf = W * 2;
  const float* base = ab + ((size_t)(2 * h) * Wf + 2 * w) * ld;
  const float4 g0 = ldg4(base), g1 = ldg4(base + ld), g2 = ldg4(base + (size_t)Wf * ld), g3 = ldg4(base + (size_t)Wf * ld + ld);
  return make_float4(g0.x + g1.x + g2.x + g3.x, g0.y + g1.y + g2.y + g3.y, g0.z + g1.z + g2.z + g3.z, g0.w + g1.w + g2.w + g3.w);
}

template <int RS, bool SILU>
__global__ void __launch_bounds__(1024, 1) gn_bwd_apply_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ ss, int ld_ss,
                                    const float* __restrict__ stats, const float* __restrict__ bstats,
                                    const float* __restrict__ dy, const float* __restrict__ addend, int ld_add, int add_mode,
                                    float* __restrict__ dx, int ld_dx, int accumulate, int H, int W, int C, int pix_chunk, int dx_f16) {
  pdl_wait();
  const int C4 = C / 4, cpg = C / GN_GROUPS, tid = threadIdx.x, b = gridDim.y - 1 - blockIdx.y, HW = H * W;
  const int bx = gridDim.x - 1 - blockIdx.x;  // backwards, see gn_apply_kernel: the reduction pass read the tail last
  const int c4 = tid % C4, prow = tid / C4, ppi = blockDim.x / C4;
  const GnChan k = gn_load_chan(stats, gamma, beta, ss, ld_ss, b, c4, C);
  const int g = (4 * c4) / cpg;
  const float m1 = bstats[((size_t)b * GN_GROUPS + g) * 2], m2 = bstats[((size_t)b * GN_GROUPS + g) * 2 + 1];
  const float* xb = x + (size_t)b * HW * ldx + 4 * c4;
  const size_t ndy = RS == RS_DOWN ? (size_t)HW / 4 : (RS == RS_UP ? (size_t)HW * 4 : (size_t)HW);
  const float* dyb = dy + (size_t)b * ndy * C + 4 * c4;
  const size_t nadd = add_mode == ADD_FROM_COARSE_QUARTER ? (size_t)HW / 4 : (add_mode == ADD_SUM4_FINE ? (size_t)HW * 4 : (size_t)HW);
  const float* ab = addend ? addend + (size_t)b * nadd * ld_add + 4 * c4 : nullptr;
  float* dxb = dx + (size_t)b * HW * ld_dx + 4 * c4;
  const int p1 = min(HW, (bx + 1) * pix_chunk);
  for (int p = bx * pix_chunk + prow; p < p1; p += 2 * ppi) {
    const int q = p + ppi;
    const bool has2 = q < p1;
    const int h0 = p / W, w0 = p - h0 * W, h1 = q / W, w1 = q - h1 * W;
    const float4 va = ldg4(xb + (size_t)p * ldx);
    const float4 ga = gn_fetch_dy<RS>(dyb, h0, w0, H, W, C);
    float4 vb = va, gb = ga;
    if (has2) { vb = ldg4(xb + (size_t)q * ldx); gb = gn_fetch_dy<RS>(dyb, h1, w1, H, W, C); }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !has2) break;
      const int pp = u ? q : p, hh = u ? h1 : h0, ww = u ? w1 : w0;
      const float4 xh = gn_xhat(k, u ? vb : va);
      const float4 d = gn_dxhat<SILU>(k, xh, u ? gb : ga);
      float4 o = make_float4(k.rstd * (d.x - m1 - xh.x * m2), k.rstd * (d.y - m1 - xh.y * m2), k.rstd * (d.z - m1 - xh.z * m2),
                             k.rstd * (d.w - m1 - xh.w * m2));
      if (add_mode != ADD_NONE) {
        const float4 a = gn_fetch_addend(ab, ld_add, add_mode, hh, ww, H, W);
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      }
      float* dst = dxb + (size_t)pp * ld_dx;
      if (accumulate) {
        const float4 pv = *reinterpret_cast<const float4*>(dst);
        o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
      }
      st4x(dst, dx, o, dx_f16);
    }
  }
  pdl_launch_dependents();   // late trigger: only the launch latency of the next kernel overlaps this one
}

static int gn_bwd_reduce_launch(const GnBwdArgs& a, cudaStream_t s);

int gn_bwd_launch(const GnBwdArgs& a, cudaStream_t s) {
  if (int e = gn_bwd_reduce_launch(a, s)) return e;
  return gn_bwd_apply_launch(a, s);
}

static int gn_bwd_reduce_launch(const GnBwdArgs& a, cudaStream_t s) {
  const GnArgs& f = a.f;
  if (int e = gn_check(f)) return e;
  int tpb, chunks, pix_chunk;
  chunking(f.H * f.W, f.C, 16, &tpb, &chunks, &pix_chunk, gn_red_max_chunks(f.B));
  const dim3 grid(chunks, f.B);
  const size_t sm = tpb * 2 * sizeof(double);
#define OSM_GN_RED(RS, SILU)                                                                                                     \
  do {                                                                                                                           \
    OSM_PREFER_SMEM((gn_bwd_reduce_kernel<RS, SILU>));                                                                           \
    launch_pdl(gn_bwd_reduce_kernel<RS, SILU>, grid, dim3(tpb), sm, s, f.x, f.ldx, f.gamma, f.beta, f.scale_shift, f.ld_ss,      \
               f.stats, a.dy, f.C / 4, f.H, f.W, pix_chunk, chunks, f.partial, f.counter, a.bstats);                            \
  } while (0)
#define OSM_GN_APP(RS, SILU)                                                                                                      \
  do {                                                                                                                            \
    OSM_PREFER_SMEM((gn_bwd_apply_kernel<RS, SILU>));                                                                             \
    launch_pdl(gn_bwd_apply_kernel<RS, SILU>, grid2, dim3(tpb), 0, s, f.x, f.ldx, f.gamma, f.beta, f.scale_shift, f.ld_ss,        \
               f.stats, a.bstats, a.dy, a.addend, a.ld_add, a.add_mode, a.dx, a.ld_dx, a.accumulate, f.H, f.W, f.C, pix_chunk2,   \
               a.dx_f16);                                                                                                        \
  } while (0)
  if (f.resample == RS_NONE) { if (f.silu) OSM_GN_RED(RS_NONE, true); else OSM_GN_RED(RS_NONE, false); }
  else if (f.resample == RS_DOWN) { if (f.silu) OSM_GN_RED(RS_DOWN, true); else OSM_GN_RED(RS_DOWN, false); }
  else { if (f.silu) OSM_GN_RED(RS_UP, true); else OSM_GN_RED(RS_UP, false); }
  OSM_LAUNCH_CHECK("gn_bwd_reduce_kernel");
#undef OSM_GN_RED
  return OSM_OK;
}

int gn_bwd_apply_launch(const GnBwdArgs& a, cudaStream_t s) {
  const GnArgs& f = a.f;
  if (int e = gn_check(f)) return e;
  if (a.ld_dx % 4 || (a.add_mode != ADD_NONE && a.ld_add % 4)) return fail(OSM_ERR_INVALID, "gn_bwd: ld must be a multiple of 4");
  if (a.dx_f16 && a.accumulate) return fail(OSM_ERR_INVALID, "gn_bwd: an fp16 dx cannot be accumulated into");
  int tpb, tpb2, chunks2, pix_chunk2;
  tpb = gn_tpb(f.C);
  // batch 1 / 2: at most one resident wave of blocks (4 per SM) - 1024 equal blocks on 592 slots take two block times for 1.73 of work
  chunking(f.H * f.W, f.C, 8, &tpb2, &chunks2, &pix_chunk2, f.B >= 3 ? GN_MAX_CHUNKS : 592 / f.B);
  const dim3 grid2(chunks2, f.B);
  if (f.resample == RS_NONE) { if (f.silu) OSM_GN_APP(RS_NONE, true); else OSM_GN_APP(RS_NONE, false); }
  else if (f.resample == RS_DOWN) { if (f.silu) OSM_GN_APP(RS_DOWN, true); else OSM_GN_APP(RS_DOWN, false); }
  else { if (f.silu) OSM_GN_APP(RS_UP, true); else OSM_GN_APP(RS_UP, false); }
  OSM_LAUNCH_CHECK("gn_bwd_apply_kernel");
#undef OSM_GN_APP
  return OSM_OK;
}

// ------------------------------------------------------------------------------------------------
// Small tensors (8x8 ... 32x32 levels): statistics AND apply in ONE launch, one CTA - or one thread-block CLUSTER - per
// (image, group).  A group slice is <= 128 KB there, so the second pass re-reads it from L1/L2; what this saves is a whole
// dependent launch (~5 us of launch + drain inside a CUDA graph, more than either pass costs on such a tensor) per GroupNorm,
// forward and backward.  Same arithmetic as the two-kernel path: fp64 sums, fixed order (lane tree, then warps in order, then
// the CTAs of the cluster in rank order).
//   thread -> fixed 4-channel slot j = tid % (cpg / 4) of the group, pixels prow + rank ppi, + CL ppi, ...
// At small batch one CTA per (image, group) is only 32 CTAs per image, so slices beyond a few thousand elements are split over
// a cluster of CL = 2 / 4 / 8 CTAs (interleaved pixel sweeps); the CL partial sums are exchanged through distributed shared
// memory, every CTA adds them in rank order (identical totals everywhere) and applies its own pixels.
// ------------------------------------------------------------------------------------------------
constexpr int GN_SMALL_THREADS = 256;
constexpr int GN_SMALL_CTA_ELEMS = 4096;   // elements one CTA takes at batch < 4 (measured, profiles/r01_gn_small.md)
// largest (image, group) slice that takes the one-launch kernel: H * W * C / 32 elements
static int gn_small_max_elems(int B) {
  static const int forced = [] { const char* e = getenv("OSM_GN_SMALL_MAX"); return e ? atoi(e) : 0; }();
  if (forced) return forced;
  static const int cl_on = [] { const char* e = getenv("OSM_GN_SMALL_CLUSTER"); return e ? atoi(e) : 1; }();
  return B >= 4 ? 32768 : (cl_on ? 8 * GN_SMALL_CTA_ELEMS : GN_SMALL_CTA_ELEMS);
}
// cluster size for a slice of `elems` elements: 1 from batch 4 (32 B CTAs are enough), else the smallest power of two that
// brings a CTA's share down to GN_SMALL_CTA_ELEMS (OSM_GN_SMALL_MAX forces single CTAs: the earlier behaviour)
static int gn_small_cluster(int B, long elems) {
  if (B >= 4) return 1;
  int cl = 1;
  while (cl < 8 && elems > (long)cl * GN_SMALL_CTA_ELEMS) cl *= 2;
  if (elems > (long)cl * GN_SMALL_CTA_ELEMS) cl = 1;
  return cl;
}

static bool gn_small_cached(const GnArgs& a, int cl);
bool gn_small_capable(const GnArgs& a) {
  const int cpg = a.C / GN_GROUPS;
  if (!(a.C % (4 * GN_GROUPS) == 0 && cpg / 4 <= 32 && (long)a.H * a.W * cpg <= gn_small_max_elems(a.B) && a.ldx % 4 == 0)) return false;
  // the 2x resample between the activation and the conv (ResBlock up / down, unet.py:317-320) only in the register-cached form:
  // a thread then owns whole 2x2 blocks (down) / writes its pixels' four copies (up)
  static const int rs_on = [] { const char* e = getenv("OSM_GN_SMALL_RESAMPLE"); return e ? atoi(e) : 1; }();
  if (a.resample != RS_NONE) return rs_on && gn_small_cached(a, gn_small_cluster(a.B, (long)a.H * a.W * cpg));
  return true;
}

__device__ __forceinline__ uint32_t gn_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ double gn_ld_dsmem_f64(const double* local, uint32_t cta) {
  const uint32_t la = (uint32_t)__cvta_generic_to_shared(local);
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(cta));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}

// Block (cl == 1) or cluster total of (s0, s1), identical in every thread of every CTA.  With cl > 1 the caller must run
// gn_small_cluster_exit() before the kernel returns: a CTA may not leave while a peer can still read its shared memory.
__device__ __forceinline__ void gn_small_block_sum(double& s0, double& s1, int cl) {
  __shared__ double red[2 * (GN_SMALL_THREADS / 32) + 2];
  __shared__ double xch[2];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[2 * warp] = s0; red[2 * warp + 1] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < GN_SMALL_THREADS / 32; ++w) { a += red[2 * w]; b += red[2 * w + 1]; }
    red[2 * (GN_SMALL_THREADS / 32)] = a;
    red[2 * (GN_SMALL_THREADS / 32) + 1] = b;
    xch[0] = a; xch[1] = b;
  }
  if (cl > 1) {
    __shared__ double peer[2 * 8];
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    // one thread per peer: the cl remote loads are in flight together (a serial loop pays the SM-to-SM latency cl times)
    if ((int)threadIdx.x < cl) {
      peer[2 * threadIdx.x] = gn_ld_dsmem_f64(&xch[0], threadIdx.x);
      peer[2 * threadIdx.x + 1] = gn_ld_dsmem_f64(&xch[1], threadIdx.x);
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");   // "my reads of the peers are done"; waited for at exit
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, b = 0;
      for (int r = 0; r < cl; ++r) { a += peer[2 * r]; b += peer[2 * r + 1]; }   // rank order
      red[2 * (GN_SMALL_THREADS / 32)] = a;
      red[2 * (GN_SMALL_THREADS / 32) + 1] = b;
    }
  }
  __syncthreads();
  s0 = red[2 * (GN_SMALL_THREADS / 32)];
  s1 = red[2 * (GN_SMALL_THREADS / 32) + 1];
  __syncthreads();
}
__device__ __forceinline__ void gn_small_cluster_exit(int cl) {
  if (cl > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// PDL launch with an optional (cl, 1, 1) thread-block cluster
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, int cl, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cl > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cl; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// grid (32 groups x cl, B), cluster (cl, 1, 1).  CACHE: a thread walks at most GN_SMALL_KEEP pixels (always true at batch < 4, where
// a CTA takes <= 4096 elements): its float4s are loaded ONCE, all at the same time, and stay in registers between the statistics and
// the apply pass - these launches are pure latency, and a second dependent pass of loads was a third of it.
constexpr int GN_SMALL_KEEP = 4;
// RS (CACHE only): RS_UP - each result goes to its four outputs; RS_DOWN - the thread's pixels are the 2x2 block of ONE output pixel
// (HW / 4 <= ppi x cl), whose activations it averages (same arithmetic as gn_apply_kernel).  W = image width (resample only).
template <bool SILU, bool RND, bool CACHE, int RS = RS_NONE>
__global__ void __launch_bounds__(GN_SMALL_THREADS)
gn_small_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ ss, int ld_ss, float* __restrict__ stats, float* __restrict__ y, int HW, int C, int out_f16,
                    int cl, int W) {
  pdl_wait();
  const int rank = cl > 1 ? (int)gn_cluster_rank() : 0;
  const int g = blockIdx.x / cl, b = blockIdx.y, cpg = C / GN_GROUPS, slots = cpg / 4;
  // ppi whole pixels per sweep; with cpg / 4 not a power of two (24 / 48 channels per group) the last few threads idle
  const int ppi = GN_SMALL_THREADS / slots, j = threadIdx.x % slots;
  const int prow = threadIdx.x < ppi * slots ? threadIdx.x / slots + rank * ppi : HW;
  const int pstep = ppi * cl;
  const int c4 = g * slots + j;
  const float* xb = x + (size_t)b * HW * ldx + 4 * c4;
  // the channel constants do not depend on the statistics: requested before the reduction (its barriers would otherwise put their
  // round trip on the critical path of a kernel that is all latency)
  GnChan k;
  k.gamma = ldg4(gamma + 4 * c4);
  k.beta = ldg4(beta + 4 * c4);
  if (ss) {
    const float4 sc = ldg4(ss + (size_t)b * ld_ss + 4 * c4);
    k.sc1 = make_float4(1.0f + sc.x, 1.0f + sc.y, 1.0f + sc.z, 1.0f + sc.w);
    k.shift = ldg4(ss + (size_t)b * ld_ss + C + 4 * c4);
  } else {
    k.sc1 = make_float4(1.f, 1.f, 1.f, 1.f);
    k.shift = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  double s = 0, q = 0;
  float4 xv[GN_SMALL_KEEP];
  // input pixel i of this thread: p + i pstep, or (RS_DOWN) corner i of the 2x2 block of output pixel prow
  auto pix = [&](int i) -> int {
    if (RS != RS_DOWN) return prow + i * pstep;
    if (prow >= HW / 4) return HW;
    const int Wo = W / 2, ho = prow / Wo, wo = prow - ho * Wo;
    return (2 * ho + (i >> 1)) * W + 2 * wo + (i & 1);
  };
  if (CACHE) {
#pragma unroll
    for (int i = 0; i < GN_SMALL_KEEP; ++i) {
      const int p = pix(i);
      if (p < HW) xv[i] = ldg4(xb + (size_t)p * ldx);
    }
#pragma unroll
    for (int i = 0; i < GN_SMALL_KEEP; ++i) {
      if (pix(i) < HW) {
        const float4 v = xv[i];
        s += (double)((v.x + v.y) + (v.z + v.w));
        q += (double)((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
      }
    }
  } else {
    for (int p = prow; p < HW; p += pstep) {
      const float4 v = ldg4(xb + (size_t)p * ldx);
      s += (double)((v.x + v.y) + (v.z + v.w));
      q += (double)((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
    }
  }
  gn_small_block_sum(s, q, cl);
  const double N = (double)HW * cpg;
  const double mean = s / N;
  double var = q / N - mean * mean;
  if (var < 0) var = 0;
  const float fm = (float)mean, fr = (float)(1.0 / sqrt(var + (double)GN_EPS));
  if (threadIdx.x == 0 && rank == 0) {
    stats[((size_t)b * GN_GROUPS + g) * 2] = fm;
    stats[((size_t)b * GN_GROUPS + g) * 2 + 1] = fr;
  }
  k.mean = fm; k.rstd = fr;
  float* yb = y + (size_t)b * HW * C + 4 * c4;
  if (CACHE && RS == RS_DOWN) {
    if (prow < HW / 4) {
      const float4 a0 = gn_act<SILU, false>(k, xv[0]), a1 = gn_act<SILU, false>(k, xv[1]), a2 = gn_act<SILU, false>(k, xv[2]),
                   a3 = gn_act<SILU, false>(k, xv[3]);
      float4 o = make_float4((a0.x + a1.x + a2.x + a3.x) * 0.25f, (a0.y + a1.y + a2.y + a3.y) * 0.25f,
                             (a0.z + a1.z + a2.z + a3.z) * 0.25f, (a0.w + a1.w + a2.w + a3.w) * 0.25f);
      if (RND) o = make_float4(round_tf32_f(o.x), round_tf32_f(o.y), round_tf32_f(o.z), round_tf32_f(o.w));
      st4x(y + (size_t)b * (HW / 4) * C + 4 * c4 + (size_t)prow * C, y, o, out_f16);
    }
  } else if (CACHE && RS == RS_UP) {
    const int Wo = 2 * W;
    float* yu = y + (size_t)b * HW * 4 * C + 4 * c4;
#pragma unroll
    for (int i = 0; i < GN_SMALL_KEEP; ++i) {
      const int p = prow + i * pstep;
      if (p < HW) {
        const int h = p / W, w = p - h * W;
        const float4 o = gn_act<SILU, RND>(k, xv[i]);
        float* d = yu + ((size_t)(2 * h) * Wo + 2 * w) * C;
        st4x(d, y, o, out_f16); st4x(d + C, y, o, out_f16); st4x(d + (size_t)Wo * C, y, o, out_f16); st4x(d + (size_t)Wo * C + C, y, o, out_f16);
      }
    }
  } else if (CACHE) {
#pragma unroll
    for (int i = 0; i < GN_SMALL_KEEP; ++i) {
      const int p = prow + i * pstep;
      if (p < HW) st4x(yb + (size_t)p * C, y, gn_act<SILU, RND>(k, xv[i]), out_f16);
    }
  } else {
    for (int p = prow; p < HW; p += pstep) st4x(yb + (size_t)p * C, y, gn_act<SILU, RND>(k, ldg4(xb + (size_t)p * ldx)), out_f16);
  }
  pdl_launch_dependents();   // late trigger: only the launch latency of the next kernel overlaps this one
  gn_small_cluster_exit(cl);
}

// does every thread of the one-launch kernels walk at most GN_SMALL_KEEP pixels?  (ppi whole pixels per sweep and CTA, cl CTAs)
static bool gn_small_cached(const GnArgs& a, int cl) {
  const int slots = a.C / GN_GROUPS / 4, ppi = GN_SMALL_THREADS / slots;
  return (long)GN_SMALL_KEEP * ppi * cl >= (long)a.H * a.W;
}

int gn_small_fwd_launch(const GnArgs& a, float* y, cudaStream_t s) {
  if (int e = gn_check(a)) return e;
  if (!gn_small_capable(a)) return fail(OSM_ERR_INVALID, "gn_small_fwd: tensor not eligible");
  const int cl = gn_small_cluster(a.B, (long)a.H * a.W * (a.C / GN_GROUPS));
  const dim3 grid(GN_GROUPS * cl, a.B);
#define OSM_GN_SMALL_C(SILU, RND, CACHE, RS)                                                                                      \
  launch_pdl_cluster(gn_small_fwd_kernel<SILU, RND, CACHE, RS>, grid, dim3(GN_SMALL_THREADS), cl, s, a.x, a.ldx, a.gamma, a.beta, \
                     a.scale_shift, a.ld_ss, a.stats, y, a.H * a.W, a.C, a.out_f16, cl, a.W)
#define OSM_GN_SMALL(SILU, RND)                                                          \
  do {                                                                                   \
    if (a.resample == RS_DOWN) OSM_GN_SMALL_C(SILU, RND, true, RS_DOWN);                 \
    else if (a.resample == RS_UP) OSM_GN_SMALL_C(SILU, RND, true, RS_UP);                \
    else if (cached) OSM_GN_SMALL_C(SILU, RND, true, RS_NONE);                           \
    else OSM_GN_SMALL_C(SILU, RND, false, RS_NONE);                                      \
  } while (0)
  const bool rnd = a.round_tf32 && !a.out_f16;
  const bool cached = gn_small_cached(a, cl);
  if (a.silu) { if (rnd) OSM_GN_SMALL(true, true); else OSM_GN_SMALL(true, false); }
  else        { if (rnd) OSM_GN_SMALL(false, true); else OSM_GN_SMALL(false, false); }
#undef OSM_GN_SMALL
#undef OSM_GN_SMALL_C
  OSM_LAUNCH_CHECK("gn_small_fwd_kernel");
  return OSM_OK;
}

// grid (32 groups x cl, B), cluster (cl, 1, 1): the two backward means and the input gradient in one launch (resample none)
template <bool SILU, bool CACHE, int RS = RS_NONE>
__global__ void __launch_bounds__(GN_SMALL_THREADS)
gn_small_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ ss, int ld_ss, const float* __restrict__ stats, const float* __restrict__ dy,
                    const float* __restrict__ addend, int ld_add, int add_mode, float* __restrict__ dx, int ld_dx, int accumulate,
                    int H, int W, int C, int dx_f16, int cl) {
  pdl_wait();
  const int rank = cl > 1 ? (int)gn_cluster_rank() : 0;
  const int g = blockIdx.x / cl, b = blockIdx.y, cpg = C / GN_GROUPS, slots = cpg / 4, HW = H * W;
  const int ppi = GN_SMALL_THREADS / slots, j = threadIdx.x % slots;
  const int prow = threadIdx.x < ppi * slots ? threadIdx.x / slots + rank * ppi : HW;
  const int pstep = ppi * cl;
  const int c4 = g * slots + j;
  const GnChan k = gn_load_chan(stats, gamma, beta, ss, ld_ss, b, c4, C);
  const float* xb = x + (size_t)b * HW * ldx + 4 * c4;
  const size_t ndy = RS == RS_DOWN ? (size_t)HW / 4 : (RS == RS_UP ? (size_t)HW * 4 : (size_t)HW);   // dy lives at the resampled resolution
  const float* dyb = dy + (size_t)b * ndy * C + 4 * c4;
  const size_t nadd = add_mode == ADD_FROM_COARSE_QUARTER ? (size_t)HW / 4 : (add_mode == ADD_SUM4_FINE ? (size_t)HW * 4 : (size_t)HW);
  const float* ab = addend ? addend + (size_t)b * nadd * ld_add + 4 * c4 : nullptr;
  float* dxb = dx + (size_t)b * HW * ld_dx + 4 * c4;
  double s0 = 0, s1 = 0;
  if (CACHE) {
    // x, dy, the skip-path addend and the previous dx of the thread's <= 4 pixels: every load of the kernel in flight at once, kept in
    // registers across the reduction (see gn_small_fwd_kernel)
    float4 xv[GN_SMALL_KEEP], dv[GN_SMALL_KEEP], av[GN_SMALL_KEEP], pv[GN_SMALL_KEEP];
#pragma unroll
    for (int i = 0; i < GN_SMALL_KEEP; ++i) {
      const int p = prow + i * pstep;
      if (p < HW) {
        xv[i] = ldg4(xb + (size_t)p * ldx);
        dv[i] = gn_fetch_dy<RS>(dyb, p / W, p % W, H, W, C);
      }
    }
#pragma unroll
    for (int i = 0; i < GN_SMALL_KEEP; ++i) {
      const int p = prow + i * pstep;
      av[i] = pv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < HW) {
        if (add_mode != ADD_NONE) { const int hh = p / W, ww = p - hh * W; av[i] = gn_fetch_addend(ab, ld_add, add_mode, hh, ww, H, W); }
        if (accumulate) pv[i] = *reinterpret_cast<const float4*>(dxb + (size_t)p * ld_dx);
      }
    }
#pragma unroll
    for (int i = 0; i < GN_SMALL_KEEP; ++i) {
      if (prow + i * pstep < HW) {
        const float4 xh = gn_xhat(k, xv[i]);
        const float4 d = gn_dxhat<SILU>(k, xh, dv[i]);
        s0 += (double)((d.x + d.y) + (d.z + d.w));
        s1 += (double)((d.x * xh.x + d.y * xh.y) + (d.z * xh.z + d.w * xh.w));
      }
    }
    gn_small_block_sum(s0, s1, cl);
    const double N = (double)HW * cpg;
    const float m1 = (float)(s0 / N), m2 = (float)(s1 / N);
#pragma unroll
    for (int i = 0; i < GN_SMALL_KEEP; ++i) {
      const int p = prow + i * pstep;
      if (p < HW) {
        const float4 xh = gn_xhat(k, xv[i]);
        const float4 d = gn_dxhat<SILU>(k, xh, dv[i]);
        float4 o = make_float4(k.rstd * (d.x - m1 - xh.x * m2), k.rstd * (d.y - m1 - xh.y * m2), k.rstd * (d.z - m1 - xh.z * m2),
                               k.rstd * (d.w - m1 - xh.w * m2));
        if (add_mode != ADD_NONE) { o.x += av[i].x; o.y += av[i].y; o.z += av[i].z; o.w += av[i].w; }
        if (accumulate) { o.x += pv[i].x; o.y += pv[i].y; o.z += pv[i].z; o.w += pv[i].w; }
        st4x(dxb + (size_t)p * ld_dx, dx, o, dx_f16);
      }
    }
    pdl_launch_dependents();
    gn_small_cluster_exit(cl);
    return;
  }
  for (int p = prow; p < HW; p += pstep) {
    const float4 xh = gn_xhat(k, ldg4(xb + (size_t)p * ldx));
    const float4 d = gn_dxhat<SILU>(k, xh, ldg4(dyb + (size_t)p * C));
    s0 += (double)((d.x + d.y) + (d.z + d.w));
    s1 += (double)((d.x * xh.x + d.y * xh.y) + (d.z * xh.z + d.w * xh.w));
  }
  gn_small_block_sum(s0, s1, cl);
  const double N = (double)HW * cpg;
  const float m1 = (float)(s0 / N), m2 = (float)(s1 / N);
  for (int p = prow; p < HW; p += pstep) {
    const float4 xh = gn_xhat(k, ldg4(xb + (size_t)p * ldx));
    const float4 d = gn_dxhat<SILU>(k, xh, ldg4(dyb + (size_t)p * C));
    float4 o = make_float4(k.rstd * (d.x - m1 - xh.x * m2), k.rstd * (d.y - m1 - xh.y * m2), k.rstd * (d.z - m1 - xh.z * m2),
                           k.rstd * (d.w - m1 - xh.w * m2));
    if (add_mode != ADD_NONE) {
      const int hh = p / W, ww = p - hh * W;
      const float4 a = gn_fetch_addend(ab, ld_add, add_mode, hh, ww, H, W);
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    float* dst = dxb + (size_t)p * ld_dx;
    if (accumulate) {
      const float4 pv = *reinterpret_cast<const float4*>(dst);
      o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
    }
    st4x(dst, dx, o, dx_f16);
  }
  pdl_launch_dependents();   // late trigger: only the launch latency of the next kernel overlaps this one
  gn_small_cluster_exit(cl);
}

int gn_small_bwd_launch(const GnBwdArgs& a, cudaStream_t s) {
  const GnArgs& f = a.f;
  if (int e = gn_check(f)) return e;
  if (!gn_small_capable(f)) return fail(OSM_ERR_INVALID, "gn_small_bwd: tensor not eligible");
  if (a.ld_dx % 4 || (a.add_mode != ADD_NONE && a.ld_add % 4)) return fail(OSM_ERR_INVALID, "gn_bwd: ld must be a multiple of 4");
  if (a.dx_f16 && a.accumulate) return fail(OSM_ERR_INVALID, "gn_bwd: an fp16 dx cannot be accumulated into");
  const int cl = gn_small_cluster(f.B, (long)f.H * f.W * (f.C / GN_GROUPS));
  const dim3 grid(GN_GROUPS * cl, f.B);
#define OSM_GN_SMALLB_C(SILU, CACHE, RS)                                                                                      \
  launch_pdl_cluster(gn_small_bwd_kernel<SILU, CACHE, RS>, grid, dim3(GN_SMALL_THREADS), cl, s, f.x, f.ldx, f.gamma, f.beta,      \
                     f.scale_shift, f.ld_ss, f.stats, a.dy, a.addend, a.ld_add, a.add_mode, a.dx, a.ld_dx, a.accumulate, f.H, f.W, \
                     f.C, a.dx_f16, cl)
#define OSM_GN_SMALLB(SILU)                                                          \
  do {                                                                               \
    if (f.resample == RS_DOWN) OSM_GN_SMALLB_C(SILU, true, RS_DOWN);                 \
    else if (f.resample == RS_UP) OSM_GN_SMALLB_C(SILU, true, RS_UP);                \
    else if (cached) OSM_GN_SMALLB_C(SILU, true, RS_NONE);                           \
    else OSM_GN_SMALLB_C(SILU, false, RS_NONE);                                      \
  } while (0)
  const bool cached = gn_small_cached(f, cl);
  if (f.silu) OSM_GN_SMALLB(true); else OSM_GN_SMALLB(false);
#undef OSM_GN_SMALLB
#undef OSM_GN_SMALLB_C
  OSM_LAUNCH_CHECK("gn_small_bwd_kernel");
  return OSM_OK;
}

// ---- statistics reduced in the conv epilogue (conv_epilogue.cuh): fold the per-(tile, warp) partials ----
// grid (32 groups, B), 256 threads; thread t sums slots t, t+256, ... in fp64 (eight loads in flight), then a fixed-order tree
// (lanes by shuffle, the eight warps in order).  A launch of this kernel sits between a conv and its consumer ~85 times per
// step, so its length is pure latency: one batch of loads for the 2048 slots of a 256x256 image, and the operands of the
// coefficients (which do not depend on the sums) requested before the reduction.
// coef != null (mode 1): also write the forward operand-transform coefficients (a, b) of the group's channels for the GroupNorm
// (gamma, beta, scale-shift ss) that consumes these statistics - what gn_coef_fwd_kernel would do in a launch of its own.
constexpr int GN_FIN_THREADS = 256;
__global__ void __launch_bounds__(GN_FIN_THREADS) gn_fused_finalize_kernel(const float* __restrict__ partial, int slots,
                                                                           const float* __restrict__ fwd_stats, float* __restrict__ out, double N,
                                                                           int mode, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                           const float* __restrict__ ss, int ld_ss, float2* __restrict__ coef, int C) {
  __shared__ double r0[GN_FIN_THREADS / 32], r1[GN_FIN_THREADS / 32];
  __shared__ float s_stat[2];
  pdl_launch_dependents();
  pdl_wait();

  const int g = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int cpg = C / GN_GROUPS;
  const bool wc = coef && mode == 1;
  float ga = 0.f, be = 0.f, sc1 = 1.f, sh = 0.f;      // channel g cpg + tid of the coefficient table (cpg <= 256 in practice)
  if (wc && tid < cpg) {
    const int c = g * cpg + tid;
    ga = gamma[c]; be = beta[c];
    if (ss) { sc1 = 1.0f + ss[(size_t)b * ld_ss + c]; sh = ss[(size_t)b * ld_ss + C + c]; }
  }
  double mean_f = 0, rstd_f = 0;
  if (mode != 1 && tid == 0) { mean_f = fwd_stats[((size_t)b * GN_GROUPS + g) * 2]; rstd_f = fwd_stats[((size_t)b * GN_GROUPS + g) * 2 + 1]; }
  const float2* src = reinterpret_cast<const float2*>(partial) + (size_t)b * slots * GN_GROUPS + g;
  double s0 = 0, s1 = 0;
  for (int i = tid; i < slots; i += 8 * GN_FIN_THREADS) {
    float2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = i + u * GN_FIN_THREADS < slots ? __ldcg(src + (size_t)(i + u * GN_FIN_THREADS) * GN_GROUPS) : make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) { s0 += (double)v[u].x; s1 += (double)v[u].y; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if ((tid & 31) == 0) { r0[tid >> 5] = s0; r1[tid >> 5] = s1; }
  __syncthreads();
  if (tid == 0) {
    double t0 = 0, t1 = 0;
    for (int w = 0; w < GN_FIN_THREADS / 32; ++w) { t0 += r0[w]; t1 += r1[w]; }
    r0[0] = t0; r1[0] = t1;
    float* o = out + ((size_t)b * GN_GROUPS + g) * 2;
    if (mode == 1) {
      const double mean = r0[0] / N;
      double var = r1[0] / N - mean * mean;
      if (var < 0) var = 0;
      o[0] = (float)mean;
      o[1] = (float)(1.0 / sqrt(var + (double)GN_EPS));
      s_stat[0] = o[0]; s_stat[1] = o[1];
    } else {  // sum d, sum d x  ->  mean d, mean d xhat  with xhat = (x - mean) rstd
      o[0] = (float)(r0[0] / N);
      o[1] = (float)((rstd_f * r1[0] - mean_f * rstd_f * r0[0]) / N);
    }
  }
  if (wc) {
    __syncthreads();
    const float mean = s_stat[0], rstd = s_stat[1];
    if (tid < cpg) coef[(size_t)b * C + g * cpg + tid] = make_float2(rstd * ga * sc1, (be - mean * rstd * ga) * sc1 + sh);
    for (int t = tid + GN_FIN_THREADS; t < cpg; t += GN_FIN_THREADS) {   // more than 256 channels per group: not in the shipped configs
      const int c = g * cpg + t;
      const float ga2 = gamma[c], be2 = beta[c];
      const float sc2 = ss ? 1.0f + ss[(size_t)b * ld_ss + c] : 1.0f;
      const float sh2 = ss ? ss[(size_t)b * ld_ss + C + c] : 0.0f;
      coef[(size_t)b * C + c] = make_float2(rstd * ga2 * sc2, (be2 - mean * rstd * ga2) * sc2 + sh2);
    }
  }
}

int gn_fused_finalize_launch(const float* partial, int slots_per_image, const float* fwd_stats, float* out, int B, int HW, int C,
                             int mode, cudaStream_t s, const GnArgs* coef_gn, float* coef) {
  OSM_PREFER_SMEM(gn_fused_finalize_kernel);
  const bool wc = coef_gn && coef && mode == 1;
  OSM_LAUNCH_PDL("gn_fused_finalize_kernel", gn_fused_finalize_kernel, dim3(GN_GROUPS, B), dim3(GN_FIN_THREADS), 0, s, partial, slots_per_image,
                 fwd_stats, out, (double)HW * (C / GN_GROUPS), mode, wc ? coef_gn->gamma : nullptr, wc ? coef_gn->beta : nullptr,
                 wc ? coef_gn->scale_shift : nullptr, wc ? coef_gn->ld_ss : 0, wc ? (float2*)coef : nullptr, C);
  return OSM_OK;
}

// Statistics of a skip-concat input [h | hs] (equal halves) from the statistics of its two halves, which the epilogues of the
// producing convs already reduced: a concat group is the union of two consecutive groups of one half.  grid (B), 32 threads = groups.
__global__ void gn_combine_stats_kernel(const float* __restrict__ sa, const float* __restrict__ sb, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, g = threadIdx.x;
  const float* src = (g < GN_GROUPS / 2 ? sa : sb) + ((size_t)b * GN_GROUPS + 2 * (g % (GN_GROUPS / 2))) * 2;
  const double m1 = src[0], r1 = src[1], m2 = src[2], r2 = src[3];
  const double v1 = 1.0 / (r1 * r1) - (double)GN_EPS, v2 = 1.0 / (r2 * r2) - (double)GN_EPS;
  const double mean = 0.5 * (m1 + m2);
  double var = 0.5 * ((v1 + m1 * m1) + (v2 + m2 * m2)) - mean * mean;
  if (var < 0) var = 0;
  out[((size_t)b * GN_GROUPS + g) * 2] = (float)mean;
  out[((size_t)b * GN_GROUPS + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)GN_EPS));
}

int gn_combine_stats_launch(const float* stats_a, const float* stats_b, float* out, int B, cudaStream_t s) {
  OSM_PREFER_SMEM(gn_combine_stats_kernel);
  OSM_LAUNCH_PDL("gn_combine_stats_kernel", gn_combine_stats_kernel, dim3(B), dim3(GN_GROUPS), 0, s, stats_a, stats_b, out);
  return OSM_OK;
}

// All backward-statistics coefficient sets of one input-VJP in ONE launch: grid (descriptors, B).  Every set depends only on the
// forward statistics and the step's scale-shift vector, so they are computed together at the start of the backward program.
__global__ void gn_coef_batch_kernel(const GnCoefDesc* __restrict__ table) {
  pdl_launch_dependents();
  pdl_wait();
  const GnCoefDesc d = table[blockIdx.x];
  const int b = blockIdx.y, C = d.C, cpg = C / GN_GROUPS;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mean = d.stats[((size_t)b * GN_GROUPS + g) * 2], rstd = d.stats[((size_t)b * GN_GROUPS + g) * 2 + 1];
    const float ga = d.gamma[c], be = d.beta[c];
    const float sc1 = d.ss ? 1.0f + d.ss[(size_t)b * d.ld_ss + c] : 1.0f;
    const float sh = d.ss ? d.ss[(size_t)b * d.ld_ss + C + c] : 0.0f;
    d.coef[(size_t)b * C + c] = make_float4(rstd * ga * sc1, (be - mean * rstd * ga) * sc1 + sh, sc1 * ga, 0.f);
  }
}

int gn_coef_batch_launch(const GnCoefDesc* table, int n, int B, cudaStream_t s) {
  if (n <= 0) return OSM_OK;
  OSM_PREFER_SMEM(gn_coef_batch_kernel);
  OSM_LAUNCH_PDL("gn_coef_batch_kernel", gn_coef_batch_kernel, dim3(n, B), dim3(256), 0, s, table);
  return OSM_OK;
}

// per (image, channel): (a, b, e, 0) with  pre-activation = x a + b  and  d(xhat-gradient) = g silu'(.) e
__global__ void gn_coef_kernel(const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ ss, int ld_ss, float4* __restrict__ coef, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int g = c / (C / GN_GROUPS);
  const float mean = stats[((size_t)b * GN_GROUPS + g) * 2], rstd = stats[((size_t)b * GN_GROUPS + g) * 2 + 1];
  const float ga = gamma[c], be = beta[c];
  const float sc1 = ss ? 1.0f + ss[(size_t)b * ld_ss + c] : 1.0f;
  const float sh = ss ? ss[(size_t)b * ld_ss + C + c] : 0.0f;
  coef[(size_t)b * C + c] = make_float4(rstd * ga * sc1, (be - mean * rstd * ga) * sc1 + sh, sc1 * ga, 0.f);
}

// the forward half alone, packed (a, b) per (image, channel): what the halo conv kernel's operand transform reads
__global__ void gn_coef_fwd_kernel(const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ ss, int ld_ss, float2* __restrict__ coef, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int g = c / (C / GN_GROUPS);
  const float mean = stats[((size_t)b * GN_GROUPS + g) * 2], rstd = stats[((size_t)b * GN_GROUPS + g) * 2 + 1];
  const float ga = gamma[c], be = beta[c];
  const float sc1 = ss ? 1.0f + ss[(size_t)b * ld_ss + c] : 1.0f;
  const float sh = ss ? ss[(size_t)b * ld_ss + C + c] : 0.0f;
  coef[(size_t)b * C + c] = make_float2(rstd * ga * sc1, (be - mean * rstd * ga) * sc1 + sh);
}

int gn_coef_fwd_launch(const GnArgs& a, float* coef, cudaStream_t s) {
  OSM_PREFER_SMEM(gn_coef_fwd_kernel);
  OSM_LAUNCH_PDL("gn_coef_fwd_kernel", gn_coef_fwd_kernel, dim3((a.C + 255) / 256, a.B), dim3(256), 0, s, (const float*)a.stats, a.gamma,
                 a.beta, a.scale_shift, a.ld_ss, (float2*)coef, a.C);
  return OSM_OK;
}

int gn_coef_launch(const GnArgs& a, float* coef, cudaStream_t s) {
  OSM_PREFER_SMEM(gn_coef_kernel);
  OSM_LAUNCH_PDL("gn_coef_kernel", gn_coef_kernel, dim3((a.C + 255) / 256, a.B), dim3(256), 0, s, (const float*)a.stats, a.gamma, a.beta,
                 a.scale_shift, a.ld_ss, (float4*)coef, a.C);
  return OSM_OK;
}

}  // namespace osm
