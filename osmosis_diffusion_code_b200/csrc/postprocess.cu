// Batched post-processing of finished samples on the device (the reference does this per image on the CPU after
// p_sample_loop returns: osmosis_sampling.py:207-292, osmosis_utils/utils.py:46-114, 748-763).
//
//   postprocess_kernel      clipped RGB, the re-degraded image 2 A_phi(x0) - 1, its per-image residual norm against the
//                           measurement, and the restored image exp(phi_a d) (y01 - backscatter)
//   minmax_quantile_kernel  min_max_norm_range / min_max_norm_range_percentile of a [B, n] plane: clip to the
//                           (q_lo, q_hi) quantiles (torch.quantile 'linear': exact order statistics by a 4-pass radix
//                           select on the float keys + torch's lerp), then affine map of [min, max] to [vmin, vmax]
//   colormap_kernel         matplotlib-style lookup of a [0,1] plane in a 256-entry RGB table
//
// HBM-bound elementwise work; one CTA per image for the quantile kernel (an image plane is 256 KB: L2/L1-resident
// across the radix passes).  Reductions are fixed-order (no float atomics).
#include <math.h>

#include "common.cuh"

namespace osm {

namespace {

struct DepthFnP {
  int kind;
  float v0, v1, v2;
};
__device__ __forceinline__ float depth_convert_p(const DepthFnP& f, float d) {  // utils.py:529-566 (convert_depth)
  if (f.kind == OSM_DEPTH_GAMMA) {
    const float base = __fmul_rn(__fadd_rn(d, f.v0), f.v1);
    return (f.v2 == 1.0f) ? base : powf(base, f.v2);
  }
  if (f.kind == OSM_DEPTH_MOVE) return __fadd_rn(d, f.v0);
  return __fmul_rn(0.5f, __fadd_rn(d, 1.0f));
}

constexpr int PP_THREADS = 256;

// grid (blocks, B).  norm_partial: [B][gridDim.x] doubles, folded by the last block of each image (ticket counter).
__global__ void __launch_bounds__(PP_THREADS)
postprocess_kernel(int op_kind, DepthFnP df, const float* __restrict__ x0, const float* __restrict__ y, const float* __restrict__ phi,
                   float* __restrict__ rgb_clip, float* __restrict__ degraded, float* __restrict__ recon, double* __restrict__ norm_partial,
                   unsigned int* __restrict__ counter, float* __restrict__ norm_out, int HW) {
  const int b = blockIdx.y;
  const float* ph = phi + 9 * b;
  float pa[3], pb[3], pinf[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    pa[c] = (op_kind == OSM_OP_HAZE) ? ph[0] : ph[c];
    pb[c] = (op_kind == OSM_OP_UNDERWATER_REVISED) ? ph[3 + c] : pa[c];
    pinf[c] = ph[6 + c];
  }
  const float* xb = x0 + (size_t)b * 4 * HW;
  const float* yb = y + (size_t)b * 3 * HW;
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const float d = depth_convert_p(df, xb[3 * (size_t)HW + i]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t o = (size_t)b * 3 * HW + (size_t)c * HW + i;
      const float J = __fmul_rn(0.5f, __fadd_rn(xb[(size_t)c * HW + i], 1.0f));      // sample_rgb_01            (:214)
      const float back = __fmul_rn(pinf[c], __fsub_rn(1.0f, expf(__fmul_rn(-pb[c], d))));  // backscatter_image  (:245)
      const float att = expf(__fmul_rn(-pa[c], d));                                   // attenuation_image        (:246)
      const float fwd = __fadd_rn(__fmul_rn(J, att), back);                           // forward_predicted_image  (:247)
      const float deg = __fsub_rn(__fmul_rn(2.0f, fwd), 1.0f);                        // degraded_image           (:250)
      const float yy = yb[(size_t)c * HW + i];
      const float y01 = __fmul_rn(0.5f, __fadd_rn(yy, 1.0f));                         // ref_img_01
      rgb_clip[o] = fminf(fmaxf(J, 0.0f), 1.0f);                                      // sample_rgb_01_clip       (:215)
      degraded[o] = deg;
      recon[o] = __fmul_rn(expf(__fmul_rn(pa[c], d)), __fsub_rn(y01, back));          // sample_rgb_recon         (:255-256)
      const float r = __fsub_rn(deg, yy);
      acc += (double)r * (double)r;
    }
  }
  __shared__ double red[PP_THREADS];
  __shared__ int s_last;
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = PP_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    __stcg(&norm_partial[(size_t)b * gridDim.x + blockIdx.x], red[0]);
    __threadfence();
    s_last = atomicAdd(&counter[b], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    for (unsigned int k = 0; k < gridDim.x; ++k) t += __ldcg(&norm_partial[(size_t)b * gridDim.x + k]);
    norm_out[b] = (float)sqrt(t);   // torch.linalg.norm(degraded_image - ref_img)                                 (:251)
    counter[b] = 0;
  }
}

// ---- quantile clip + min-max normalisation, one CTA per image ----
__device__ __forceinline__ uint32_t f2key(float f) {  // order-preserving map float -> uint32
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(u);
}

constexpr int MQ_THREADS = 1024;

// exact k-th smallest (0-based) of v[0..n): 4 radix passes of 8 bits over the keys, 256-bin shared histogram
__device__ float block_select(const float* __restrict__ v, int n, int k, unsigned int* hist, unsigned int* sh) {
  uint32_t prefix = 0, mask = 0;
  int kk = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = f2key(v[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int cum = 0;
      int bsel = 255;
      for (int bkt = 0; bkt < 256; ++bkt) {
        if (cum + hist[bkt] > (unsigned int)kk) { bsel = bkt; break; }
        cum += hist[bkt];
      }
      sh[0] = (unsigned int)bsel;
      sh[1] = cum;
    }
    __syncthreads();
    prefix |= sh[0] << shift;
    mask |= 255u << shift;
    kk -= (int)sh[1];
    __syncthreads();
  }
  return key2f(prefix);
}

// value of rank k+1 given v_k: the smallest element greater than v_k unless v_k itself occupies rank k+1 too
__device__ float block_next(const float* __restrict__ v, int n, int k, float vk, float* fred, unsigned int* ured) {
  unsigned int cnt = 0;
  float nxt = INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = v[i];
    cnt += x <= vk;
    if (x > vk) nxt = fminf(nxt, x);
  }
  fred[threadIdx.x] = nxt;
  ured[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      fred[threadIdx.x] = fminf(fred[threadIdx.x], fred[threadIdx.x + o]);
      ured[threadIdx.x] += ured[threadIdx.x + o];
    }
    __syncthreads();
  }
  const float r = (ured[0] >= (unsigned int)(k + 2)) ? vk : fred[0];
  __syncthreads();
  return r;
}

// torch.quantile(..., interpolation='linear') in the input dtype: rank = q (n - 1) (fp32), lerp(below, above, weight)
__device__ float block_quantile(const float* __restrict__ v, int n, float q, unsigned int* hist, unsigned int* sh, float* fred,
                                unsigned int* ured) {
  const float rank = __fmul_rn(q, (float)(n - 1));
  const float below = floorf(rank);
  const float w = __fsub_rn(rank, below);
  const int k = (int)below;
  const float lo = block_select(v, n, k, hist, sh);
  if (w == 0.0f || k + 1 >= n) return lo;
  const float hi = block_next(v, n, k, lo, fred, ured);
  // at::lerp: weight < 0.5 ? a + w (b - a) : b - (b - a) (1 - w), each a fused multiply-add in torch's CPU and CUDA builds
  // (checked against torch.quantile bit for bit: the unfused form differs in ~1 % of the cases by one ulp)
  const float diff = __fsub_rn(hi, lo);
  return (w < 0.5f) ? __fmaf_rn(w, diff, lo) : __fmaf_rn(-diff, __fsub_rn(1.0f, w), hi);
}

__global__ void __launch_bounds__(MQ_THREADS)
minmax_quantile_kernel(const float* __restrict__ img, float* __restrict__ out, int n, float q_lo, float q_hi, float vmin, float vmax) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int sh[2];
  __shared__ float fred[MQ_THREADS];
  __shared__ unsigned int ured[MQ_THREADS];
  const float* v = img + (size_t)blockIdx.x * n;
  float* o = out + (size_t)blockIdx.x * n;
  float lo, hi;
  if (q_lo <= 0.0f && q_hi >= 1.0f) {  // plain min / max (min_max_norm_range, utils.py:46-76)
    float mn = INFINITY, mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { mn = fminf(mn, v[i]); mx = fmaxf(mx, v[i]); }
    fred[threadIdx.x] = mn;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) fred[threadIdx.x] = fminf(fred[threadIdx.x], fred[threadIdx.x + s]);
      __syncthreads();
    }
    lo = fred[0];
    __syncthreads();
    fred[threadIdx.x] = mx;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) fred[threadIdx.x] = fmaxf(fred[threadIdx.x], fred[threadIdx.x + s]);
      __syncthreads();
    }
    hi = fred[0];
    __syncthreads();
  } else {  // clip to the quantiles first; min / max of the clipped image are then the quantiles themselves (utils.py:85-100)
    lo = block_quantile(v, n, q_lo, hist, sh, fred, ured);
    hi = block_quantile(v, n, q_hi, hist, sh, fred, ured);
  }
  if (lo == hi) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) o[i] = 0.0f;
    return;
  }
  const float scale = __fdiv_rn(__fsub_rn(vmax, vmin), __fsub_rn(hi, lo));
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float c = fminf(fmaxf(v[i], lo), hi);
    o[i] = __fadd_rn(__fmul_rn(__fsub_rn(c, lo), scale), vmin);
  }
}

// matplotlib Colormap.__call__ on floats in [0,1] with an N = 256 table: index = int(x * 256), x == 1 -> 255
__global__ void colormap_kernel(const float* __restrict__ img, const float* __restrict__ lut, float* __restrict__ out, int n) {
  const int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float x = img[(size_t)b * n + i];
    int idx = (int)(x * 256.0f);
    idx = idx < 0 ? 0 : (idx > 255 ? 255 : idx);
#pragma unroll
    for (int c = 0; c < 3; ++c) out[((size_t)b * 3 + c) * n + i] = __ldg(lut + 3 * idx + c);
  }
}

struct PostScratch {
  double* partial = nullptr;
  unsigned int* counter = nullptr;
  int cap_b = 0;
  int ensure(int B, int blocks) {
    if (B <= cap_b) return OSM_OK;
    if (partial) { cudaFree(partial); cudaFree(counter); }
    OSM_CUDA_CHECK(cudaMalloc(&partial, (size_t)B * blocks * sizeof(double)));
    OSM_CUDA_CHECK(cudaMalloc(&counter, (size_t)B * sizeof(unsigned int)));
    OSM_CUDA_CHECK(cudaMemset(counter, 0, (size_t)B * sizeof(unsigned int)));
    cap_b = B;
    return OSM_OK;
  }
};
PostScratch g_post;
constexpr int PP_BLOCKS = 64;

}  // namespace

int postprocess_launch(int op_kind, int depth_kind, const float* dv, const float* x0, const float* y, const float* phi, float* rgb_clip,
                       float* degraded, float* recon, float* norm_out, int B, int HW, cudaStream_t s) {
  if (int e = g_post.ensure(B, PP_BLOCKS)) return e;   // one-time scratch (per process), like the GroupNorm test scratch
  DepthFnP df{depth_kind, dv[0], dv[1], dv[2]};
  postprocess_kernel<<<dim3(PP_BLOCKS, B), PP_THREADS, 0, s>>>(op_kind, df, x0, y, phi, rgb_clip, degraded, recon, g_post.partial,
                                                                g_post.counter, norm_out, HW);
  OSM_LAUNCH_CHECK("postprocess_kernel");
  return OSM_OK;
}

int minmax_quantile_launch(const float* img, float* out, int B, int n, float q_lo, float q_hi, float vmin, float vmax, cudaStream_t s) {
  if (n < 2) return fail(OSM_ERR_INVALID, "minmax_quantile: need at least 2 elements per image");
  minmax_quantile_kernel<<<B, MQ_THREADS, 0, s>>>(img, out, n, q_lo, q_hi, vmin, vmax);
  OSM_LAUNCH_CHECK("minmax_quantile_kernel");
  return OSM_OK;
}

int colormap_launch(const float* img, const float* lut, float* out, int B, int n, cudaStream_t s) {
  int blocks = (n + 255) / 256;
  if (blocks > 592) blocks = 592;
  colormap_kernel<<<dim3(blocks, B), 256, 0, s>>>(img, lut, out, n);
  OSM_LAUNCH_CHECK("colormap_kernel");
  return OSM_OK;
}

}  // namespace osm
