// Small kernels: sinusoidal timestep embedding, the timestep MLP / per-block emb linears, NCHW<->NHWC
// boundary transposes, conv weight repacking.
#include <cuda_fp16.h>

#include "common.cuh"

namespace osm {

// emb[b, j] = cos(t_b f_j) (j < half) | sin(t_b f_{j-half}),  f_j = exp(-ln(1e4) j / half)  (nn.py:103-121)
__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int dim) {
  const int b = blockIdx.x, half = dim / 2;
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    // the reference evaluates -log(10000) * arange / half in fp32 (mul then div), then exp, then t * freq
    const float f = expf(__fdiv_rn(__fmul_rn(-9.210340371976184f, (float)j), (float)half));
    const float a = __fmul_rn(t[b], f);
    out[(size_t)b * dim + j] = cosf(a);
    out[(size_t)b * dim + half + j] = sinf(a);
  }
  if ((dim & 1) && threadIdx.x == 0) out[(size_t)b * dim + dim - 1] = 0.f;
}

int timestep_embedding_launch(const float* t, float* out, int B, int dim, cudaStream_t s) {
  timestep_embedding_kernel<<<B, 128, 0, s>>>(t, out, dim);
  OSM_LAUNCH_CHECK("timestep_embedding_kernel");
  return OSM_OK;
}

// out[b,n] = bias[n] + sum_k act(in[b,k]) W[n,k].  One warp per output feature n, looping over the batch in chunks of 8 so
// each weight row is read from HBM once per chunk (weight-bandwidth bound: K*N*4 bytes - 212 MB for the packed emb linears).
// The (SiLU'd) activations of a chunk are staged once per block in shared memory (the first version recomputed the SiLU in
// every warp: 53 M expf per call at B = 1), and a lane issues all its weight loads of a 1024-wide slab before the FMAs.
constexpr int LIN_BCHUNK = 8;
constexpr int LIN_WARPS = 8;
constexpr int LIN_UNROLL = 8;   // float4 weight loads in flight per lane (8 x 128 columns = one K = 1024 row)
__global__ void __launch_bounds__(LIN_WARPS * 32)
linear_kernel(const float* __restrict__ in, int ld_in, const float* __restrict__ W, const float* __restrict__ bias,
              float* __restrict__ out, int ld_out, int B, int K, int N, int silu_in) {
  extern __shared__ float4 s_in4[];  // [LIN_BCHUNK][K / 4]
  const int lane = threadIdx.x & 31, n = blockIdx.x * LIN_WARPS + (threadIdx.x >> 5);
  const int K4 = K / 4;
  const float4* wr = reinterpret_cast<const float4*>(W + (size_t)(n < N ? n : 0) * K);
  for (int b0 = 0; b0 < B; b0 += LIN_BCHUNK) {
    const int nb = min(LIN_BCHUNK, B - b0);
    for (int i = threadIdx.x; i < nb * K4; i += blockDim.x) {
      const int bi = i / K4, k4 = i - bi * K4;
      float4 v = *reinterpret_cast<const float4*>(in + (size_t)(b0 + bi) * ld_in + 4 * k4);
      if (silu_in) {
        v.x = v.x / (1.0f + expf(-v.x)); v.y = v.y / (1.0f + expf(-v.y));
        v.z = v.z / (1.0f + expf(-v.z)); v.w = v.w / (1.0f + expf(-v.w));
      }
      s_in4[i] = v;
    }
    __syncthreads();
    if (n < N) {
      float acc[LIN_BCHUNK];
#pragma unroll
      for (int i = 0; i < LIN_BCHUNK; ++i) acc[i] = 0.f;
      for (int k0 = 0; k0 < K4; k0 += 32 * LIN_UNROLL) {
        float4 w4[LIN_UNROLL];
#pragma unroll
        for (int u = 0; u < LIN_UNROLL; ++u) {
          const int k4 = k0 + u * 32 + lane;
          w4[u] = k4 < K4 ? __ldg(wr + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < LIN_UNROLL; ++u) {
          const int k4 = k0 + u * 32 + lane;
          if (k4 < K4) {
#pragma unroll
            for (int i = 0; i < LIN_BCHUNK; ++i) {
              if (i < nb) {
                const float4 v = s_in4[i * K4 + k4];
                acc[i] += v.x * w4[u].x + v.y * w4[u].y + v.z * w4[u].z + v.w * w4[u].w;
              }
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < LIN_BCHUNK; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && i < nb) out[(size_t)(b0 + i) * ld_out + n] = v + (bias ? bias[n] : 0.f);
      }
    }
    __syncthreads();
  }
}

int linear_launch(const float* in, int ld_in, const float* W, const float* bias, float* out, int ld_out, int B, int K, int N,
                  int silu_in, cudaStream_t s) {
  if (K % 4 || ld_in % 4) return fail(OSM_ERR_INVALID, "linear: K and ld_in must be multiples of 4");
  const size_t smem = (size_t)LIN_BCHUNK * K * sizeof(float);
  if (smem > 200 * 1024) return fail(OSM_ERR_INVALID, "linear: K too large for the staged activations");
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    OSM_CUDA_CHECK(cudaFuncSetAttribute(linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  linear_kernel<<<(N + LIN_WARPS - 1) / LIN_WARPS, LIN_WARPS * 32, smem, s>>>(in, ld_in, W, bias, out, ld_out, B, K, N, silu_in);
  OSM_LAUNCH_CHECK("linear_kernel");
  return OSM_OK;
}

// Power-of-two scale derived from the bit pattern of a per-image maximum |x| (amax_bits_kernel): amax * scale lies in
// [2^texp, 2^(texp+1)).  Multiplying by a power of two is exact in fp32, so scaling a linear pass on the way in and unscaling it on
// the way out changes nothing but the exponent range its intermediate values occupy (what the fp16-operand convs need).
__device__ __forceinline__ float pow2_scale_from_bits(unsigned int bits, int texp, int inverse) {
  const int e = (int)((bits >> 23) & 0xffu);
  if (e == 0 || e == 255) return 1.0f;           // zero / denormal / non-finite maximum: leave the data alone
  int k = texp - (e - 127);
  k = k < -126 ? -126 : (k > 126 ? 126 : k);
  return __uint_as_float((unsigned int)(127 + (inverse ? -k : k)) << 23);
}

// bits[b] = max over image b of the bit pattern of |x| (non-negative floats order like unsigned integers); bits zeroed by the caller.
// The maximum is order-independent, so the atomics do not cost reproducibility.
__global__ void amax_bits_kernel(const float* __restrict__ src, unsigned int* __restrict__ bits, size_t n) {
  const int b = blockIdx.y;
  const float* s = src + (size_t)b * n;
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(s[i]));
  unsigned int u = __float_as_uint(m);
  for (int d = 16; d; d >>= 1) u = max(u, __shfl_xor_sync(0xffffffffu, u, d));
  if ((threadIdx.x & 31) == 0 && u) atomicMax(bits + b, u);
}

int amax_bits_launch(const float* src, unsigned int* bits, int B, size_t n_per_image, cudaStream_t s) {
  OSM_CUDA_CHECK(cudaMemsetAsync(bits, 0, (size_t)B * sizeof(unsigned int), s));
  int blocks = (int)((n_per_image + 256 * 8 - 1) / (256 * 8));
  if (blocks > 256) blocks = 256;
  if (blocks < 1) blocks = 1;
  amax_bits_kernel<<<dim3(blocks, B), 256, 0, s>>>(src, bits, n_per_image);
  OSM_LAUNCH_CHECK("amax_bits_kernel");
  return OSM_OK;
}

// [B,C,HW] -> [B,HW,Cp] (channels >= C zero-filled); scale_bits != null: image b is multiplied by its power-of-two scale
__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW, int Cp,
                                        const unsigned int* __restrict__ scale_bits, int texp) {
  const int b = blockIdx.y;
  const float sc = scale_bits ? pow2_scale_from_bits(scale_bits[b], texp, 0) : 1.0f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    float* d = dst + ((size_t)b * HW + p) * Cp;
    for (int c = 0; c < Cp; c += 4) {
      float4 v;
      v.x = (c + 0 < C) ? sc * src[((size_t)b * C + c + 0) * HW + p] : 0.f;
      v.y = (c + 1 < C) ? sc * src[((size_t)b * C + c + 1) * HW + p] : 0.f;
      v.z = (c + 2 < C) ? sc * src[((size_t)b * C + c + 2) * HW + p] : 0.f;
      v.w = (c + 3 < C) ? sc * src[((size_t)b * C + c + 3) * HW + p] : 0.f;
      *reinterpret_cast<float4*>(d + c) = v;
    }
  }
}

int nchw_to_nhwc_pad_launch(const float* src, float* dst, int B, int C, int HW, int Cp, cudaStream_t s, const unsigned int* scale_bits,
                            int texp) {
  int blocks = (HW + 255) / 256;
  nchw_to_nhwc_pad_kernel<<<dim3(blocks, B), 256, 0, s>>>(src, dst, C, HW, Cp, scale_bits, texp);
  OSM_LAUNCH_CHECK("nchw_to_nhwc_pad_kernel");
  return OSM_OK;
}

// [B,HW,ld] (first C channels) -> [B,C,HW]; scale_bits != null: image b is divided by its power-of-two scale
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, int ld, float* __restrict__ dst, int C, int HW,
                                    const unsigned int* __restrict__ scale_bits, int texp) {
  const int b = blockIdx.y;
  const float sc = scale_bits ? pow2_scale_from_bits(scale_bits[b], texp, 1) : 1.0f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const float* s = src + ((size_t)b * HW + p) * ld;
    for (int c = 0; c < C; ++c) dst[((size_t)b * C + c) * HW + p] = sc * s[c];
  }
}

int nhwc_to_nchw_launch(const float* src, int ld, float* dst, int B, int C, int HW, cudaStream_t s, const unsigned int* scale_bits,
                        int texp) {
  int blocks = (HW + 255) / 256;
  nhwc_to_nchw_kernel<<<dim3(blocks, B), 256, 0, s>>>(src, ld, dst, C, HW, scale_bits, texp);
  OSM_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return OSM_OK;
}

// OIHW (or OI1 for 1x1 / Conv1d) -> K-block-major packs (a B-operand tile of any width is ONE contiguous run in DRAM):
//   forward  Wf[tap][ci/32][co][ci%32]
//   dgrad    Wd[tap'][co/32][ci][co%32],  tap' = taps-1-tap   (spatially flipped, transposed)
// Padded rows/cols are zero.  round_tf32: store the RN-rounded TF32 value so the tensor core's mantissa
// truncation is exact on the weight operand.
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, float* __restrict__ wf, float* __restrict__ wd, int Cout,
                                        int Cin, int Cout_p, int Cin_p, int taps, int round_tf32) {
  const size_t total = (size_t)taps * Cout_p * Cin_p;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin_p);
    const int co = (int)((i / Cin_p) % Cout_p);
    const int tap = (int)(i / ((size_t)Cin_p * Cout_p));
    float v = 0.f;
    if (ci < Cin && co < Cout) v = w[((size_t)co * Cin + ci) * taps + tap];
    if (round_tf32) {
      uint32_t r;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
      v = __uint_as_float(r);
    }
    if (wf) wf[(((size_t)tap * (Cin_p / 32) + ci / 32) * Cout_p + co) * 32 + (ci & 31)] = v;
    if (wd) wd[(((size_t)(taps - 1 - tap) * (Cout_p / 32) + co / 32) * Cin_p + ci) * 32 + (co & 31)] = v;
  }
}

int pack_conv_weight_launch(const float* w_oihw, float* w_fwd, float* w_dgrad, int Cout, int Cin, int Cout_p, int Cin_p,
                            int taps, int round_tf32, cudaStream_t s) {
  const size_t total = (size_t)taps * Cout_p * Cin_p;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  pack_conv_weight_kernel<<<(unsigned)blocks, 256, 0, s>>>(w_oihw, w_fwd, w_dgrad, Cout, Cin, Cout_p, Cin_p, taps, round_tf32);
  OSM_LAUNCH_CHECK("pack_conv_weight_kernel");
  return OSM_OK;
}

__global__ void pack_conv_weight_f16_kernel(const float* __restrict__ w, __half* __restrict__ wf, __half* __restrict__ wd, int Cout,
                                            int Cin, int Cout_p, int Cin_p, int taps) {
  const size_t total = (size_t)taps * Cout_p * Cin_p;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin_p);
    const int co = (int)((i / Cin_p) % Cout_p);
    const int tap = (int)(i / ((size_t)Cin_p * Cout_p));
    float v = 0.f;
    if (ci < Cin && co < Cout) v = w[((size_t)co * Cin + ci) * taps + tap];
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half hv = __float2half_rn(v);
    if (wf) wf[(((size_t)tap * (Cin_p / 64) + ci / 64) * Cout_p + co) * 64 + (ci & 63)] = hv;
    if (wd) wd[(((size_t)(taps - 1 - tap) * (Cout_p / 64) + co / 64) * Cin_p + ci) * 64 + (co & 63)] = hv;
  }
}

int pack_conv_weight_f16_launch(const float* w_oihw, void* w_fwd, void* w_dgrad, int Cout, int Cin, int Cout_p, int Cin_p, int taps,
                                cudaStream_t s) {
  if ((w_fwd && Cin_p % 64) || (w_dgrad && Cout_p % 64)) return fail(OSM_ERR_INVALID, "pack_conv_weight_f16: K dimension must be a multiple of 64");
  const size_t total = (size_t)taps * Cout_p * Cin_p;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  pack_conv_weight_f16_kernel<<<(unsigned)blocks, 256, 0, s>>>(w_oihw, (__half*)w_fwd, (__half*)w_dgrad, Cout, Cin, Cout_p, Cin_p, taps);
  OSM_LAUNCH_CHECK("pack_conv_weight_f16_kernel");
  return OSM_OK;
}

}  // namespace osm
