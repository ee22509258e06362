// QKVAttentionLegacy forward / backward (unet.py:407-437), fp32 like the reference's einsum + fp32 softmax.
// Token-major layout: qkv [B, L, 3C]; head h owns channels [3*ch*h, 3*ch*(h+1)) split as (q | k | v),
// which is exactly the reference's `qkv.reshape(bs*n_heads, ch*3, length).split(ch, dim=1)`.
//
// Round-1 implementation: a strided-batched fp32 CUDA-core GEMM (64x64x16 tiles) + warp-per-row softmax,
// P materialised per (image, head) in an L2-friendly scratch and recomputed in the backward (the
// reference checkpoints the whole block, nn.py:124-170, so nothing but qkv is kept here either).
// Attention is 0.54 % of the step's FLOPs; a tcgen05 flash kernel replaces this in a later round.
#include "common.cuh"

namespace osm {

struct BGemm {
  const float* A; const float* B; float* C;
  long sam, sak, sbk, sbn, scm;       // element strides: A(m,k), B(k,n), C(m, n contiguous)
  long sAb, sAh, sBb, sBh, sCb, sCh;  // batch strides for (image, head)
  int M, N, K, heads;
  float alpha;
};

constexpr int BG_T = 64, BG_K = 16;

__global__ void __launch_bounds__(256) bgemm_kernel(BGemm g) {
  __shared__ float As[BG_K][BG_T + 4];
  __shared__ float Bs[BG_K][BG_T + 4];
  const int z = blockIdx.z, b = z / g.heads, h = z % g.heads;
  const float* A = g.A + b * g.sAb + h * g.sAh;
  const float* B = g.B + b * g.sBb + h * g.sBh;
  float* C = g.C + b * g.sCb + h * g.sCh;
  const int m0 = blockIdx.y * BG_T, n0 = blockIdx.x * BG_T;
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += BG_K) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int m, k;
      if (g.sak == 1) { m = e / BG_K; k = e % BG_K; } else { m = e % BG_T; k = e / BG_T; }
      float v = 0.f;
      if (m0 + m < g.M && k0 + k < g.K) v = A[(long)(m0 + m) * g.sam + (long)(k0 + k) * g.sak];
      As[k][m] = v;
      int n, kk;
      if (g.sbn == 1) { n = e % BG_T; kk = e / BG_T; } else { kk = e % BG_K; n = e / BG_K; }
      float u = 0.f;
      if (n0 + n < g.N && k0 + kk < g.K) u = B[(long)(k0 + kk) * g.sbk + (long)(n0 + n) * g.sbn];
      Bs[kk][n] = u;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BG_K; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < g.N) C[(long)m * g.scm + n] = g.alpha * acc[i][j];
    }
  }
}

static int bgemm_launch(const BGemm& g, int batches, cudaStream_t s) {
  dim3 grid((g.N + BG_T - 1) / BG_T, (g.M + BG_T - 1) / BG_T, batches);
  bgemm_kernel<<<grid, 256, 0, s>>>(g);
  OSM_LAUNCH_CHECK("bgemm_kernel");
  return OSM_OK;
}

// one warp per row of length L: P = softmax(S) in place
__global__ void softmax_rows_kernel(float* __restrict__ S, long rows, int L) {
  const long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* r = S + row * L;
  float mx = -INFINITY;
  for (int i = lane; i < L; i += 32) mx = fmaxf(mx, r[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int i = lane; i < L; i += 32) {
    const float e = expf(r[i] - mx);
    r[i] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  for (int i = lane; i < L; i += 32) r[i] *= inv;
}

// D <- alpha * P o (D - rowsum(D o P))
__global__ void softmax_bwd_rows_kernel(const float* __restrict__ P, float* __restrict__ D, long rows, int L, float alpha) {
  const long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = P + row * L;
  float* d = D + row * L;
  float dot = 0.f;
  for (int i = lane; i < L; i += 32) dot += p[i] * d[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  for (int i = lane; i < L; i += 32) d[i] = alpha * p[i] * (d[i] - dot);
}

static int scores_softmax(const float* qkv, float* P, int B, int L, int C, int heads, cudaStream_t s) {
  const int ch = C / heads;
  const long C3 = 3L * C;
  BGemm g{};
  g.A = qkv; g.B = qkv + ch; g.C = P;
  g.sam = C3; g.sak = 1; g.sbk = 1; g.sbn = C3; g.scm = L;
  g.sAb = (long)L * C3; g.sAh = 3L * ch; g.sBb = (long)L * C3; g.sBh = 3L * ch;
  g.sCb = (long)heads * L * L; g.sCh = (long)L * L;
  g.M = L; g.N = L; g.K = ch; g.heads = heads;
  g.alpha = 1.0f / sqrtf((float)ch);  // (q ch^-1/4) . (k ch^-1/4)
  if (int e = bgemm_launch(g, B * heads, s)) return e;
  const long rows = (long)B * heads * L;
  softmax_rows_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, s>>>(P, rows, L);
  OSM_LAUNCH_CHECK("softmax_rows_kernel");
  return OSM_OK;
}

int attention_fwd_launch(const float* qkv, float* out, float* P, int B, int L, int C, int heads, cudaStream_t s) {
  if (C % heads) return fail(OSM_ERR_INVALID, "attention: C must be divisible by heads");
  const int ch = C / heads;
  const long C3 = 3L * C;
  if (int e = scores_softmax(qkv, P, B, L, C, heads, s)) return e;
  BGemm g{};
  g.A = P; g.B = qkv + 2 * ch; g.C = out;
  g.sam = L; g.sak = 1; g.sbk = C3; g.sbn = 1; g.scm = C;
  g.sAb = (long)heads * L * L; g.sAh = (long)L * L; g.sBb = (long)L * C3; g.sBh = 3L * ch;
  g.sCb = (long)L * C; g.sCh = ch;
  g.M = L; g.N = ch; g.K = L; g.heads = heads; g.alpha = 1.0f;
  return bgemm_launch(g, B * heads, s);
}

int attention_bwd_launch(const float* qkv, const float* g_out, float* g_qkv, float* P, float* D, int B, int L, int C, int heads,
                         cudaStream_t s) {
  if (C % heads) return fail(OSM_ERR_INVALID, "attention: C must be divisible by heads");
  const int ch = C / heads;
  const long C3 = 3L * C, LL = (long)L * L;
  if (int e = scores_softmax(qkv, P, B, L, C, heads, s)) return e;
  BGemm g{};
  g.heads = heads; g.alpha = 1.0f;
  // g_V[s,c] = sum_t P[t,s] g_a[t,c]
  g.A = P; g.sam = 1; g.sak = L; g.sAb = heads * LL; g.sAh = LL;
  g.B = g_out; g.sbk = C; g.sbn = 1; g.sBb = (long)L * C; g.sBh = ch;
  g.C = g_qkv + 2 * ch; g.scm = C3; g.sCb = (long)L * C3; g.sCh = 3L * ch;
  g.M = L; g.N = ch; g.K = L;
  if (int e = bgemm_launch(g, B * heads, s)) return e;
  // D[t,s] = sum_c g_a[t,c] V[s,c]
  g.A = g_out; g.sam = C; g.sak = 1; g.sAb = (long)L * C; g.sAh = ch;
  g.B = qkv + 2 * ch; g.sbk = 1; g.sbn = C3; g.sBb = (long)L * C3; g.sBh = 3L * ch;
  g.C = D; g.scm = L; g.sCb = heads * LL; g.sCh = LL;
  g.M = L; g.N = L; g.K = ch;
  if (int e = bgemm_launch(g, B * heads, s)) return e;
  const long rows = (long)B * heads * L;
  softmax_bwd_rows_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, s>>>(P, D, rows, L, 1.0f / sqrtf((float)ch));
  OSM_LAUNCH_CHECK("softmax_bwd_rows_kernel");
  // g_Q[t,c] = sum_s dS[t,s] K[s,c]
  g.A = D; g.sam = L; g.sak = 1; g.sAb = heads * LL; g.sAh = LL;
  g.B = qkv + ch; g.sbk = C3; g.sbn = 1; g.sBb = (long)L * C3; g.sBh = 3L * ch;
  g.C = g_qkv; g.scm = C3; g.sCb = (long)L * C3; g.sCh = 3L * ch;
  g.M = L; g.N = ch; g.K = L;
  if (int e = bgemm_launch(g, B * heads, s)) return e;
  // g_K[s,c] = sum_t dS[t,s] Q[t,c]
  g.A = D; g.sam = 1; g.sak = L; g.sAb = heads * LL; g.sAh = LL;
  g.B = qkv; g.sbk = C3; g.sbn = 1; g.sBb = (long)L * C3; g.sBh = 3L * ch;
  g.C = g_qkv + ch; g.scm = C3; g.sCb = (long)L * C3; g.sCh = 3L * ch;
  g.M = L; g.N = ch; g.K = L;
  return bgemm_launch(g, B * heads, s);
}

int attention_launches(int which) { return which == 0 ? 3 : 7; }

}  // namespace osm
