// QKVAttentionLegacy forward / backward (unet.py:407-437), fp32 like the reference's einsum + fp32 softmax.
// Token-major layout: qkv [B, L, 3C]; head h owns channels [3*ch*h, 3*ch*(h+1)) split as (q | k | v),
// which is exactly the reference's `qkv.reshape(bs*n_heads, ch*3, length).split(ch, dim=1)`.
//
// Round-1 implementation: a strided-batched tensor-core GEMM (mma.sync m16n8k8 TF32 with the 3xTF32 split, i.e. fp32-level
// accuracy; 64x64x16 tiles) + warp-per-row fp32 softmax, P materialised per (image, head) in an L2-friendly scratch and
// recomputed in the backward (the reference checkpoints the whole block, nn.py:124-170, so nothing but qkv is kept here
// either).  Attention is 0.54 % of the step's FLOPs; a fused tcgen05 flash kernel (no P round trip) is the next step.
#include "common.cuh"

namespace osm {

struct BGemm {
  const float* A; const float* B; float* C;
  long sam, sak, sbk, sbn, scm;       // element strides: A(m,k), B(k,n), C(m, n contiguous)
  long sAb, sAh, sBb, sBh, sCb, sCh;  // batch strides for (image, head)
  int M, N, K, heads;
  float alpha;
};

constexpr int BG_T = 64, BG_K = 16, BG_LD = BG_T + 8;  // row stride 72 = 8 (mod 32): conflict-free mma fragment loads

// 3xTF32 split: x = hi + lo with hi, lo representable in TF32; a*b ~= hi_a*hi_b + hi_a*lo_b + lo_a*hi_b keeps fp32-level
// accuracy on the tensor cores (the reference computes attention in fp32).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// C[z](m,n) = alpha * sum_k A[z](m,k) B[z](k,n) with arbitrary element strides.  64x64x16 tiles staged in shared memory
// (k-major), 8 warps as 2 (m) x 4 (n), each warp 32x16 of C = 2x2 mma.sync m16n8k8 tiles, 3 MMAs per tile (3xTF32).
__global__ void __launch_bounds__(256) bgemm_kernel(BGemm g) {
  __shared__ float As[BG_K][BG_LD];
  __shared__ float Bs[BG_K][BG_LD];
  const int z = blockIdx.z, b = z / g.heads, h = z % g.heads;
  const float* A = g.A + b * g.sAb + h * g.sAh;
  const float* B = g.B + b * g.sBb + h * g.sBh;
  float* C = g.C + b * g.sCb + h * g.sCh;
  const int m0 = blockIdx.y * BG_T, n0 = blockIdx.x * BG_T;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;           // mma fragment coordinates
  const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;
  float acc[2][2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

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
    for (int ks = 0; ks < BG_K; ks += 8) {
      uint32_t ah[2][4], al[2][4], bh[2][2], bl[2][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int m = wm + 16 * i + gq;
        split_tf32(As[ks + tq][m], ah[i][0], al[i][0]);
        split_tf32(As[ks + tq][m + 8], ah[i][1], al[i][1]);
        split_tf32(As[ks + tq + 4][m], ah[i][2], al[i][2]);
        split_tf32(As[ks + tq + 4][m + 8], ah[i][3], al[i][3]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = wn + 8 * j + gq;
        split_tf32(Bs[ks + tq][n], bh[j][0], bl[j][0]);
        split_tf32(Bs[ks + tq + 4][n], bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          mma_tf32_16x8x8(acc[i][j], al[i], bh[j]);   // small terms first
          mma_tf32_16x8x8(acc[i][j], ah[i], bl[j]);
          mma_tf32_16x8x8(acc[i][j], ah[i], bh[j]);
        }
    }
    __syncthreads();
  }
  // accumulator fragment: c0,c1 -> (row gq, cols 2tq, 2tq+1); c2,c3 -> (row gq+8, same cols)
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int m = m0 + wm + 16 * i + gq + 8 * hrow;
        const int n = n0 + wn + 8 * j + 2 * tq;
        if (m < g.M) {
          if (n < g.N) C[(long)m * g.scm + n] = g.alpha * acc[i][j][2 * hrow];
          if (n + 1 < g.N) C[(long)m * g.scm + n + 1] = g.alpha * acc[i][j][2 * hrow + 1];
        }
      }
}

static int bgemm_launch(const BGemm& g, int batches, cudaStream_t s) {
  dim3 grid((g.N + BG_T - 1) / BG_T, (g.M + BG_T - 1) / BG_T, batches);
  OSM_PREFER_SMEM(bgemm_kernel);
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
  OSM_PREFER_SMEM(softmax_rows_kernel);
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
  OSM_PREFER_SMEM(softmax_bwd_rows_kernel);
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
