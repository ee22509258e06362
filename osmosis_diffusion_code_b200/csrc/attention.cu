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

// ------------------------------------------------------------------------------------------------
// One-launch fp32 attention for the 8x8 level (L = 64 tokens, 64 channels per head): everything of one (image, head) - q, k, v, the
// 64x64 probabilities and, in the backward, dO / dP / dS - lives in shared memory of ONE CTA, so the block is one launch forward and
// one backward instead of 2 + 3 (transpose + fused tcgen05 kernels) whose serial TMA -> MMA -> softmax -> MMA chain, not the 2 x 0.5
// MFLOP of math, set the time (8 + 23 us -> see DESIGN.md).  Plain fp32 FMAs: exact mode and product mode share it.
//   thread (tx, ty) of 16 x 16 owns the 4 x 4 outputs (ty + 16 i, tx + 16 j): operand rows are read as broadcasts, columns at
//   consecutive banks (row stride 65 floats).  Reference: unet.py:407-437 (QKVAttentionLegacy) and its autograd backward.
// ------------------------------------------------------------------------------------------------
constexpr int AS_L = 64, AS_CH = 64, AS_LD = 68;   // row stride 68 floats: rows stay 16-byte aligned, 8 rows at float4 = 32 distinct banks
constexpr int AS_TILE = AS_L * AS_LD;   // floats per [64][68] tile

bool attention_small_ok(int L, int C, int heads) {
  static const int on = [] { const char* e = getenv("OSM_ATTN_SMALL"); return e ? atoi(e) : 1; }();
  return on && L == AS_L && heads > 0 && C == heads * AS_CH;
}

__device__ __forceinline__ float4 as_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// NT tiles (up to 4) of 64 x 64 (row = token, 64 consecutive channels, ld floats between tokens) -> shared [64][68]; all loads of a
// thread are issued before its first store (one L2 round trip per kernel instead of one per loop iteration)
template <int NT>
__device__ __forceinline__ void as_load_tiles(float* const (&dst)[NT], const float* const (&src)[NT], const long (&ld)[NT]) {
  float4 v[NT][4];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = threadIdx.x + 256 * u, r = e >> 4, c4 = e & 15;
      v[t][u] = __ldg(reinterpret_cast<const float4*>(src[t] + (long)r * ld[t]) + c4);
    }
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = threadIdx.x + 256 * u, r = e >> 4, c4 = e & 15;
      *reinterpret_cast<float4*>(dst[t] + r * AS_LD + 4 * c4) = v[t][u];
    }
}
__device__ __forceinline__ void as_zero(float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}
#define AS_FMA4(ai, bj, i, j) acc[i][j] = fmaf(ai.w, bj.w, fmaf(ai.z, bj.z, fmaf(ai.y, bj.y, fmaf(ai.x, bj.x, acc[i][j]))))
// "A B^T": acc[i][j] = sum_k A[ty + 16 i][k] * Bm[tx + 16 j][k]   (k contiguous in both: float4 reads; A rows are warp broadcasts,
// the 16 B rows of a half-warp sit 272 bytes apart = conflict-free per quarter-warp)
__device__ __forceinline__ void as_mm_abt(const float* A, const float* Bm, int tx, int ty, float (&acc)[4][4]) {
  as_zero(acc);
#pragma unroll 2
  for (int k = 0; k < 64; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = as_ld4(A + (ty + 16 * i) * AS_LD + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = as_ld4(Bm + (tx + 16 * j) * AS_LD + k);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) AS_FMA4(a[i], b[j], i, j);
  }
}
// "A B": acc[i][j] = sum_k A[ty + 16 i][k] * Bm[k][4 tx + j]       (the thread owns 4 CONSECUTIVE output columns)
__device__ __forceinline__ void as_mm_ab(const float* A, const float* Bm, int tx, int ty, float (&acc)[4][4]) {
  as_zero(acc);
#pragma unroll 2
  for (int k = 0; k < 64; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = as_ld4(A + (ty + 16 * i) * AS_LD + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) b[kk] = as_ld4(Bm + (k + kk) * AS_LD + 4 * tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[i][0] = fmaf(a[i].w, b[3].x, fmaf(a[i].z, b[2].x, fmaf(a[i].y, b[1].x, fmaf(a[i].x, b[0].x, acc[i][0]))));
      acc[i][1] = fmaf(a[i].w, b[3].y, fmaf(a[i].z, b[2].y, fmaf(a[i].y, b[1].y, fmaf(a[i].x, b[0].y, acc[i][1]))));
      acc[i][2] = fmaf(a[i].w, b[3].z, fmaf(a[i].z, b[2].z, fmaf(a[i].y, b[1].z, fmaf(a[i].x, b[0].z, acc[i][2]))));
      acc[i][3] = fmaf(a[i].w, b[3].w, fmaf(a[i].z, b[2].w, fmaf(a[i].y, b[1].w, fmaf(a[i].x, b[0].w, acc[i][3]))));
    }
  }
}
// "A^T B": acc[i][j] = sum_k A[k][4 ty + i] * Bm[k][4 tx + j]       (4 consecutive output rows and columns)
__device__ __forceinline__ void as_mm_atb(const float* A, const float* Bm, int tx, int ty, float (&acc)[4][4]) {
  as_zero(acc);
#pragma unroll 4
  for (int k = 0; k < 64; ++k) {
    const float4 a = as_ld4(A + k * AS_LD + 4 * ty), b = as_ld4(Bm + k * AS_LD + 4 * tx);
    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[i][0] = fmaf(av[i], b.x, acc[i][0]); acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
      acc[i][2] = fmaf(av[i], b.z, acc[i][2]); acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
    }
  }
}
// sum / max over the 16 lanes that share ty (one row of the 16 x 16 thread grid = half a warp)
__device__ __forceinline__ float as_row_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float as_row_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// acc = (q k^T) / sqrt(ch)  ->  row softmax in place (fp32, like the reference's softmax(weight.float())); the thread holds columns
// tx + 16 j of rows ty + 16 i, the other columns of a row are in the 15 other lanes of its half-warp
__device__ __forceinline__ void as_softmax_rows(float (&acc)[4][4], float alpha) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] *= alpha; m = fmaxf(m, acc[i][j]); }
    m = as_row_max(m);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = expf(acc[i][j] - m); sum += acc[i][j]; }
    sum = as_row_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] *= inv;
  }
}
// interleaved-column results (rows ty + 16 i, columns tx + 16 j) -> shared tile
__device__ __forceinline__ void as_put(float* T, int tx, int ty, const float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) T[(ty + 16 * i) * AS_LD + tx + 16 * j] = acc[i][j];
}
// rows r0 + rs i (tokens), columns 4 tx .. 4 tx + 3 (channels of one head slice) -> global, ld floats between tokens
__device__ __forceinline__ void as_store4(float* __restrict__ dst, long ld, int tx, int r0, int rs, const float (&acc)[4][4], float scale) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(dst + (long)(r0 + rs * i) * ld + 4 * tx) =
        make_float4(acc[i][0] * scale, acc[i][1] * scale, acc[i][2] * scale, acc[i][3] * scale);
}

// grid (heads, B), 256 threads, 4 tiles of dynamic shared memory
__global__ void __launch_bounds__(256) attn_small_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, int C, int heads) {
  extern __shared__ __align__(16) float as_sm[];
  float *Q = as_sm, *K = Q + AS_TILE, *V = K + AS_TILE, *P = V + AS_TILE;
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long C3 = 3L * C;
  const float* base = qkv + (long)b * AS_L * C3 + 3L * AS_CH * h;
  {
    float* const dst[3] = {Q, K, V};
    const float* const src[3] = {base, base + AS_CH, base + 2 * AS_CH};
    const long ld[3] = {C3, C3, C3};
    as_load_tiles<3>(dst, src, ld);
  }
  __syncthreads();
  float acc[4][4];
  as_mm_abt(Q, K, tx, ty, acc);
  as_softmax_rows(acc, 1.0f / sqrtf((float)AS_CH));
  as_put(P, tx, ty, acc);
  __syncthreads();
  as_mm_ab(P, V, tx, ty, acc);           // rows ty + 16 i, channels 4 tx ..
  as_store4(out + (long)b * AS_L * C + (long)AS_CH * h, C, tx, ty, 16, acc, 1.0f);
  pdl_launch_dependents();
}

// grid (heads, B), 256 threads, 6 tiles: q, k, v, dO, P, dS
__global__ void __launch_bounds__(256) attn_small_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ g_out,
                                                             float* __restrict__ g_qkv, int C, int heads) {
  extern __shared__ __align__(16) float as_sm[];
  float *Q = as_sm, *K = Q + AS_TILE, *V = K + AS_TILE, *G = V + AS_TILE, *P = G + AS_TILE, *S = P + AS_TILE;
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long C3 = 3L * C;
  const float* base = qkv + (long)b * AS_L * C3 + 3L * AS_CH * h;
  {
    float* const dst[4] = {Q, K, V, G};
    const float* const src[4] = {base, base + AS_CH, base + 2 * AS_CH, g_out + (long)b * AS_L * C + (long)AS_CH * h};
    const long ld[4] = {C3, C3, C3, (long)C};
    as_load_tiles<4>(dst, src, ld);
  }
  __syncthreads();
  const float alpha = 1.0f / sqrtf((float)AS_CH);
  float p[4][4], d[4][4];
  as_mm_abt(Q, K, tx, ty, p);
  as_softmax_rows(p, alpha);
  as_mm_abt(G, V, tx, ty, d);            // dP[t][s] = sum_c dO[t][c] V[s][c]
#pragma unroll
  for (int i = 0; i < 4; ++i) {          // dS = P o (dP - rowsum(P o dP))
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) dot = fmaf(p[i][j], d[i][j], dot);
    dot = as_row_sum(dot);
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = p[i][j] * (d[i][j] - dot);
  }
  as_put(P, tx, ty, p);
  as_put(S, tx, ty, d);
  __syncthreads();
  float* gb = g_qkv + (long)b * AS_L * C3 + 3L * AS_CH * h;
  float acc[4][4];
  as_mm_atb(P, G, tx, ty, acc);          // dV[s][c] = sum_t P[t][s] dO[t][c]: rows 4 ty + i, channels 4 tx ..
  as_store4(gb + 2 * AS_CH, C3, tx, 4 * ty, 1, acc, 1.0f);
  as_mm_ab(S, K, tx, ty, acc);           // dQ[t][c] = alpha sum_s dS[t][s] K[s][c]: rows ty + 16 i
  as_store4(gb, C3, tx, ty, 16, acc, alpha);
  as_mm_atb(S, Q, tx, ty, acc);          // dK[s][c] = alpha sum_t dS[t][s] Q[t][c]: rows 4 ty + i
  as_store4(gb + AS_CH, C3, tx, 4 * ty, 1, acc, alpha);
  pdl_launch_dependents();
}

static int attention_small_fwd(const float* qkv, float* out, int B, int C, int heads, cudaStream_t s) {
  constexpr int SMEM = 4 * AS_TILE * (int)sizeof(float);
  static bool done = false;
  if (!done) { OSM_CUDA_CHECK(cudaFuncSetAttribute(attn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); done = true; }
  OSM_LAUNCH_PDL("attn_small_fwd_kernel", attn_small_fwd_kernel, dim3(heads, B), dim3(256), SMEM, s, qkv, out, C, heads);
  return OSM_OK;
}
static int attention_small_bwd(const float* qkv, const float* g_out, float* g_qkv, int B, int C, int heads, cudaStream_t s) {
  constexpr int SMEM = 6 * AS_TILE * (int)sizeof(float);
  static bool done = false;
  if (!done) { OSM_CUDA_CHECK(cudaFuncSetAttribute(attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); done = true; }
  OSM_LAUNCH_PDL("attn_small_bwd_kernel", attn_small_bwd_kernel, dim3(heads, B), dim3(256), SMEM, s, qkv, g_out, g_qkv, C, heads);
  return OSM_OK;
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
  if (attention_small_ok(L, C, heads)) return attention_small_fwd(qkv, out, B, C, heads, s);
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
  if (attention_small_ok(L, C, heads)) return attention_small_bwd(qkv, g_out, g_qkv, B, C, heads, s);
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
