// Fused QKVAttentionLegacy forward / backward on the 5th-gen tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
//
// Reference: unet.py:407-437 (QKVAttentionLegacy.forward: w = softmax_fp32((q ch^-1/4)^T (k ch^-1/4)); a = w v^T) and its
// autograd backward, which the reference reaches through the checkpointed AttentionBlock (unet.py:376, nn.py:124-170).
//
// Layout: qkv token-major [B, L, 3C]; head h owns channels [192 h, 192 h + 192) as (q | k | v) with ch = 64 - exactly the
// reference's `qkv.reshape(bs * n_heads, ch * 3, length).split(ch, dim=1)`.  Every GEMM of the block has the shape
// M = 128 (one row tile of tokens) x N = 64 x K = 64, so one instruction shape serves all of them:
//
//   forward      S  = Q K^T          P = exp2(S c - m)      O  += P V            (online softmax, O rescaled in registers)
//   dQ kernel    S  = Q K^T, dP = dO V^T,  dS = P o (dP - D),  dQ += dS K        (D = rowsum(dO o O), P from the saved LSE)
//   dK/dV kernel S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q          (row tile = keys, streamed chunks = queries)
//
// The L x L score / probability matrices never leave the SM: S lands in TMEM, one thread per row reads it with
// tcgen05.ld, and writes P (TF32-rounded) straight into the 128-byte-swizzled K-major shared-memory tile the next
// tcgen05.mma consumes.  tcgen05 reads TF32 operands only K-major (the MN-major form needs a different swizzle), so the
// operands whose contraction runs over tokens (V in P V, K in dS K, dO in P^T dO, Q in dS^T Q) come from channel-major
// copies qkvT [B, 3C, L] / dOT [B, C, L] written by a small transpose kernel.
//
// Pipeline per CTA (128 threads = 128 rows; thread 0 also issues TMA and MMA): the MMAs that depend only on freshly
// loaded tiles (S, dP of chunk j+1) are issued together with the accumulating MMAs of chunk j and tracked by ONE
// tcgen05.commit, so there is a single tensor-core round trip per 64-token chunk; the TMA loads of the next chunk are
// issued at the start of the iteration and land while the threads do the softmax arithmetic.
// No atomics anywhere: dQ and dK/dV are produced by separate kernels, each owning its output rows - bit-reproducible.
#include <math.h>
#include <stdlib.h>

#include "tcgen05.cuh"

namespace osm {

constexpr int FA_CH = 64;  // channels per head (num_head_channels = 64 in every shipped config)
constexpr int FA_CK = 64;  // tokens per streamed chunk
constexpr uint32_t FA_IDESC = make_idesc_tf32(128, 64);
constexpr uint32_t FA_ROWTILE_BYTES = 128 * FA_CH * 4;  // 32 KB: a (<=128)-row token-major tile of one head slice
constexpr uint32_t FA_CHUNK_BYTES = FA_CK * FA_CH * 4;  // 16 KB

struct FlashParams {
  int B, L, C, heads, R;  // R = min(128, L): rows per row tile
  float sl2;              // log2(e) / sqrt(ch)
  float scale;            // 1 / sqrt(ch)
  float* out;             // fwd: attention output a [B, L, C]
  const float* O;         // bwd: the saved forward output
  float* lse;             // [B, heads, L]: log2-domain log-sum-exp of the scaled scores
  float* Dv;              // [B, heads, L]: rowsum(dO o O)
  const float* dO;        // [B, L, C]
  float* g_qkv;           // [B, L, 3C]
  int split;              // 1, or 2: the streamed chunks of a row tile are shared by the two CTAs of a (2,1,1) cluster (see fa_split)
};

// Small grids (batch 1: 64 row-tile CTAs at L = 1024, each walking 16 chunks in a serial MMA -> softmax -> MMA chain) leave more than
// half of the SMs idle while the chain length sets the kernel time.  With split == 2 the two CTAs of a cluster take the first and the
// second half of the streamed chunks of the SAME row tile and rank 0 folds in rank 1's partial result through distributed shared
// memory at the end (own part first, then the peer's: a fixed order, still no atomics).  Rows are staged as 16 float4 per row with
// the float4 index XOR-ed by the row (conflict-free for the 256-byte row stride).
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_saddr, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(cta));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t fa_stage_addr(uint32_t region, int row, int f4) {   // float4 slot f4 (0..15) of a 64-float row
  return region + (uint32_t)row * 256u + (uint32_t)((f4 ^ (row & 15)) << 4);
}
__device__ __forceinline__ void fa_stage_put32(uint32_t region, int row, int half, const uint32_t (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(fa_stage_addr(region, row, 8 * half + i)), "r"(v[4 * i]), "r"(v[4 * i + 1]),
                 "r"(v[4 * i + 2]), "r"(v[4 * i + 3])
                 : "memory");
}
__device__ __forceinline__ void fa_stage_add32(uint32_t region, int row, int half, uint32_t (&v)[32]) {   // v += peer (rank 1) values
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 q = ld_dsmem_f4(fa_stage_addr(region, row, 8 * half + i), 1u);
    v[4 * i] = __float_as_uint(__uint_as_float(v[4 * i]) + q.x);
    v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + q.y);
    v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + q.z);
    v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + q.w);
  }
}

// 8 MMAs (K = 8 each) covering a 64-deep contraction.  a_kb / b_kb: bytes between the two 32-element K blocks.
__device__ __forceinline__ void fa_mma64(uint32_t d_tmem, uint32_t sa, uint32_t a_kb, uint32_t sb, uint32_t b_kb, bool accumulate) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    mma_tf32(d_tmem, make_smem_desc(sa + (uint32_t)(i >> 2) * a_kb + (uint32_t)(i & 3) * 32u),
             make_smem_desc(sb + (uint32_t)(i >> 2) * b_kb + (uint32_t)(i & 3) * 32u), FA_IDESC, (uint32_t)(accumulate || i > 0));
}
// token-major tile [rows x 64 ch] = two K blocks of [rows x 32 ch]
__device__ __forceinline__ void fa_load_tok(uint32_t dst, const CUtensorMap* m, uint32_t bar, int ch0, int row0, int b, int rows) {
  tma_load_3d(dst, m, bar, ch0, row0, b);
  tma_load_3d(dst + (uint32_t)rows * 128u, m, bar, ch0 + 32, row0, b);
}
// channel-major tile [64 ch x 64 tokens] = two K blocks of [64 ch x 32 tokens]
__device__ __forceinline__ void fa_load_chan(uint32_t dst, const CUtensorMap* m, uint32_t bar, int tok0, int ch0, int b) {
  tma_load_3d(dst, m, bar, tok0, ch0, b);
  tma_load_3d(dst + 8192u, m, bar, tok0 + 32, ch0, b);
}
// one row (64 values, TF32-rounded) of a [128 x 64] K-major SWIZZLE_128B operand tile
__device__ __forceinline__ void fa_store_row(uint32_t tile, int row, const uint32_t (&v)[64]) {
  const uint32_t base = tile + (uint32_t)row * 128u, swz = (uint32_t)row & 7u;
#pragma unroll
  for (int kb = 0; kb < 2; ++kb)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int j = kb * 32 + c * 4;
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)kb * 16384u + (((uint32_t)c ^ swz) << 4)), "r"(v[j]),
                   "r"(v[j + 1]), "r"(v[j + 2]), "r"(v[j + 3])
                   : "memory");
    }
}
// 32 values (one 32-token K block) of one row of a K-major SWIZZLE_128B operand tile; kblock = tile + 16384 * block index
__device__ __forceinline__ void fa_store_row32(uint32_t kblock, int row, const uint32_t (&v)[32]) {
  const uint32_t base = kblock + (uint32_t)row * 128u, swz = (uint32_t)row & 7u;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (((uint32_t)c ^ swz) << 4)), "r"(v[4 * c]), "r"(v[4 * c + 1]),
                 "r"(v[4 * c + 2]), "r"(v[4 * c + 3])
                 : "memory");
}
__device__ __forceinline__ void fa_tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
  tmem_ld32(taddr + 32u, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
}
__device__ __forceinline__ float fa_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#define FA_PROLOGUE(NBARS, TMEM_COLS)                                                                                       \
  extern __shared__ uint8_t smem_raw[];                                                                                     \
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;                                                         \
  __shared__ __align__(8) uint64_t bars[NBARS];                                                                             \
  __shared__ uint32_t tmem_base_smem;                                                                                       \
  const int tid = threadIdx.x, warp = tid >> 5;                                                                             \
  pdl_launch_dependents();                                                                                                  \
  if (tid == 0) {                                                                                                           \
    for (int i = 0; i < NBARS; ++i) mbar_init(smem_u32(&bars[i]), 1);                                                       \
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");                                                      \
  }                                                                                                                         \
  if (warp == 1) {                                                                                                          \
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),      \
                 "r"(TMEM_COLS)                                                                                             \
                 : "memory");                                                                                               \
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");                                \
  }                                                                                                                         \
  tcgen05_fence_before();                                                                                                   \
  __syncthreads();                                                                                                          \
  tcgen05_fence_after();                                                                                                    \
  const uint32_t tmem_base = tmem_base_smem;                                                                                \
  const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16); /* this warp's TMEM lane quadrant */        \
  pdl_wait();

#define FA_EPILOGUE(TMEM_COLS)                                                                                              \
  tcgen05_fence_before();                                                                                                   \
  __syncthreads();                                                                                                          \
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");

// ------------------------------------------------------------------------------------------------
// forward: grid (L / R, heads, B)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
flash_fwd_kernel(const __grid_constant__ CUtensorMap tmTokR, const __grid_constant__ CUtensorMap tmTok64,
                 const __grid_constant__ CUtensorMap tmChan, const FlashParams p) {
  FA_PROLOGUE(4, 128)
  const uint32_t barQ = smem_u32(&bars[0]), barK = smem_u32(&bars[1]), barV = smem_u32(&bars[2]), barM = smem_u32(&bars[3]);
  const uint32_t sQ = smem_base, sK = sQ + FA_ROWTILE_BYTES, sVT = sK + FA_CHUNK_BYTES, sP = sVT + FA_CHUNK_BYTES;
  __shared__ float s_ml[2][128];
  const int rank = p.split > 1 ? (int)cluster_ctarank() : 0;
  const int q0 = (blockIdx.x / p.split) * p.R, h = blockIdx.y, b = blockIdx.z;
  const int cq = 3 * FA_CH * h, n_chunks = p.L / FA_CK / p.split, j0 = rank * n_chunks;
  const uint32_t tS = tmem_base, tO = tmem_base + 64u;
  const uint32_t a_kb = (uint32_t)p.R * 128u;

  if (tid == 0) {
    mbar_expect_tx(barQ, (uint32_t)p.R * 256u + FA_CHUNK_BYTES);
    fa_load_tok(sQ, &tmTokR, barQ, cq, q0, b, p.R);
    fa_load_tok(sK, &tmTok64, barQ, cq + FA_CH, j0 * FA_CK, b, FA_CK);
    mbar_wait(barQ, 0);
    tcgen05_fence_after();
    fa_mma64(tS, sQ, a_kb, sK, 8192u, false);
    tcgen05_commit(barM);
  }
  float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
  float o[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) o[i] = 0.f;

  for (int j = 0; j < n_chunks; ++j) {
    mbar_wait(barM, (uint32_t)j & 1u);
    tcgen05_fence_after();
    if (tid == 0) {  // the K buffer (read by S(j)) and the V^T buffer (read by PV(j-1)) are free again
      if (j + 1 < n_chunks) {
        mbar_expect_tx(barK, FA_CHUNK_BYTES);
        fa_load_tok(sK, &tmTok64, barK, cq + FA_CH, (j0 + j + 1) * FA_CK, b, FA_CK);
      }
      mbar_expect_tx(barV, FA_CHUNK_BYTES);
      fa_load_chan(sVT, &tmChan, barV, (j0 + j) * FA_CK, cq + 2 * FA_CH, b);
    }
    if (j > 0) {  // fold the previous chunk's P V (computed against the running max of that chunk)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t r[32];
        tmem_ld32(tmem_row + (tO - tmem_base) + (uint32_t)hf * 32u, r);
#pragma unroll
        for (int i = 0; i < 32; ++i) o[hf * 32 + i] = o[hf * 32 + i] * alpha_prev + __uint_as_float(r[i]);
      }
    }
    uint32_t s[64];
    fa_tmem_ld64(tmem_row + (tS - tmem_base), s);
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const float t = __uint_as_float(s[i]) * p.sl2;
      s[i] = __float_as_uint(t);
      mx = fmaxf(mx, t);
    }
    const float m_new = fmaxf(m, mx);
    const float alpha = fa_exp2(m - m_new);
    float rs = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const uint32_t pr = f32_to_tf32_rn(fa_exp2(__uint_as_float(s[i]) - m_new));
      s[i] = pr;
      rs += __uint_as_float(pr);
    }
    l = l * alpha + rs;
    m = m_new;
    alpha_prev = alpha;
    fa_store_row(sP, tid, s);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      mbar_wait(barV, (uint32_t)j & 1u);
      tcgen05_fence_after();
      fa_mma64(tO, sP, 16384u, sVT, 8192u, false);
      if (j + 1 < n_chunks) {
        mbar_wait(barK, (uint32_t)j & 1u);
        tcgen05_fence_after();
        fa_mma64(tS, sQ, a_kb, sK, 8192u, false);
      }
      tcgen05_commit(barM);
    }
  }
  mbar_wait(barM, (uint32_t)n_chunks & 1u);
  tcgen05_fence_after();
  const int row = q0 + tid;
  const bool row_ok = tid < p.R && row < p.L;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {   // fold the last chunk's P V: o = un-normalised output of this CTA's chunks
    uint32_t r[32];
    tmem_ld32(tmem_row + (tO - tmem_base) + (uint32_t)hf * 32u, r);
#pragma unroll
    for (int i = 0; i < 32; ++i) o[hf * 32 + i] = o[hf * 32 + i] * alpha_prev + __uint_as_float(r[i]);
  }
  if (p.split > 1) {
    // merge the two halves of the key range: (m, l, o) of rank 1 -> rank 0 (the P tile's shared memory is free: every MMA has completed)
    if (rank == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(fa_stage_addr(sP, tid, i)), "f"(o[4 * i]), "f"(o[4 * i + 1]), "f"(o[4 * i + 2]),
                     "f"(o[4 * i + 3])
                     : "memory");
      s_ml[0][tid] = m; s_ml[1][tid] = l;
    }
    cluster_sync_all();
    if (rank == 0) {
      const float m1 = ld_dsmem_f32(smem_u32(&s_ml[0][tid]), 1u), l1 = ld_dsmem_f32(smem_u32(&s_ml[1][tid]), 1u);
      const float ms = fmaxf(m, m1), w0 = fa_exp2(m - ms), w1 = fa_exp2(m1 - ms);
      l = l * w0 + l1 * w1;
      m = ms;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 q = ld_dsmem_f4(fa_stage_addr(sP, tid, i), 1u);
        o[4 * i] = o[4 * i] * w0 + q.x * w1; o[4 * i + 1] = o[4 * i + 1] * w0 + q.y * w1;
        o[4 * i + 2] = o[4 * i + 2] * w0 + q.z * w1; o[4 * i + 3] = o[4 * i + 3] * w0 + q.w * w1;
      }
    }
    cluster_sync_all();   // rank 1 keeps its shared memory alive until rank 0 has read it
  }
  if (rank == 0 && row_ok) {
    const float inv = 1.0f / l;
    float* dst = p.out + ((size_t)b * p.L + row) * p.C + FA_CH * h;
#pragma unroll
    for (int i = 0; i < 64; i += 4)
      *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
    p.lse[((size_t)b * p.heads + h) * p.L + row] = m + log2f(l);
  }
  FA_EPILOGUE(128)
}

// ------------------------------------------------------------------------------------------------
// backward 1: dQ (and D = rowsum(dO o O)).  grid (L / R, heads, B), 256 threads: TWO threads per row, thread (row, half)
// owns the 32-token K block `half` of every 64-token chunk (there is no row reduction inside the loop - P comes from the
// saved LSE - so the split is free, and eight warps keep the four schedulers busy where four ran at ~1/4 IPC).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
flash_dq_kernel(const __grid_constant__ CUtensorMap tmTokR, const __grid_constant__ CUtensorMap tmTok64,
                const __grid_constant__ CUtensorMap tmChan, const __grid_constant__ CUtensorMap tmDoR, const FlashParams p) {
  FA_PROLOGUE(4, 256)
  __shared__ float s_Dh[2][128];
  const uint32_t barT = smem_u32(&bars[0]), barKV = smem_u32(&bars[1]), barKT = smem_u32(&bars[2]), barM = smem_u32(&bars[3]);
  const uint32_t sQ = smem_base, sdO = sQ + FA_ROWTILE_BYTES, sK = sdO + FA_ROWTILE_BYTES, sV = sK + FA_CHUNK_BYTES,
                 sKT = sV + FA_CHUNK_BYTES, sdS = sKT + FA_CHUNK_BYTES;
  const int rank = p.split > 1 ? (int)cluster_ctarank() : 0;
  const int q0 = (blockIdx.x / p.split) * p.R, h = blockIdx.y, b = blockIdx.z;
  const int cq = 3 * FA_CH * h, n_chunks = p.L / FA_CK / p.split, j0 = rank * n_chunks;
  const uint32_t tS = 0u, tdP = 64u, tdQ = 128u;  // column offsets
  const uint32_t a_kb = (uint32_t)p.R * 128u;
  const int r = tid & 127, half = tid >> 7;       // tile row, K-block / column half

  if (tid == 0) {
    mbar_expect_tx(barT, 2u * (uint32_t)p.R * 256u + 2u * FA_CHUNK_BYTES);
    fa_load_tok(sQ, &tmTokR, barT, cq, q0, b, p.R);
    fa_load_tok(sdO, &tmDoR, barT, FA_CH * h, q0, b, p.R);
    fa_load_tok(sK, &tmTok64, barT, cq + FA_CH, j0 * FA_CK, b, FA_CK);
    fa_load_tok(sV, &tmTok64, barT, cq + 2 * FA_CH, j0 * FA_CK, b, FA_CK);
    mbar_wait(barT, 0);
    tcgen05_fence_after();
    fa_mma64(tmem_base + tS, sQ, a_kb, sK, 8192u, false);
    fa_mma64(tmem_base + tdP, sdO, a_kb, sV, 8192u, false);
    tcgen05_commit(barM);
  }
  const int row = q0 + r;
  const bool row_ok = r < p.R && row < p.L;
  float Dr = 0.f, lse = 0.f;
  if (row_ok) {
    const float4* a = reinterpret_cast<const float4*>(p.dO + ((size_t)b * p.L + row) * p.C + FA_CH * h + 32 * half);
    const float4* c = reinterpret_cast<const float4*>(p.O + ((size_t)b * p.L + row) * p.C + FA_CH * h + 32 * half);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 x = a[i], y = c[i];
      Dr += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
    lse = p.lse[((size_t)b * p.heads + h) * p.L + row];
  }
  s_Dh[half][r] = Dr;
  __syncthreads();
  Dr = s_Dh[0][r] + s_Dh[1][r];   // fixed order: both threads of a row hold the same D
  if (row_ok && half == 0 && rank == 0) p.Dv[((size_t)b * p.heads + h) * p.L + row] = Dr;

  for (int j = 0; j < n_chunks; ++j) {
    mbar_wait(barM, (uint32_t)j & 1u);
    tcgen05_fence_after();
    if (tid == 0) {
      if (j + 1 < n_chunks) {
        mbar_expect_tx(barKV, 2u * FA_CHUNK_BYTES);
        fa_load_tok(sK, &tmTok64, barKV, cq + FA_CH, (j0 + j + 1) * FA_CK, b, FA_CK);
        fa_load_tok(sV, &tmTok64, barKV, cq + 2 * FA_CH, (j0 + j + 1) * FA_CK, b, FA_CK);
      }
      mbar_expect_tx(barKT, FA_CHUNK_BYTES);
      fa_load_chan(sKT, &tmChan, barKT, (j0 + j) * FA_CK, cq + FA_CH, b);
    }
    uint32_t s[32], d[32];
    tmem_ld32(tmem_row + tS + 32u * (uint32_t)half, s);
    tmem_ld32(tmem_row + tdP + 32u * (uint32_t)half, d);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float pr = fa_exp2(__uint_as_float(s[i]) * p.sl2 - lse);
      s[i] = f32_to_tf32_rn(pr * (__uint_as_float(d[i]) - Dr));
    }
    fa_store_row32(sdS + 16384u * (uint32_t)half, r, s);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      mbar_wait(barKT, (uint32_t)j & 1u);
      tcgen05_fence_after();
      fa_mma64(tmem_base + tdQ, sdS, 16384u, sKT, 8192u, j > 0);
      if (j + 1 < n_chunks) {
        mbar_wait(barKV, (uint32_t)j & 1u);
        tcgen05_fence_after();
        fa_mma64(tmem_base + tS, sQ, a_kb, sK, 8192u, false);
        fa_mma64(tmem_base + tdP, sdO, a_kb, sV, 8192u, false);
      }
      tcgen05_commit(barM);
    }
  }
  mbar_wait(barM, (uint32_t)n_chunks & 1u);
  tcgen05_fence_after();
  {
    float* dst = p.g_qkv + ((size_t)b * p.L + row) * (3 * p.C) + cq + 32 * half;
    uint32_t v[32];
    tmem_ld32(tmem_row + tdQ + 32u * (uint32_t)half, v);
    if (p.split > 1) {   // dQ = this CTA's keys + the peer's (the dS tile's shared memory is free: every MMA has completed)
      if (rank == 1) fa_stage_put32(sdS, r, half, v);
      cluster_sync_all();
      if (rank == 0) fa_stage_add32(sdS, r, half, v);
      cluster_sync_all();
    }
    if (row_ok && rank == 0) {
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(v[i]) * p.scale, __uint_as_float(v[i + 1]) * p.scale,
                                                          __uint_as_float(v[i + 2]) * p.scale, __uint_as_float(v[i + 3]) * p.scale);
    }
  }
  FA_EPILOGUE(256)
}

// ------------------------------------------------------------------------------------------------
// backward 2: dK, dV.  Row tile = keys, streamed chunks = queries.  grid (L / R, heads, B), 256 threads (two per row, as above)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
flash_dkv_kernel(const __grid_constant__ CUtensorMap tmTokR, const __grid_constant__ CUtensorMap tmTok64,
                 const __grid_constant__ CUtensorMap tmChan, const __grid_constant__ CUtensorMap tmDo64,
                 const __grid_constant__ CUtensorMap tmDoChan, const FlashParams p) {
  FA_PROLOGUE(4, 256)
  __shared__ float s_lse[FA_CK], s_D[FA_CK];
  const uint32_t barT = smem_u32(&bars[0]), barC = smem_u32(&bars[1]), barCT = smem_u32(&bars[2]), barM = smem_u32(&bars[3]);
  const uint32_t sK = smem_base, sV = sK + FA_ROWTILE_BYTES, sQc = sV + FA_ROWTILE_BYTES, sdOc = sQc + FA_CHUNK_BYTES,
                 sQT = sdOc + FA_CHUNK_BYTES, sdOT = sQT + FA_CHUNK_BYTES, sPT = sdOT + FA_CHUNK_BYTES, sdST = sPT + 32768u;
  const int rank = p.split > 1 ? (int)cluster_ctarank() : 0;
  const int k0 = (blockIdx.x / p.split) * p.R, h = blockIdx.y, b = blockIdx.z;
  const int cq = 3 * FA_CH * h, n_chunks = p.L / FA_CK / p.split, j0 = rank * n_chunks;
  const uint32_t tST = 0u, tdPT = 64u, tdV = 128u, tdK = 192u;
  const uint32_t a_kb = (uint32_t)p.R * 128u;
  const int r = tid & 127, half = tid >> 7;

  if (tid == 0) {
    mbar_expect_tx(barT, 2u * (uint32_t)p.R * 256u + 2u * FA_CHUNK_BYTES);
    fa_load_tok(sK, &tmTokR, barT, cq + FA_CH, k0, b, p.R);
    fa_load_tok(sV, &tmTokR, barT, cq + 2 * FA_CH, k0, b, p.R);
    fa_load_tok(sQc, &tmTok64, barT, cq, j0 * FA_CK, b, FA_CK);
    fa_load_tok(sdOc, &tmDo64, barT, FA_CH * h, j0 * FA_CK, b, FA_CK);
    mbar_wait(barT, 0);
    tcgen05_fence_after();
    fa_mma64(tmem_base + tST, sK, a_kb, sQc, 8192u, false);
    fa_mma64(tmem_base + tdPT, sV, a_kb, sdOc, 8192u, false);
    tcgen05_commit(barM);
  }
  const float* lse_g = p.lse + ((size_t)b * p.heads + h) * p.L;
  const float* D_g = p.Dv + ((size_t)b * p.heads + h) * p.L;

  for (int j = 0; j < n_chunks; ++j) {
    mbar_wait(barM, (uint32_t)j & 1u);
    tcgen05_fence_after();
    if (tid == 0) {
      if (j + 1 < n_chunks) {
        mbar_expect_tx(barC, 2u * FA_CHUNK_BYTES);
        fa_load_tok(sQc, &tmTok64, barC, cq, (j0 + j + 1) * FA_CK, b, FA_CK);
        fa_load_tok(sdOc, &tmDo64, barC, FA_CH * h, (j0 + j + 1) * FA_CK, b, FA_CK);
      }
      mbar_expect_tx(barCT, 2u * FA_CHUNK_BYTES);
      fa_load_chan(sQT, &tmChan, barCT, (j0 + j) * FA_CK, cq, b);
      fa_load_chan(sdOT, &tmDoChan, barCT, (j0 + j) * FA_CK, FA_CH * h, b);
    }
    if (tid < FA_CK) {
      s_lse[tid] = lse_g[(j0 + j) * FA_CK + tid];
      s_D[tid] = D_g[(j0 + j) * FA_CK + tid];
    }
    __syncthreads();
    uint32_t s[32], d[32];
    tmem_ld32(tmem_row + tST + 32u * (uint32_t)half, s);
    tmem_ld32(tmem_row + tdPT + 32u * (uint32_t)half, d);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float pr = fa_exp2(__uint_as_float(s[i]) * p.sl2 - s_lse[32 * half + i]);
      s[i] = f32_to_tf32_rn(pr);
      d[i] = f32_to_tf32_rn(pr * (__uint_as_float(d[i]) - s_D[32 * half + i]));
    }
    fa_store_row32(sPT + 16384u * (uint32_t)half, r, s);
    fa_store_row32(sdST + 16384u * (uint32_t)half, r, d);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      mbar_wait(barCT, (uint32_t)j & 1u);
      tcgen05_fence_after();
      fa_mma64(tmem_base + tdV, sPT, 16384u, sdOT, 8192u, j > 0);
      fa_mma64(tmem_base + tdK, sdST, 16384u, sQT, 8192u, j > 0);
      if (j + 1 < n_chunks) {
        mbar_wait(barC, (uint32_t)j & 1u);
        tcgen05_fence_after();
        fa_mma64(tmem_base + tST, sK, a_kb, sQc, 8192u, false);
        fa_mma64(tmem_base + tdPT, sV, a_kb, sdOc, 8192u, false);
      }
      tcgen05_commit(barM);
    }
  }
  mbar_wait(barM, (uint32_t)n_chunks & 1u);
  tcgen05_fence_after();
  const int row = k0 + r;
  const bool row_ok = r < p.R && row < p.L;
  float* dst = p.g_qkv + ((size_t)b * p.L + row) * (3 * p.C) + cq + 32 * half;
#pragma unroll
  for (int part = 0; part < 2; ++part) {  // 0: dK (scaled), 1: dV
    const float sc = part == 0 ? p.scale : 1.0f;
    uint32_t v[32];
    tmem_ld32(tmem_row + (part == 0 ? tdK : tdV) + 32u * (uint32_t)half, v);
    if (p.split > 1) {   // this CTA's queries + the peer's (the P^T / dS^T tiles' shared memory is free: every MMA has completed)
      const uint32_t region = part == 0 ? sPT : sdST;
      if (rank == 1) fa_stage_put32(region, r, half, v);
      cluster_sync_all();
      if (rank == 0) fa_stage_add32(region, r, half, v);
      cluster_sync_all();
    }
    if (row_ok && rank == 0) {
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(dst + FA_CH * (1 + part) + i) =
            make_float4(__uint_as_float(v[i]) * sc, __uint_as_float(v[i + 1]) * sc, __uint_as_float(v[i + 2]) * sc,
                        __uint_as_float(v[i + 3]) * sc);
    }
  }
  FA_EPILOGUE(256)
}

// token-major [B, L, Cs] -> channel-major [B, Cs, L]; optional TF32 rounding of the copy is not needed (the tensor core
// truncates both copies identically).  grid (L / 32, Cs / 32, B), block (32, 8)
__global__ void tok_to_chan_kernel(const float* __restrict__ src, float* __restrict__ dst, int L, int Cs) {
  __shared__ float t[32][33];
  pdl_launch_dependents();
  pdl_wait();
  const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32, b = blockIdx.z;
  const float* s = src + (size_t)b * L * Cs;
  float* d = dst + (size_t)b * L * Cs;
#pragma unroll
  for (int i = threadIdx.y; i < 32; i += 8) t[i][threadIdx.x] = s[(size_t)(l0 + i) * Cs + c0 + threadIdx.x];
  __syncthreads();
#pragma unroll
  for (int i = threadIdx.y; i < 32; i += 8) d[(size_t)(c0 + i) * L + l0 + threadIdx.x] = t[threadIdx.x][i];
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
bool attn_flash_supported(int L, int C, int heads) { return heads > 0 && C == heads * FA_CH && L >= 64 && L % 64 == 0; }

static int encode3(void* map, const float* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1, const char* what) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(OSM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (needs a CUDA 12 driver)");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 4, d0 * d1 * 4};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc((CUtensorMap*)map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(OSM_ERR_CUDA, std::string("cuTensorMapEncodeTiled(") + what + ") failed: code " + std::to_string((int)r));
  return OSM_OK;
}

int attn_flash_plan(AttnFlashPlan* pl, const float* qkv, float* qkvT, float* O, float* lse, float* Dv, const float* dO, float* dOT,
                    float* g_qkv, int B, int L, int C, int heads) {
  if (!attn_flash_supported(L, C, heads)) return fail(OSM_ERR_INVALID, "flash attention: needs 64 channels per head and L % 64 == 0");
  pl->qkv = qkv; pl->qkvT = qkvT; pl->O = O; pl->lse = lse; pl->Dv = Dv; pl->dO = dO; pl->dOT = dOT; pl->g_qkv = g_qkv;
  pl->B = B; pl->L = L; pl->C = C; pl->heads = heads; pl->R = L < 128 ? L : 128;
  const uint64_t C3 = 3ull * C;
  if (int e = encode3(pl->tm[0], qkv, C3, L, B, 32, (uint32_t)pl->R, "qkv row tile")) return e;
  if (int e = encode3(pl->tm[1], qkv, C3, L, B, 32, FA_CK, "qkv chunk")) return e;
  if (int e = encode3(pl->tm[2], qkvT, L, C3, B, 32, FA_CH, "qkvT chunk")) return e;
  if (dO) {
    if (int e = encode3(pl->tm[3], dO, C, L, B, 32, (uint32_t)pl->R, "dO row tile")) return e;
    if (int e = encode3(pl->tm[4], dO, C, L, B, 32, FA_CK, "dO chunk")) return e;
    if (int e = encode3(pl->tm[5], dOT, L, C, B, 32, FA_CH, "dOT chunk")) return e;
  }
  return OSM_OK;
}

// split the streamed chunks over a CTA pair when the grid would otherwise leave more than half of the SMs idle (OSM_ATTN_SPLIT=0: never)
static int fa_split(const AttnFlashPlan& pl) {
  static const int on = [] { const char* e = getenv("OSM_ATTN_SPLIT"); return e ? atoi(e) : 1; }();
  static int num_sms = 0;
  if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
  const int n = pl.L / FA_CK;
  const long ctas = (long)(pl.L / pl.R) * pl.heads * pl.B;
  return (on && n >= 4 && n % 2 == 0 && (on == 2 || 2 * ctas <= num_sms)) ? 2 : 1;
}

template <typename... KArgs, typename... Args>
static cudaError_t fa_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster_x; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
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
#define FA_LAUNCH(name, kernel, grid, block, smem, stream, cl, ...)                                   \
  do {                                                                                                \
    cudaError_t _e = fa_launch(kernel, grid, block, smem, stream, cl, __VA_ARGS__);                   \
    if (_e != cudaSuccess) return osm::cuda_fail(_e, name);                                           \
  } while (0)

static FlashParams make_params(const AttnFlashPlan& pl) {
  FlashParams p{};
  p.split = fa_split(pl);
  p.B = pl.B; p.L = pl.L; p.C = pl.C; p.heads = pl.heads; p.R = pl.R;
  p.scale = 1.0f / sqrtf((float)FA_CH);
  p.sl2 = p.scale * 1.4426950408889634f;
  p.out = pl.O; p.O = pl.O; p.lse = pl.lse; p.Dv = pl.Dv; p.dO = pl.dO; p.g_qkv = pl.g_qkv;
  return p;
}

template <typename K>
static int set_smem(K kernel, int bytes, bool* done) {
  if (!*done) {
    OSM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    OSM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    *done = true;
  }
  return OSM_OK;
}

int attn_flash_fwd_launch(const AttnFlashPlan& pl, cudaStream_t s) {
  const FlashParams p = make_params(pl);
  OSM_PREFER_SMEM(tok_to_chan_kernel);
  OSM_LAUNCH_PDL("tok_to_chan_kernel", tok_to_chan_kernel, dim3(pl.L / 32, 3 * pl.C / 32, pl.B), dim3(32, 8), 0, s, pl.qkv, pl.qkvT, pl.L,
                 3 * pl.C);
  constexpr int SMEM = FA_ROWTILE_BYTES + 2 * FA_CHUNK_BYTES + 32768 + 1024;
  static bool done = false;
  if (int e = set_smem(flash_fwd_kernel, SMEM, &done)) return e;
  FA_LAUNCH("flash_fwd_kernel", flash_fwd_kernel, dim3(pl.L / pl.R * p.split, pl.heads, pl.B), dim3(128), SMEM, s, p.split,
            *(const CUtensorMap*)pl.tm[0], *(const CUtensorMap*)pl.tm[1], *(const CUtensorMap*)pl.tm[2], p);
  return OSM_OK;
}

int attn_flash_bwd_launch(const AttnFlashPlan& pl, cudaStream_t s) {
  if (!pl.dO) return fail(OSM_ERR_STATE, "flash attention backward: plan has no gradient buffers");
  const FlashParams p = make_params(pl);
  OSM_PREFER_SMEM(tok_to_chan_kernel);
  OSM_LAUNCH_PDL("tok_to_chan_kernel", tok_to_chan_kernel, dim3(pl.L / 32, pl.C / 32, pl.B), dim3(32, 8), 0, s, pl.dO, pl.dOT, pl.L, pl.C);
  const dim3 grid(pl.L / pl.R * p.split, pl.heads, pl.B);
  {
    constexpr int SMEM = 2 * FA_ROWTILE_BYTES + 3 * FA_CHUNK_BYTES + 32768 + 1024;
    static bool done = false;
    if (int e = set_smem(flash_dq_kernel, SMEM, &done)) return e;
    FA_LAUNCH("flash_dq_kernel", flash_dq_kernel, grid, dim3(256), SMEM, s, p.split, *(const CUtensorMap*)pl.tm[0], *(const CUtensorMap*)pl.tm[1],
              *(const CUtensorMap*)pl.tm[2], *(const CUtensorMap*)pl.tm[3], p);
  }
  {
    constexpr int SMEM = 2 * FA_ROWTILE_BYTES + 4 * FA_CHUNK_BYTES + 2 * 32768 + 1024;
    static bool done = false;
    if (int e = set_smem(flash_dkv_kernel, SMEM, &done)) return e;
    FA_LAUNCH("flash_dkv_kernel", flash_dkv_kernel, grid, dim3(256), SMEM, s, p.split, *(const CUtensorMap*)pl.tm[0], *(const CUtensorMap*)pl.tm[1],
              *(const CUtensorMap*)pl.tm[2], *(const CUtensorMap*)pl.tm[4], *(const CUtensorMap*)pl.tm[5], p);
  }
  return OSM_OK;
}

}  // namespace osm
