// tcgen05 (5th-gen tensor core) TF32 implicit-GEMM convolution for sm_100a.
//
//   out[pixel, co] = sum_{tap, ci} x[pixel + off(tap), ci] * Wp[tap][co][ci]  (+bias, +residual, +=)
//
// NHWC fp32 activations, 3x3 (pad 1) or 1x1, stride 1.  Reference call sites: every nn.Conv2d / Conv1d(k=1)
// of UNetModel (unet.py:264, 290, 301, 365, 373, 561, 694) and, with the flipped/transposed weight pack,
// their input gradients (condition_methods.py:186-191 back-propagates to x_prev only).
//
// Mapping to the hardware
//   * GEMM view: M = 128 output pixels (a tw x th x tn box of one or more images), N = BN output
//     channels, K = taps * Cin in blocks of 32 fp32 (one 128-byte swizzle row).
//   * A tile: ONE 4-D TMA box load {32 ch, tw, th, tn} at coordinates shifted by the tap offset; rows
//     that fall outside the image are zero-filled by the TMA unit, which IS the conv's zero padding -
//     no im2col buffer, no halo handling in the kernel.  B tile: 3-D box {32 ci, BN co, 1 tap} of the
//     packed weights.  Both land in shared memory in the canonical K-major SWIZZLE_128B layout that
//     tcgen05.mma consumes directly.
//   * warp 0 / lane 0: TMA producer over a STAGES-deep mbarrier ring; warp 1 / lane 0: issues
//     tcgen05.mma.kind::tf32 (M128 x N=BN x K8, 4 per stage) accumulating in TMEM, releases stages with
//     tcgen05.commit; all 4 warps: epilogue (tcgen05.ld 32 lanes x 32 columns -> bias/residual -> 128-bit
//     stores).
//   * split-K for small-M layers (8x8 .. 64x64 at small batch, where a 128-pixel tile grid cannot fill 148 SMs and the
//     layer is paced by the length of the per-CTA K loop): a thread-block CLUSTER of `split` (2/4/8/16) CTAs shares one
//     output tile, each CTA accumulates a K-slice in its own TMEM and writes the fp32 partial tile to an L2-resident scratch;
//     after ONE cluster barrier CTA r sums the float4 columns [r BN / (4 split), ...) of all partials in a FIXED rank order -
//     deterministic, no atomics, nothing reaches DRAM - and runs the epilogue.  (The first version exchanged the partials
//     through distributed shared memory; ld.shared::cluster moved ~6 B/clk per SM in that all-to-all gather, 6 us of a 14 us
//     kernel - it is kept behind OSM_CONV_SKRED=0 and for partials that exceed the scratch.)
//   * TF32: the tensor core reads fp32 words from shared memory and ignores the low 13 mantissa bits.
//     Weights are pre-rounded (RN) at pack time and the GroupNorm/SiLU producer rounds the activation it
//     writes, so the truncation is exact on both operands wherever the producer is ours.
#include <stdlib.h>

#include "conv_epilogue.cuh"
#include "tcgen05.cuh"

namespace osm {

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
// one MMA of the non-halo kernels: TF32 (K = 8) or fp16 (K = 16) operands, 32 bytes along the swizzled row either way
__device__ __forceinline__ void mma_tf32_or_f16(int f16, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (f16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    mma_tf32(d_tmem, adesc, bdesc, idesc, accumulate);
  }
}
__device__ __forceinline__ void mma_tf32_or_f16_2sm(int f16, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (f16) mma_f16_2sm(d_tmem, adesc, bdesc, idesc, accumulate);
  else mma_tf32_2sm(d_tmem, adesc, bdesc, idesc, accumulate);
}

// Development-only phase timeline of the split-K kernel (tools/splitk_trace.py builds a separate library with -DOSM_TRACE;
// the product library contains none of it): per CTA and phase, %clock64 and %globaltimer of the last launch.
#ifdef OSM_TRACE
__device__ int g_trace_flags;   // experiments: bit 0 = do not load the A boxes, bit 1 = do not load the weight blocks (results are garbage)
__device__ unsigned long long g_conv_trace[256 * 16 * 2];
__device__ __forceinline__ void trace_put(int ev) {
  const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (cta < 256) {
    unsigned long long c, g;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    g_conv_trace[(cta * 16 + ev) * 2] = c;
    g_conv_trace[(cta * 16 + ev) * 2 + 1] = g;
  }
}
#define OSM_TRACE_PUT(ev) trace_put(ev)
#else
#define OSM_TRACE_PUT(ev)
#endif

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                        // fp32 elements per K block = one 128-byte swizzle row
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;    // 16 KB

struct ConvTcParams {
  int split;  // cluster size along grid.z = number of K slices (1 = no cluster reduction)
  int taps, kblocks_per_tap;
  int tw, th, tn, tiles_w, tiles_h;
  int n_mtiles;  // tiles_w * tiles_h * tiles_b (persistent kernel)
  int B, H, W, Cout_p;
  EpiArgs epi;
  float* sk_ws;             // stream-K: one raw 128 x 256 fp32 partial tile per CTA; cluster split-K: the L2 scratch (null = DSMEM)
  unsigned int* sk_flags;   // stream-K: per-CTA "partial published" counters (self-resetting)
  int f16, bk;              // non-halo kernels: fp16 operands straight from memory (kind::f16), bk = elements per 128-byte K block (32 / 64)
  const float4* xf_coef;    // halo kernel, XFORM: [B][Cin_p] float2 (a, b): the A operand is tf32(SiLU(x a + b)) (gn_coef_fwd_kernel)
  int xf_silu;
};

// Work of one persistent CTA (pair): whole tiles with a grid stride, or - stream-K - a CONTIGUOUS share of the
// tile-major list of K blocks, so that every CTA runs the same number of K blocks and a tile may be split between
// consecutive CTAs.  A segment is (tile, K blocks [kb0, kb1)).
template <bool SK>
struct SegWalk {
  long long g, gend;
  int tile, n_tiles, stride, total_k;
  __device__ SegWalk(int worker, int n_workers, int n_tiles_, int total_k_) : n_tiles(n_tiles_), stride(n_workers), total_k(total_k_) {
    const long long tot = (long long)n_tiles_ * total_k_;
    g = SK ? tot * worker / n_workers : 0;
    gend = SK ? tot * (worker + 1) / n_workers : 0;
    tile = worker;
  }
  __device__ bool valid() const { return SK ? g < gend : tile < n_tiles; }
  __device__ void get(int& t, int& kb0, int& kb1) const {
    if (SK) {
      t = (int)(g / total_k);
      kb0 = (int)(g - (long long)t * total_k);
      const long long rem = gend - g;
      kb1 = rem < (long long)(total_k - kb0) ? kb0 + (int)rem : total_k;
    } else {
      t = tile; kb0 = 0; kb1 = total_k;
    }
  }
  __device__ void next() {
    if (SK) { int t, a, b; get(t, a, b); g += b - a; } else tile += stride;
  }
};

// Split-K reduction of one tile row through the L2 scratch (see conv_tc_kernel): this rank's BN / (4 SPLIT) float4 columns of the
// row, U <= 4 columns per batch with all U x SPLIT partial loads and the epilogue's own operands (bias / residual / previous value) in
// flight before the first add; partials are added in rank order 0, 1, ... (fixed: bit-reproducible, identical to the DSMEM path).
// NSUB threads share a row (the kernel's 8-warp form: thread = (row, sub)): thread `sub` takes the sub-th part of the rank's columns.
template <int BN, int SPLIT, int NSUB>
__device__ __forceinline__ void splitk_reduce_l2(const EpiArgs& epi, const float4* part0, size_t qstride, int rank, int n, int h, int w, int co0,
                                                 int sub) {
  if constexpr (BN / 4 >= SPLIT) {   // (the host takes the DSMEM path otherwise: BN = 32 with a 16-way split)
  constexpr int C4_PER = BN / 4 / SPLIT;
  constexpr int C4_SUB = C4_PER >= NSUB ? C4_PER / NSUB : C4_PER;   // columns per thread (one column: only sub 0 works)
  constexpr int U = C4_SUB < 4 ? C4_SUB : (32 / SPLIT < 4 ? 32 / SPLIT : 4);   // up to 32 partial loads (128 registers) in flight
  static_assert(C4_SUB % U == 0, "split-K: BN / 4 must be a multiple of the split");
  if (C4_PER < NSUB && sub != 0) return;
  const int ibeg = C4_PER >= NSUB ? sub * C4_SUB : 0;
#pragma unroll 1
  for (int i0 = ibeg; i0 < ibeg + C4_SUB; i0 += U) {
    float4 v[U][SPLIT];
    EpiOperands4 ad[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int q = 0; q < SPLIT; ++q) v[u][q] = __ldcg(part0 + (size_t)q * qstride + (size_t)(rank * C4_PER + i0 + u) * TC_BM);
#pragma unroll
    for (int u = 0; u < U; ++u) ad[u] = conv_epilogue_load4(epi, n, h, w, co0 + (rank * C4_PER + i0 + u) * 4);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float4 acc = v[u][0];
#pragma unroll
      for (int q = 1; q < SPLIT; ++q) acc = f4_add(acc, v[u][q]);
      conv_epilogue_apply_store4(epi, n, h, w, co0 + (rank * C4_PER + i0 + u) * 4, acc, ad[u]);
    }
  }
  }
}

// NW = 4 warps, or 8 for the cluster split-K plans that reduce through the L2 scratch: warps 4..7 only join after the main loop -
// two threads per tile row stage the partial (alternate 32-channel chunks) and reduce it (half of the rank's columns each), which
// doubles the loads in flight of that latency-bound tail.
template <int BN, int STAGES, int MINB, int NW = 4>
__global__ void __launch_bounds__(NW * 32, MINB)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
  constexpr int B_BYTES = BN * TC_BK * 4;
  constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 1];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]), accum_bar = smem_u32(&bars[2 * STAGES]);
  if (threadIdx.x == 0) OSM_TRACE_PUT(0);

  // tile coordinates
  int mt = blockIdx.x;
  const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
  const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
  const int w0 = tile_w * p.tw, h0 = tile_h * p.th, n0 = mt * p.tn;
  const int co0 = blockIdx.y * BN;

  const int total_k = p.taps * p.kblocks_per_tap;
  const int rank = p.split > 1 ? (int)cluster_ctarank() : 0;
  const int it0 = rank * total_k / p.split, it1 = (rank + 1) * total_k / p.split;
  // one K block: the tap-shifted A box and the contiguous weight block into ring slot s
  auto produce = [&](int it, int s) {
    const int tap = it / p.kblocks_per_tap, kc = it - tap * p.kblocks_per_tap;
    const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
    const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + TC_A_BYTES;
#ifdef OSM_TRACE
    const int fl = *(volatile int*)&g_trace_flags;
    if (fl) {
      mbar_expect_tx(full0 + 8 * s, ((fl & 1) ? 0 : TC_A_BYTES) + ((fl & 2) ? 0 : B_BYTES) + ((fl & 3) == 3 ? 16 : 0));
      if (!(fl & 1)) tma_load_4d(sa, &tmA, full0 + 8 * s, kc * p.bk, w0 + dx, h0 + dy, n0);
      if (!(fl & 2)) tma_load_3d(sb, &tmB, full0 + 8 * s, 0, co0, it);
      if ((fl & 3) == 3) asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8 * s), "r"(16) : "memory");
      return;
    }
#endif
    mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
    tma_load_4d(sa, &tmA, full0 + 8 * s, kc * p.bk, w0 + dx, h0 + dy, n0);
    tma_load_3d(sb, &tmB, full0 + 8 * s, 0, co0, it);  // packed weights [tap*kpt + kc][co][32]: one contiguous run
  };

  if (threadIdx.x == 32) { prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); }
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // First fill of the ring BEFORE the CTA-wide barrier: the slots are trivially empty, and the loads' latency (~0.5 us) then
    // overlaps the TMEM allocation instead of following it (these launches are latency-bound: a dozen K blocks per CTA).
    pdl_wait();  // the A operand is the previous kernel's output
    for (int it = it0; it < it1 && it < it0 + STAGES; ++it) produce(it, it - it0);
    OSM_TRACE_PUT(2);
  }
  if (warp == 2) {  // one full warp allocates BN TMEM columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();  // everything above is CTA-local (thread 0 has waited before its loads); global memory is first touched below
  if (threadIdx.x == 0) OSM_TRACE_PUT(1);

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: the rest of the K blocks, each waiting for its ring slot =====
      for (int it = it0 + STAGES; it < it1; ++it) {
        const int s = (it - it0) % STAGES;
        const uint32_t ph = (uint32_t)((it - it0) / STAGES) & 1u;
        mbar_wait(empty0 + 8 * s, ph ^ 1u);
        produce(it, s);
      }
      OSM_TRACE_PUT(3);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = p.f16 ? make_idesc_f16(TC_BM, BN) : make_idesc_tf32(TC_BM, BN);
      for (int it = it0; it < it1; ++it) {
        const int s = (it - it0) % STAGES;
        const uint32_t ph = (uint32_t)((it - it0) / STAGES) & 1u;
        mbar_wait(full0 + 8 * s, ph);
        tcgen05_fence_after();
        if (it == it0) OSM_TRACE_PUT(4);
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + TC_A_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {  // UMMA_K = 8 for tf32 = 32 bytes along the swizzled row
          mma_tf32_or_f16(p.f16, tmem_base, make_smem_desc(sa + k * 32), make_smem_desc(sb + k * 32), idesc, (uint32_t)((it != it0) || (k != 0)));
        }
        tcgen05_commit(empty0 + 8 * s);  // stage reusable once these MMAs have read it
      }
      tcgen05_commit(accum_bar);         // accumulator complete
      OSM_TRACE_PUT(5);
    }
    __syncwarp();
  }

  // ===== epilogue: all 4 warps; warp w owns TMEM lanes [32w, 32w+32) = tile rows =====
  mbar_wait(accum_bar, 0);
  pdl_launch_dependents();   // main loop done: the next kernel may launch while the epilogue / cluster reduction runs
  tcgen05_fence_after();
  if (threadIdx.x == 0) OSM_TRACE_PUT(6);
  if (p.split == 1) {
    // Row-per-thread epilogue (thread = tile row = one pixel, 32 consecutive channels per tcgen05.ld).  A shared-memory
    // transpose to make the stores 128-byte contiguous was measured SLOWER (288 vs 398 TFLOP/s on 256->256@256x256): with
    // 4 warps the epilogue is issue/latency-bound, not transaction-bound, and L2 merges the 16-byte pieces of a line.
    const int row = threadIdx.x & 127, sub = threadIdx.x >> 7;
    const int ww = row % p.tw, hh = (row / p.tw) % p.th, nn = row / (p.tw * p.th);
    const int w = w0 + ww, h = h0 + hh, n = n0 + nn;
    const bool row_ok = (w < p.W) && (h < p.H) && (n < p.B);
#pragma unroll 1
    for (int c = sub; c < BN / 32; c += NW / 4) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(c * 32), r);
      if (row_ok) {
        float st[16];
        conv_epilogue_chunk32(p.epi, n, h, w, co0 + c * 32, p.Cout_p, r, st);
      }
    }
  } else if (p.sk_ws) {
    // --- split-K through an L2-resident scratch (the default): every CTA writes its partial tile as [BN / 4 float4 columns][128 rows]
    //     (a warp's store is 512 contiguous bytes), cluster barrier (release / acquire at cluster scope), then thread = row sums,
    //     in rank order, the float4 columns [rank BN / (4 split), ...) of all `split` partials and runs the epilogue.  Only valid
    //     rows are written and read (half of an 8x8 image's 128-row tile is padding).  Why not DSMEM: ld.shared::cluster moves
    //     ~6 B/clk per SM in this all-to-all pattern - the 64 KB a CTA gathers took 6 us of a 14 us kernel (tools/splitk_trace.py,
    //     profiles/r02_splitk_trace.md); L2 moves the same bytes in well under 1 us and needs no second barrier before exit.
    constexpr int TILE_F4 = TC_BM * BN / 4;
    const int row = threadIdx.x & 127, sub = threadIdx.x >> 7;   // (row, sub): NW / 4 threads per tile row
    const int ww = row % p.tw, hh = (row / p.tw) % p.th, nn = row / (p.tw * p.th);
    const int w = w0 + ww, h = h0 + hh, n = n0 + nn;
    const bool row_ok = (w < p.W) && (h < p.H) && (n < p.B);
    float4* ws4 = reinterpret_cast<float4*>(p.sk_ws);
    const size_t lin0 = (size_t)blockIdx.x + (size_t)gridDim.x * blockIdx.y, qstride = (size_t)gridDim.x * gridDim.y * TILE_F4;
    {
      float4* mine = ws4 + lin0 * TILE_F4 + (size_t)rank * qstride + row;
#pragma unroll 1
      for (int c = sub; c < BN / 32; c += NW / 4) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(c * 32), r);
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            __stcg(mine + (size_t)(c * 8 + j) * TC_BM, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                                    __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
        }
      }
    }
    if (threadIdx.x == 0) OSM_TRACE_PUT(7);
    cluster_sync_all();
    if (threadIdx.x == 0) OSM_TRACE_PUT(8);
    if (row_ok) {
      const float4* part0 = ws4 + lin0 * TILE_F4 + row;
      switch (p.split) {
        case 2: splitk_reduce_l2<BN, 2, NW / 4>(p.epi, part0, qstride, rank, n, h, w, co0, sub); break;
        case 4: splitk_reduce_l2<BN, 4, NW / 4>(p.epi, part0, qstride, rank, n, h, w, co0, sub); break;
        case 8: splitk_reduce_l2<BN, 8, NW / 4>(p.epi, part0, qstride, rank, n, h, w, co0, sub); break;
        default: splitk_reduce_l2<BN, 16, NW / 4>(p.epi, part0, qstride, rank, n, h, w, co0, sub); break;
      }
    }
    if (threadIdx.x == 0) OSM_TRACE_PUT(9);
  } else {
    // --- split-K through distributed shared memory (kept for grids whose partials exceed the scratch, and as the reference the
    //     tests compare the L2 path with - same summation order, same bits): stage the partial tile in shared memory (the pipeline
    //     buffers are idle now: every TMA write has been consumed and every MMA has completed), cluster barrier, reduce my row
    //     slice over all ranks through DSMEM ---
    constexpr int LDR = BN + 4;  // padded row stride (floats): conflict-free 128-bit stores from 32 rows at once
    if (threadIdx.x < 128) {     // (the four-warp form; with eight warps the other four only take part in the barriers)
      const int row = threadIdx.x;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), r);
        const uint32_t dst = smem_base + (uint32_t)(row * LDR + c * 32) * 4u;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + j * 4), "r"(r[j]), "r"(r[j + 1]), "r"(r[j + 2]), "r"(r[j + 3])
                       : "memory");
      }
    }
    if (threadIdx.x == 0) OSM_TRACE_PUT(7);
    cluster_sync_all();
    if (threadIdx.x == 0) OSM_TRACE_PUT(8);
    const int rows_per = TC_BM / p.split;
    constexpr int C4N = BN / 4;
    for (int e = threadIdx.x < 128 ? threadIdx.x : rows_per * C4N; e < rows_per * C4N; e += 128) {
      const int row = rank * rows_per + e / C4N, c4 = e % C4N;
      const uint32_t src = smem_base + (uint32_t)(row * LDR + c4 * 4) * 4u;
      float4 acc = ld_dsmem_f4(src, 0);
      for (int q = 1; q < p.split; ++q) acc = f4_add(acc, ld_dsmem_f4(src, (uint32_t)q));
      const int ww = row % p.tw, hh = (row / p.tw) % p.th, nn = row / (p.tw * p.th);
      const int w = w0 + ww, h = h0 + hh, n = n0 + nn, co = co0 + c4 * 4;
      if (w < p.W && h < p.H && n < p.B && co < p.Cout_p) conv_epilogue_store4(p.epi, n, h, w, co, acc);
    }
    if (threadIdx.x == 0) OSM_TRACE_PUT(9);
    cluster_sync_all();  // nobody may exit (and release its shared memory) while a peer is still reading it
    if (threadIdx.x == 0) OSM_TRACE_PUT(10);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
  }
  if (threadIdx.x == 64) OSM_TRACE_PUT(11);
}


// ------------------------------------------------------------------------------------------------
// Persistent variant (split == 1): one CTA per SM walks the tile list; the TMA ring and the MMA issue run continuously
// across tiles (no per-tile pipeline ramp / barrier init / TMEM allocation), accumulators are DOUBLE-BUFFERED in TMEM
// (2 x BN columns) and four dedicated epilogue warps drain tile i while the tensor core already works on tile i+1.
//   warp 0: TMA producer   warp 1: MMA issuer   warps 2..9: epilogue (TMEM lane quadrant = warp % 4, two warps per quadrant)
// Barriers: full/empty per smem stage, tfull/tempty per accumulator buffer (tempty counts one arrival per epilogue warp).
// ------------------------------------------------------------------------------------------------

// EPI_WARPS: 4 (one per TMEM lane quadrant) or 8 (two per quadrant, alternate 32-channel chunks).  A lone warp per
// scheduler issues at ~1/4 IPC (dependent chains, nothing to interleave with): fine for the plain epilogue (~6 us per
// 128x256 tile against a ~30 us main loop), but the fused backward GroupNorm statistics make the epilogue ~4x longer and
// the critical path, so those launches use 8 warps.  (8 warps everywhere cost the plain kernel 3 % at B=8.)
template <int BN, int STAGES, int EPI_WARPS>
__global__ void __launch_bounds__((2 + EPI_WARPS) * 32, 1)
conv_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
  constexpr int B_BYTES = BN * TC_BK * 4;
  constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * STAGES]), tempty0 = smem_u32(&bars[2 * STAGES + 2]);

  if (threadIdx.x == 32) { prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); }
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();

  const int total_k = p.taps * p.kblocks_per_tap;
  const int n_ntiles = p.Cout_p / BN;
  const int n_tiles = p.n_mtiles * n_ntiles;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: one continuous ring over all tiles of this CTA =====
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int mt = tile / n_ntiles;
        const int co0 = (tile - mt * n_ntiles) * BN;
        const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
        const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
        const int w0 = tile_w * p.tw, h0 = tile_h * p.th, n0 = mt * p.tn;
        for (int it = 0; it < total_k; ++it, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          const int tap = it / p.kblocks_per_tap, kc = it - tap * p.kblocks_per_tap;
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
          const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + TC_A_BYTES;
          mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
          tma_load_4d(sa, &tmA, full0 + 8 * s, kc * p.bk, w0 + dx, h0 + dy, n0);
          tma_load_3d(sb, &tmB, full0 + 8 * s, 0, co0, it);
        }
      }
      pdl_launch_dependents();   // every load of this CTA is issued: the next kernel may launch under the last tiles' MMAs / epilogue
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = p.f16 ? make_idesc_f16(TC_BM, BN) : make_idesc_tf32(TC_BM, BN);
      uint32_t g = 0, j = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
        const uint32_t acc = j & 1u, aph = (j >> 1) & 1u;
        mbar_wait(tempty0 + 8 * acc, aph ^ 1u);   // the epilogue has drained this accumulator buffer
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < total_k; ++it, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          mbar_wait(full0 + 8 * s, ph);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + TC_A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k)
            mma_tf32_or_f16(p.f16, d_tmem, make_smem_desc(sa + k * 32), make_smem_desc(sb + k * 32), idesc, (uint32_t)((it != 0) || (k != 0)));
          tcgen05_commit(empty0 + 8 * s);
        }
        tcgen05_commit(tfull0 + 8 * acc);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 2..9: TMEM lane quadrant q = warp % 4, thread = tile row = one pixel; the two warps of a
    //       quadrant take alternate 32-channel chunks =====
    const int q = warp & 3, chunk0 = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int ww = row % p.tw, hh = (row / p.tw) % p.th, nn = row / (p.tw * p.th);
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const uint32_t acc = j & 1u, aph = (j >> 1) & 1u;
      int mt = tile / n_ntiles;
      const int co0 = (tile - mt * n_ntiles) * BN;
      const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
      const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
      const int w = tile_w * p.tw + ww, h = tile_h * p.th + hh, n = mt * p.tn + nn;
      const bool row_ok = (w < p.W) && (h < p.H) && (n < p.B);
      // The epilogue warps get here long before the tile's main loop is done: pull the global operands this row's epilogue
      // will read (residual / accumulate source / GroupNorm input of the fused backward statistics) into L2 now, so the
      // chunk loop below does not expose one DRAM round trip per 32-channel chunk.
      if (row_ok) {
        const size_t pix = ((size_t)n * p.H + h) * p.W + w;
        if (p.epi.res_mode == RES_SAME) {
          const float* q1 = p.epi.res + pix * p.epi.ldr + co0;
          for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(q1 + 32 * c));
        }
        if (p.epi.accumulate) {
          const float* q2 = p.epi.out + pix * p.epi.ldo + co0;
          for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(q2 + 32 * c));
        }
      }
      mbar_wait(tfull0 + 8 * acc, aph);
      tcgen05_fence_after();
      const int mtile = tile / n_ntiles;
#pragma unroll 1
      for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)(c * 32), r);
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        if (row_ok) conv_epilogue_chunk32(p.epi, n, h, w, co0 + c * 32, p.Cout_p, r, st);
        if (p.epi.stat_mode)
          conv_epilogue_stat_flush(st, lane, p.epi.stat_cpg, p.epi.stat_partial + ((size_t)mtile * 4 + q) * 64,
                                   (co0 + c * 32) / p.epi.stat_cpg);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);   // this warp's quadrant of the buffer may be overwritten
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair persistent variant (cta_group::2): the two CTAs of a (2,1,1) cluster - the two SMs of one TPC - run ONE
// 256-row x 256-channel tcgen05.mma per K step.  CTA r of the pair stages its OWN 128-pixel tile of A (the pair's two
// halves are simply two consecutive entries of the 128-pixel tile list) and its own 128-channel HALF of the weight tile:
// 32 KB of shared-memory fill per K block and CTA for the same 128 x 256 x 32 MACs the single-CTA kernel pays 48 KB for,
// i.e. 1/3 less L2 -> SM traffic and 6 stages in flight instead of 4.  The leader (cluster rank 0) issues the MMAs; the
// hardware reads the other halves of A / B from the peer's shared memory and writes each CTA's 128 accumulator rows into
// that CTA's own TMEM, so the epilogue (TMEM double buffer, epilogue warps, fused GroupNorm statistics) is the
// single-CTA one, per CTA.
//   full[s]   lives in the leader: 2 arrivals (both producers) + the bytes of both CTAs' TMA loads (their mbarrier operand
//             has the peer bit cleared, so the complete_tx lands in the leader).
//   empty[s], tfull[a]: one copy per CTA, arrived by the leader's tcgen05.commit with a 2-CTA multicast mask.
//   tempty[a] lives in the leader: one arrival per epilogue warp of BOTH CTAs.
// ------------------------------------------------------------------------------------------------
template <int STAGES, int EPI_WARPS, bool SK>
__global__ void __launch_bounds__((2 + EPI_WARPS) * 32, 1)
conv_tc_persist_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
  constexpr int BN = 256;                       // channels per pair tile
  constexpr int B_BYTES = (BN / 2) * TC_BK * 4; // this CTA's half of the weight tile
  constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * STAGES]), tempty0 = smem_u32(&bars[2 * STAGES + 2]);

  if (threadIdx.x == 32) { prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); }
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full0 + 8 * i, 2);
      mbar_init(empty0 + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 2 * EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // pair-wide TMEM allocation: one warp of EACH CTA issues it
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers are initialised before anything can signal them
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();

  const int total_k = p.taps * p.kblocks_per_tap;
  const int n_ntiles = p.Cout_p / BN;
  const int n_mpairs = (p.n_mtiles + 1) / 2;
  const int n_tiles = n_mpairs * n_ntiles;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs) =====
      uint32_t g = 0;
      for (SegWalk<SK> sw(pair, n_pairs, n_tiles, total_k); sw.valid(); sw.next()) {
        int tile, kb0, kb1;
        sw.get(tile, kb0, kb1);
        const int mp = tile / n_ntiles;
        const int co0 = (tile - mp * n_ntiles) * BN + (int)rank * (BN / 2);
        int mt = 2 * mp + (int)rank;      // an odd tile count leaves the last peer half out of range: TMA zero-fills it
        const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
        const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
        const int w0 = tile_w * p.tw, h0 = tile_h * p.th, n0 = mt * p.tn;
        for (int it = kb0; it < kb1; ++it, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          const int tap = it / p.kblocks_per_tap, kc = it - tap * p.kblocks_per_tap;
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
          const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + TC_A_BYTES;
          if (leader) mbar_expect_tx(full0 + 8 * s, 2 * STAGE_BYTES);
          else mbar_arrive_leader(full0 + 8 * s);
          tma_load_4d_2sm(sa, &tmA, full0 + 8 * s, kc * p.bk, w0 + dx, h0 + dy, n0);
          tma_load_3d_2sm(sb, &tmB, full0 + 8 * s, 0, co0, it);
        }
      }
      pdl_launch_dependents();
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===== MMA issuer (leader only) =====
      const uint32_t idesc = p.f16 ? make_idesc_f16(2 * TC_BM, BN) : make_idesc_tf32(2 * TC_BM, BN);
      uint32_t g = 0, j = 0;
      for (SegWalk<SK> sw(pair, n_pairs, n_tiles, total_k); sw.valid(); sw.next(), ++j) {
        int tile, kb0, kb1;
        sw.get(tile, kb0, kb1);
        const uint32_t acc = j & 1u, aph = (j >> 1) & 1u;
        mbar_wait(tempty0 + 8 * acc, aph ^ 1u);   // both CTAs' epilogues have drained this accumulator buffer
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = kb0; it < kb1; ++it, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          mbar_wait(full0 + 8 * s, ph);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + TC_A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k)
            mma_tf32_or_f16_2sm(p.f16, d_tmem, make_smem_desc(sa + k * 32), make_smem_desc(sb + k * 32), idesc, (uint32_t)((it != kb0) || (k != 0)));
          tcgen05_commit_2sm(empty0 + 8 * s);
        }
        tcgen05_commit_2sm(tfull0 + 8 * acc);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps (both CTAs): as in conv_tc_persist_kernel, on this CTA's own 128 rows =====
    const int q = warp & 3, chunk0 = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int ww = row % p.tw, hh = (row / p.tw) % p.th, nn = row / (p.tw * p.th);
    uint32_t j = 0;
    for (SegWalk<SK> sw(pair, n_pairs, n_tiles, total_k); sw.valid(); sw.next(), ++j) {
      int tile, kb0, kb1;
      sw.get(tile, kb0, kb1);
      const uint32_t acc = j & 1u, aph = (j >> 1) & 1u;
      const int mp = tile / n_ntiles;
      const int co0 = (tile - mp * n_ntiles) * BN;
      const int mtile = 2 * mp + (int)rank;
      int mt = mtile;
      const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
      const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
      const int w = tile_w * p.tw + ww, h = tile_h * p.th + hh, n = mt * p.tn + nn;
      const bool tile_ok = mtile < p.n_mtiles;
      const bool row_ok = tile_ok && (w < p.W) && (h < p.H) && (n < p.B);
      if (row_ok) {
        const size_t pix = ((size_t)n * p.H + h) * p.W + w;
        if (p.epi.res_mode == RES_SAME) {
          const float* q1 = p.epi.res + pix * p.epi.ldr + co0;
          for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(q1 + 32 * c));
        }
        if (p.epi.accumulate) {
          const float* q2 = p.epi.out + pix * p.epi.ldo + co0;
          for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(q2 + 32 * c));
        }
      }
      mbar_wait(tfull0 + 8 * acc, aph);
      tcgen05_fence_after();
      int n_succ = 0;
      if (SK) {
        if (kb0 > 0) {
          // ---- this CTA holds a LATER part of a tile another CTA owns: publish the raw partial sums and move on ----
          // column-major scratch [BN][128 rows]: the 32 lanes of a warp (32 consecutive rows) write one 128-byte line per store
          float* slot = p.sk_ws + (size_t)blockIdx.x * TC_BM * BN + row;
#pragma unroll 1
          for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)(c * 32), r);
#pragma unroll
            for (int i = 0; i < 32; ++i) __stcg(slot + (size_t)(c * 32 + i) * TC_BM, __uint_as_float(r[i]));
          }
          __threadfence();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.sk_flags + blockIdx.x) : "memory");
            mbar_arrive_leader(tempty0 + 8 * acc);
          }
          continue;
        }
        if (kb1 < total_k) {
          // ---- owner of a tile whose tail other CTAs computed: they ran it FIRST (their range starts inside this tile),
          //      so by the time this CTA reaches its last segment the partials are normally there already ----
          const long long tile_end = (long long)(tile + 1) * total_k, tot = (long long)n_tiles * total_k;
          for (int s2 = pair + 1; s2 < n_pairs && tot * s2 / n_pairs < tile_end; ++s2) ++n_succ;
          if (lane == 0) {
            for (int u = 1; u <= n_succ; ++u) {
              const unsigned int* f = p.sk_flags + (blockIdx.x + 2 * u);
              unsigned int v = 0;
              for (uint32_t spin = 0; v < (unsigned)EPI_WARPS; ++spin) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if (v < (unsigned)EPI_WARPS && spin > (1u << 22)) __trap();   // a protocol bug traps instead of hanging the GPU
              }
            }
          }
          __syncwarp();
        }
      }
#pragma unroll 1
      for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)(c * 32), r);
        if (SK) {
          for (int u = 1; u <= n_succ; ++u) {   // fixed order: own K blocks, then the successors' in K order
            const float* src = p.sk_ws + (size_t)(blockIdx.x + 2 * u) * TC_BM * BN + (size_t)(c * 32) * TC_BM + row;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __ldcg(src + (size_t)i * TC_BM);
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + v[i]);
          }
        }
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        if (row_ok) conv_epilogue_chunk32(p.epi, n, h, w, co0 + c * 32, p.Cout_p, r, st);
        if (p.epi.stat_mode && tile_ok)
          conv_epilogue_stat_flush(st, lane, p.epi.stat_cpg, p.epi.stat_partial + ((size_t)mtile * 4 + q) * 64,
                                   (co0 + c * 32) / p.epi.stat_cpg);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty0 + 8 * acc);
      if (SK && n_succ > 0) {
        // every epilogue warp of this CTA has read the partials: clear the successors' counters for the next launch
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
        if (warp == 2 && lane == 0)
          for (int u = 1; u <= n_succ; ++u) p.sk_flags[blockIdx.x + 2 * u] = 0u;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();          // neither CTA may free TMEM / exit while the pair's MMAs or multicast arrivals are in flight
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// HALO variant of the CTA-pair kernel, 3x3 convs: the nine taps of a K block share ONE input tile.
//
// The kernels above load a tap-shifted 128-pixel x 32-channel A tile per (tap, K block): every input element travels
// L2 -> shared memory nine times.  Here a tile is 8 x 16 output pixels of one image and the A operand of a K block
// (32 input channels) is loaded ONCE as the 10 x 18-pixel halo box around it: 180 rows of 128 bytes (23 KB instead of
// 9 x 16 KB), written by TMA in the canonical SWIZZLE_128B pattern (row r of the box at r * 128 bytes, 16-byte chunks
// XOR-ed with r & 7).  The tcgen05 shared-memory descriptor of tap (dy, dx) simply STARTS at halo row (dy+1) * 10 + (dx+1)
// and uses a stride of 1280 bytes (one halo row of 10 pixels) between its 8-row groups: the 8 rows of group g are then the
// 8 pixels of output row g, shifted by the tap.  The swizzle is a function of the absolute shared-memory address, so the
// shifted views read exactly what TMA wrote.  TMA zero-fills the part of the box outside the image = the conv's zero padding.
//
// With the A operand resident once per K block, the GroupNorm + SiLU that precedes every 3x3 conv of a ResBlock
// (unet.py:315-335: h = conv(SiLU(GN(x))), and SiLU(GN(h) (1 + scale) + shift) for the second conv) is applied IN SHARED
// MEMORY (XFORM): four transform warps turn the raw halo rows into cvt.rna.tf32(SiLU(x a[n,c] + b[n,c])) between the TMA
// landing and the first MMA that reads them, once per element (not once per tap); halo rows outside the image stay zero
// (the padding is applied AFTER the activation, as in the reference).  The stand-alone normalise pass (read + write of the
// whole activation) disappears; a and b come from the tiny per-(image, channel) coefficient kernel (gn_coef_kernel).
//
// Rings: A (HALO_NA slots of 23 KB) and B (HALO_NB slots of 16 KB = this CTA's 128-channel half of a weight tile) are
// independent; K order is K-block-major, tap-minor.  Barriers:
//   fullA[s]  (XFORM only) local: TMA of this CTA's halo box        readyA[s] in the leader: XFORM: one arrival per transform
//   warp of BOTH CTAs; else: both producers + the bytes of both boxes (as full[] of the kernel above)
//   fullB[s] in the leader: both producers + both halves' bytes     emptyA / emptyB / tfull: per CTA, multicast commits
//   tempty[a] in the leader: one arrival per epilogue warp of both CTAs.
// ------------------------------------------------------------------------------------------------
constexpr int HALO_TW = 8, HALO_TH = 16;
constexpr int HALO_BW = HALO_TW + 2, HALO_BH = HALO_TH + 2;      // the 10 x 18-pixel box
constexpr int HALO_A_BYTES = HALO_BW * HALO_BH * 128;             // 23040 bytes land per K block
// BN = output channels per pair tile: 256 (128 exists for the measurement recorded in conv_tc_plan and for tests).
constexpr int HALO_PITCH = HALO_BW;                               // shared-memory rows between consecutive halo rows (dense box)
template <int BN> struct HaloCfg {
  static constexpr int B_BYTES = (BN / 2) * TC_BK * 4;            // this CTA's half of a weight tile: 16 / 8 KB
  static constexpr int A_SLOT = 23 * 1024;                        // slots stay 1024-byte aligned
  static constexpr int NA = 3;
  static constexpr int NB = BN == 256 ? 9 : 18;
  static constexpr int SMEM = NA * A_SLOT + NB * B_BYTES + 1024;
};
constexpr int HALO_XF_MAX_C = 1536;                               // input channels whose (a, b) pairs fit behind the rings
constexpr int HALO_XF_COEF_BYTES = HALO_XF_MAX_C * 8;

// One halo row (32 channels) of the operand transform, written as whole-row phases so that the 32 independent element
// chains overlap: the transform warps run one per scheduler, with nothing else to hide an FMA -> ex2 -> rcp -> mul chain
// behind (measured: the element-by-element form made the transform, not the tensor core, the critical path).
// Same arithmetic as gn_apply_kernel's silu_f (ex2.approx / rcp.approx) followed by cvt.rna.tf32.
template <bool SILU>
__device__ __forceinline__ void halo_xf_row(float4 (&v)[8], const float4* __restrict__ cq) {
  float u[32];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 q0 = cq[2 * c], q1 = cq[2 * c + 1];   // (a, b) of channels 4c, 4c+1 | 4c+2, 4c+3
    u[4 * c + 0] = fmaf(v[c].x, q0.x, q0.y); u[4 * c + 1] = fmaf(v[c].y, q0.z, q0.w);
    u[4 * c + 2] = fmaf(v[c].z, q1.x, q1.y); u[4 * c + 3] = fmaf(v[c].w, q1.z, q1.w);
  }
  if (SILU) {
    float e[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) e[i] = __expf(-u[i]);
#pragma unroll
    for (int i = 0; i < 32; ++i) e[i] = __fdividef(1.0f, 1.0f + e[i]);
#pragma unroll
    for (int i = 0; i < 32; ++i) u[i] *= e[i];
  }
#pragma unroll
  for (int c = 0; c < 8; ++c)
    v[c] = make_float4(__uint_as_float(f32_to_tf32_rn(u[4 * c])), __uint_as_float(f32_to_tf32_rn(u[4 * c + 1])),
                       __uint_as_float(f32_to_tf32_rn(u[4 * c + 2])), __uint_as_float(f32_to_tf32_rn(u[4 * c + 3])));
}

template <int EPI_WARPS, bool XFORM, int BN>
__global__ void __launch_bounds__((2 + EPI_WARPS + (XFORM ? 4 : 0)) * 32, 1)
conv_tc_halo_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
  constexpr int PITCH = HALO_PITCH;
  constexpr int TMEM_COLS = 2 * BN;
  constexpr int NA = HaloCfg<BN>::NA, NB = HaloCfg<BN>::NB, A_SLOT = HaloCfg<BN>::A_SLOT, HALO_B_BYTES = HaloCfg<BN>::B_BYTES;
  constexpr int NBAR = 3 * NA + 2 * NB + 4;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smemA = smem_base, smemB = smem_base + NA * A_SLOT;
  __shared__ __align__(8) uint64_t bars[NBAR];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t fullA0 = smem_u32(&bars[0]), readyA0 = smem_u32(&bars[NA]), emptyA0 = smem_u32(&bars[2 * NA]);
  const uint32_t fullB0 = smem_u32(&bars[3 * NA]), emptyB0 = smem_u32(&bars[3 * NA + NB]);
  const uint32_t tfull0 = smem_u32(&bars[3 * NA + 2 * NB]), tempty0 = tfull0 + 16;

  if (threadIdx.x == 32) { prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) {
      mbar_init(fullA0 + 8 * i, 1);
      mbar_init(readyA0 + 8 * i, XFORM ? 8 : 2);
      mbar_init(emptyA0 + 8 * i, 1);
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(fullB0 + 8 * i, 2);
      mbar_init(emptyB0 + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 2 * EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();

  const int kpt = p.kblocks_per_tap;
  const int n_ntiles = p.Cout_p / BN;
  const int n_mpairs = (p.n_mtiles + 1) / 2;
  const int n_tiles = n_mpairs * n_ntiles;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): the A cursor runs up to two K blocks ahead of the B cursor =====
      uint32_t gA = 0, gB = 0, gk = 0;
      int a_tile = pair, a_kc = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        const int co0 = (tile % n_ntiles) * BN + (int)rank * (BN / 2);
        for (int kc = 0; kc < kpt; ++kc, ++gk) {
          while (a_tile < n_tiles && gA < gk + 2) {
            int mt = 2 * (a_tile / n_ntiles) + (int)rank;   // an odd tile count leaves the last peer half out of range: zero-filled
            const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
            const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
            const uint32_t s = gA % NA, ph = (gA / NA) & 1u;
            mbar_wait(emptyA0 + 8 * s, ph ^ 1u);
            const uint32_t dst = smemA + s * A_SLOT;
            const uint32_t bar = (XFORM ? fullA0 : readyA0) + 8 * s;
            const int cx = tile_w * HALO_TW - 1, cy = tile_h * HALO_TH - 1;
            if (XFORM) mbar_expect_tx(bar, HALO_A_BYTES);
            else if (leader) mbar_expect_tx(bar, 2 * HALO_A_BYTES);
            else mbar_arrive_leader(bar);
            if (XFORM) tma_load_4d(dst, &tmA, bar, a_kc * TC_BK, cx, cy, mt);
            else tma_load_4d_2sm(dst, &tmA, bar, a_kc * TC_BK, cx, cy, mt);
            ++gA;
            if (++a_kc == kpt) { a_kc = 0; a_tile += n_pairs; }
          }
          for (int tap = 0; tap < 9; ++tap, ++gB) {
            const uint32_t s = gB % NB, ph = (gB / NB) & 1u;
            mbar_wait(emptyB0 + 8 * s, ph ^ 1u);
            if (leader) mbar_expect_tx(fullB0 + 8 * s, 2 * HALO_B_BYTES);
            else mbar_arrive_leader(fullB0 + 8 * s);
            tma_load_3d_2sm(smemB + s * HALO_B_BYTES, &tmB, fullB0 + 8 * s, 0, co0, tap * kpt + kc);
          }
        }
      }
      pdl_launch_dependents();
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===== MMA issuer (leader only): 9 taps x 4 K steps per halo tile =====
      constexpr uint32_t idesc = make_idesc_tf32(2 * TC_BM, BN);
      uint32_t gA = 0, gB = 0, j = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs, ++j) {
        const uint32_t acc = j & 1u, aph = (j >> 1) & 1u;
        mbar_wait(tempty0 + 8 * acc, aph ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = 0; kc < kpt; ++kc, ++gA) {
          const uint32_t sa = gA % NA, pha = (gA / NA) & 1u;
          mbar_wait(readyA0 + 8 * sa, pha);
          tcgen05_fence_after();
          const uint32_t a_slot = smemA + sa * A_SLOT;
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap, ++gB) {
            const uint32_t sb = gB % NB, phb = (gB / NB) & 1u;
            mbar_wait(fullB0 + 8 * sb, phb);
            tcgen05_fence_after();
            const uint32_t a_tap = a_slot + (uint32_t)((tap / 3) * PITCH + (tap % 3)) * 128u;
            const uint32_t b_slot = smemB + sb * HALO_B_BYTES;
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k)
              mma_tf32_2sm(d_tmem, make_smem_desc_sbo(a_tap + k * 32, PITCH * 128), make_smem_desc(b_slot + k * 32), idesc,
                           (uint32_t)((kc != 0) || (tap != 0) || (k != 0)));
            tcgen05_commit_2sm(emptyB0 + 8 * sb);
          }
          tcgen05_commit_2sm(emptyA0 + 8 * sa);
        }
        tcgen05_commit_2sm(tfull0 + 8 * acc);
      }
    }
    __syncwarp();
  } else if (warp < 2 + EPI_WARPS) {
    // ===== epilogue warps (both CTAs), as in conv_tc_persist_2sm_kernel =====
    const int q = warp & 3, chunk0 = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int ww = row % HALO_TW, hh = row / HALO_TW;
    uint32_t j = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs, ++j) {
      const uint32_t acc = j & 1u, aph = (j >> 1) & 1u;
      const int mp = tile / n_ntiles;
      const int co0 = (tile - mp * n_ntiles) * BN;
      const int mtile = 2 * mp + (int)rank;
      int mt = mtile;
      const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
      const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
      const int w = tile_w * HALO_TW + ww, h = tile_h * HALO_TH + hh, n = mt;
      const bool tile_ok = mtile < p.n_mtiles;
      const bool row_ok = tile_ok && (w < p.W) && (h < p.H) && (n < p.B);
      if (row_ok) {
        const size_t pix = ((size_t)n * p.H + h) * p.W + w;
        if (p.epi.res_mode == RES_SAME) {
          const float* q1 = p.epi.res + pix * p.epi.ldr + co0;
          for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(q1 + 32 * c));
        }
        if (p.epi.accumulate) {
          const float* q2 = p.epi.out + pix * p.epi.ldo + co0;
          for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(q2 + 32 * c));
        }
      }
      mbar_wait(tfull0 + 8 * acc, aph);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)(c * 32), r);
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        if (row_ok) conv_epilogue_chunk32(p.epi, n, h, w, co0 + c * 32, p.Cout_p, r, st);
        if (p.epi.stat_mode && tile_ok)
          conv_epilogue_stat_flush(st, lane, p.epi.stat_cpg, p.epi.stat_partial + ((size_t)mtile * 4 + q) * 64,
                                   (co0 + c * 32) / p.epi.stat_cpg);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty0 + 8 * acc);
    }
  } else if (XFORM) {
    // ===== transform warps (both CTAs): raw halo rows -> tf32(SiLU(x a + b)) in place, once per K block =====
    const int t = (warp - (2 + EPI_WARPS)) * 32 + lane;   // 0..127: halo pixels t and t + 128 of the 180
    const int C = kpt * TC_BK;
    const int hy0 = t / HALO_BW, hx0 = t - hy0 * HALO_BW;
    const int hy1 = (t + 128) / HALO_BW, hx1 = (t + 128) - hy1 * HALO_BW;
    const int r0 = hy0 * PITCH + hx0, r1 = hy1 * PITCH + hx1;   // shared-memory rows of the two pixels
    float4* coef_s = reinterpret_cast<float4*>(smem_raw + (smem_base - smem_u32(smem_raw)) + NA * A_SLOT + NB * HALO_B_BYTES);
    int coef_img = -1;
    uint32_t gA = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs) {
      const int mtile = 2 * (tile / n_ntiles) + (int)rank;
      int mt = mtile;
      const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
      const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
      const bool tile_ok = mtile < p.n_mtiles;
      const int w0 = tile_w * HALO_TW - 1, h0 = tile_h * HALO_TH - 1;
      // pixels of this thread that lie inside the image (the others are the zero padding and stay untouched)
      const bool in0 = tile_ok && (unsigned)(w0 + hx0) < (unsigned)p.W && (unsigned)(h0 + hy0) < (unsigned)p.H;
      const bool in1 = tile_ok && (t + 128) < HALO_BW * HALO_BH && (unsigned)(w0 + hx1) < (unsigned)p.W && (unsigned)(h0 + hy1) < (unsigned)p.H;
      // (a, b) of every input channel of this tile's image, staged once per image in shared memory: every thread needs all
      // 32 pairs of a K block, so they are read back as warp-wide broadcasts (one wavefront each)
      const int img = tile_ok ? mt : 0;
      if (img != coef_img) {
        asm volatile("bar.sync 2, 128;" ::: "memory");          // nobody still reads the previous image's pairs
        const float4* src = p.xf_coef + ((size_t)img * C >> 1);
        for (int i = t; i < C / 2; i += 128) coef_s[i] = __ldg(src + i);
        asm volatile("bar.sync 2, 128;" ::: "memory");
        coef_img = img;
      }
      for (int kc = 0; kc < kpt; ++kc, ++gA) {
        const uint32_t s = gA % NA, ph = (gA / NA) & 1u;
        mbar_wait(fullA0 + 8 * s, ph);
        uint8_t* slot = smem_raw + (smem_base - smem_u32(smem_raw)) + s * A_SLOT;
        uint8_t* row0 = slot + r0 * 128;
        uint8_t* row1 = slot + r1 * 128;
        const float4* cq = coef_s + kc * (TC_BK / 2);
#pragma unroll 1
        for (int u = 0; u < 2; ++u) {
          if (u == 0 ? in0 : in1) {
            uint8_t* rowp = u == 0 ? row0 : row1;
            const int sw = (u == 0 ? r0 : r1) & 7;
            float4 v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(rowp + ((c ^ sw) << 4));
            if (p.xf_silu) halo_xf_row<true>(v, cq); else halo_xf_row<false>(v, cq);
#pragma unroll
            for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(rowp + ((c ^ sw) << 4)) = v[c];
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader_release(readyA0 + 8 * s);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// fp16-operand halo kernel (kind::f16, fp32 accumulation in TMEM): the halo kernel above with K blocks of 64 channels.
//
// TF32 and fp16 carry the same 11-bit significand; fp16 only has the narrower exponent.  tcgen05 runs kind::f16 at twice the
// kind::tf32 rate and a 128-byte shared-memory row holds 64 fp16 channels instead of 32 fp32 ones, so the same bytes through
// the shared-memory port feed twice the MACs - the port, not the tensor pipe, is what bounds the TF32 kernels above.
// Activations and gradients stay fp32 in HBM: the raw fp32 halo boxes (two 32-channel boxes per K block) land in a RAW ring
// by TMA and six transform warps (one halo pixel per thread) write the fp16 operand rows (same canonical SWIZZLE_128B K-major
// layout, row = halo pixel) into the OPERAND ring, applying on the way
//     mode 0: nothing (conversion only; gradients, and activations some other kernel normalised)
//     mode 1: x a[n,c] + b[n,c]                 (GroupNorm without activation)
//     mode 2: SiLU(x a[n,c] + b[n,c])           (GroupNorm + scale-shift + SiLU, unet.py:315-335)
// followed by cvt.rn.satfinite.f16x2 (round to nearest; a value beyond +-65504 saturates instead of becoming inf).  Range:
// forward operands are normalised activations (O(1..10)); backward operands are gradients the engine pre-scales per image
// by a power of two (exact in fp32, undone on the way out; unet_engine.cu vjp) so that they sit mid-range.  Weights are
// packed as fp16 [tap][Cin/64][Cout][64] (pack_conv_weight_f16_kernel).
//
// 16 warps, 128 registers each: warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue (two per TMEM lane quadrant, alternate
// 32-channel chunks), warps 10..15 transform.  With the main loop of a tile twice as short as in the TF32 kernels the epilogue
// (one thread per pixel row: 16-byte stores 1 KB apart, 32 L1 wavefronts per instruction) is as long as the main loop, hence
// eight warps and the register-lean epilogue below (a spilled scalar costs an L2 round trip: the L1 is configured away).
// Rings per CTA: RAW NR x (2 x 23 KB), OPERAND NA x 23 KB, B NB x 16 KB (this CTA's 128-channel half of a 64-channel weight
// tile).  Barriers:
//   fullR[s]  local: TMA bytes of the two raw boxes          emptyR[s] local: one arrival (transform thread 0, after the warps' barrier)
//   readyA[s] in the leader: one arrival per CTA (after fence.proxy.async by every writer + the transform warps' named barrier)
//   emptyA[s], emptyB[s], tfull[a]: per CTA, multicast tcgen05.commit of the leader
//   fullB[s] in the leader: both producers + both halves' bytes    tempty[a] in the leader: epilogue warps of both CTAs
// ------------------------------------------------------------------------------------------------
constexpr int H16_BK = 64;                          // fp16 channels per K block
constexpr int H16_RAW_BOX = 23 * 1024;              // one fp32 halo box of 32 channels (23040 bytes) in a 1024-byte-aligned slot
constexpr int H16_NR = 2, H16_NA = 2, H16_NB = 5;
constexpr int H16_A_SLOT = 23 * 1024;
constexpr int H16_B_BYTES = 128 * H16_BK * 2;       // 16 KB
constexpr int H16_EPI_WARPS = 8, H16_XF_WARPS = 6;
constexpr int H16_THREADS = (2 + H16_EPI_WARPS + H16_XF_WARPS) * 32;
constexpr int H16_COEF_BYTES = 2 * H16_BK * 8;      // (a, b) of the 64 channels of a K block, double-buffered
constexpr int H16_ECOEF_BYTES = 256 * 16;           // (a, b, e, 0) of the tile's 256 output channels for the backward-statistics epilogue
constexpr int H16_SMEM = H16_NR * 2 * H16_RAW_BOX + H16_NA * H16_A_SLOT + H16_NB * H16_B_BYTES + H16_COEF_BYTES + H16_ECOEF_BYTES + 1024;

// 16 channels of one halo pixel: raw fp32 (4 x float4) -> 8 packed fp16x2 words
template <int MODE>
__device__ __forceinline__ void h16_xf16(const float4 (&r)[4], const float4* __restrict__ cq, uint4& o0, uint4& o1) {
  float u[16];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (MODE == 0) {
      u[4 * c + 0] = r[c].x; u[4 * c + 1] = r[c].y; u[4 * c + 2] = r[c].z; u[4 * c + 3] = r[c].w;
    } else {
      const float4 q0 = cq[2 * c], q1 = cq[2 * c + 1];   // (a, b) of channels 4c, 4c+1 | 4c+2, 4c+3 (shared memory, warp-wide broadcast)
      u[4 * c + 0] = fmaf(r[c].x, q0.x, q0.y); u[4 * c + 1] = fmaf(r[c].y, q0.z, q0.w);
      u[4 * c + 2] = fmaf(r[c].z, q1.x, q1.y); u[4 * c + 3] = fmaf(r[c].w, q1.z, q1.w);
    }
  }
  if (MODE == 2) {
    float e[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) e[i] = __expf(-u[i]);
#pragma unroll
    for (int i = 0; i < 16; ++i) e[i] = __fdividef(1.0f, 1.0f + e[i]);
#pragma unroll
    for (int i = 0; i < 16; ++i) u[i] *= e[i];
  }
  o0 = make_uint4(pack_f16x2_sat(u[0], u[1]), pack_f16x2_sat(u[2], u[3]), pack_f16x2_sat(u[4], u[5]), pack_f16x2_sat(u[6], u[7]));
  o1 = make_uint4(pack_f16x2_sat(u[8], u[9]), pack_f16x2_sat(u[10], u[11]), pack_f16x2_sat(u[12], u[13]), pack_f16x2_sat(u[14], u[15]));
}

// BN = output channels per CTA-pair tile: 256, or 64 for the 4- / 8-channel input-gradient / output convs (channels padded to 64)
template <int BN>
__global__ void __launch_bounds__(H16_THREADS, 1)
conv_tc_halo16_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
  constexpr uint32_t B_BYTES = (BN / 2) * H16_BK * 2;   // this CTA's half of a weight tile (the ring slots keep the 16 KB stride)
  constexpr int PITCH = HALO_PITCH;
  constexpr int TMEM_COLS = 2 * BN;
  constexpr int NR = H16_NR, NA = H16_NA, NB = H16_NB;
  constexpr int EPI_WARPS = H16_EPI_WARPS;
  constexpr int NBAR = 2 * NR + 2 * NA + 2 * NB + 4;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smemR = smem_base, smemA = smemR + NR * 2 * H16_RAW_BOX, smemB = smemA + NA * H16_A_SLOT;
  __shared__ __align__(8) uint64_t bars[NBAR];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t fullR0 = smem_u32(&bars[0]), emptyR0 = fullR0 + 8 * NR;
  const uint32_t readyA0 = emptyR0 + 8 * NR, emptyA0 = readyA0 + 8 * NA;
  const uint32_t fullB0 = emptyA0 + 8 * NA, emptyB0 = fullB0 + 8 * NB;
  const uint32_t tfull0 = emptyB0 + 8 * NB, tempty0 = tfull0 + 16;

  if (threadIdx.x == 32) { prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NR; ++i) {
      mbar_init(fullR0 + 8 * i, 1);
      mbar_init(emptyR0 + 8 * i, 1);
    }
    for (int i = 0; i < NA; ++i) {
      mbar_init(readyA0 + 8 * i, 2);
      mbar_init(emptyA0 + 8 * i, 1);
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(fullB0 + 8 * i, 2);
      mbar_init(emptyB0 + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 2 * EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();

  const int kpt = p.kblocks_per_tap;           // K blocks of 64 channels
  const int n_ntiles = p.Cout_p / BN;
  const int n_mpairs = (p.n_mtiles + 1) / 2;
  const int n_tiles = n_mpairs * n_ntiles;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): raw boxes into this CTA's RAW ring (up to NR K blocks ahead), weight halves into the B ring =====
      uint32_t gA = 0, gB = 0, gk = 0;
      int a_tile = pair, a_kc = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        const int co0 = (tile % n_ntiles) * BN + (int)rank * (BN / 2);
        for (int kc = 0; kc < kpt; ++kc, ++gk) {
          while (a_tile < n_tiles && gA < gk + NR) {
            int mt = 2 * (a_tile / n_ntiles) + (int)rank;   // an odd tile count leaves the last peer half out of range: zero-filled
            const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
            const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
            const uint32_t s = gA % NR, ph = (gA / NR) & 1u;
            mbar_wait(emptyR0 + 8 * s, ph ^ 1u);
            const uint32_t dst = smemR + s * 2 * H16_RAW_BOX;
            const uint32_t bar = fullR0 + 8 * s;
            const int cx = tile_w * HALO_TW - 1, cy = tile_h * HALO_TH - 1;
            mbar_expect_tx(bar, 2 * HALO_A_BYTES);
            tma_load_4d(dst, &tmA, bar, a_kc * H16_BK, cx, cy, mt);
            tma_load_4d(dst + H16_RAW_BOX, &tmA, bar, a_kc * H16_BK + 32, cx, cy, mt);
            ++gA;
            if (++a_kc == kpt) { a_kc = 0; a_tile += n_pairs; }
          }
          for (int tap = 0; tap < 9; ++tap, ++gB) {
            const uint32_t s = gB % NB, ph = (gB / NB) & 1u;
            mbar_wait(emptyB0 + 8 * s, ph ^ 1u);
            if (leader) mbar_expect_tx(fullB0 + 8 * s, 2 * B_BYTES);
            else mbar_arrive_leader(fullB0 + 8 * s);
            tma_load_3d_2sm(smemB + s * H16_B_BYTES, &tmB, fullB0 + 8 * s, 0, co0, tap * kpt + kc);
          }
        }
      }
      pdl_launch_dependents();
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===== MMA issuer (leader only): 9 taps x 4 K steps of 16 per 64-channel K block =====
      constexpr uint32_t idesc = make_idesc_f16(2 * TC_BM, BN);
      uint32_t gA = 0, gB = 0, j = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs, ++j) {
        const uint32_t acc = j & 1u, aph = (j >> 1) & 1u;
        mbar_wait(tempty0 + 8 * acc, aph ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = 0; kc < kpt; ++kc, ++gA) {
          const uint32_t sa = gA % NA, pha = (gA / NA) & 1u;
          mbar_wait(readyA0 + 8 * sa, pha);
          tcgen05_fence_after();
          const uint32_t a_slot = smemA + sa * H16_A_SLOT;
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap, ++gB) {
            const uint32_t sb = gB % NB, phb = (gB / NB) & 1u;
            mbar_wait(fullB0 + 8 * sb, phb);
            tcgen05_fence_after();
            const uint32_t a_tap = a_slot + (uint32_t)((tap / 3) * PITCH + (tap % 3)) * 128u;
            const uint32_t b_slot = smemB + sb * H16_B_BYTES;
#pragma unroll
            for (int k = 0; k < H16_BK / 16; ++k)
              mma_f16_2sm(d_tmem, make_smem_desc_sbo(a_tap + k * 32, PITCH * 128), make_smem_desc(b_slot + k * 32), idesc,
                          (uint32_t)((kc != 0) || (tap != 0) || (k != 0)));
            tcgen05_commit_2sm(emptyB0 + 8 * sb);
          }
          tcgen05_commit_2sm(emptyA0 + 8 * sa);
        }
        tcgen05_commit_2sm(tfull0 + 8 * acc);
      }
    }
    __syncwarp();
  } else if (warp < 2 + EPI_WARPS) {
    // ===== epilogue warps 2..9 (both CTAs): TMEM lane quadrant q = warp % 4, thread = tile row = one pixel; the two warps of a
    //       quadrant take alternate 32-channel chunks =====
    const int q = warp & 3, chunk0 = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int ww = row % HALO_TW, hh = row / HALO_TW;
    float4* const ecoef_s = reinterpret_cast<float4*>(smem_raw + (smem_base - smem_u32(smem_raw)) + (size_t)NR * 2 * H16_RAW_BOX +
                                                      (size_t)NA * H16_A_SLOT + (size_t)NB * H16_B_BYTES + H16_COEF_BYTES);
    uint32_t j = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs, ++j) {
      const uint32_t acc = j & 1u, aph = (j >> 1) & 1u;
      const int mp = tile / n_ntiles;
      const int co0 = (tile - mp * n_ntiles) * BN;
      const int mtile = 2 * mp + (int)rank;
      int mt = mtile;
      const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
      const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
      const int w = tile_w * HALO_TW + ww, h = tile_h * HALO_TH + hh, n = mt;
      const bool tile_ok = mtile < p.n_mtiles;
      const bool row_ok = tile_ok && (w < p.W) && (h < p.H) && (n < p.B);
      if (p.epi.stat_mode == 2) {
        // backward statistics: the tile's 256 (a, b, e) coefficient triples go to shared memory once per tile (while the tile's
        // main loop still runs) instead of 32 global loads per chunk and thread
        asm volatile("bar.sync 3, %0;" ::"n"(EPI_WARPS * 32) : "memory");   // every reader of the previous tile's triples is done
        const int e_ = (warp - 2) * 32 + lane;
        if (e_ < BN) ecoef_s[e_] = tile_ok ? __ldg(p.epi.stat_coef + (size_t)n * p.Cout_p + co0 + e_) : make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("bar.sync 3, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      }
      // Pull the global operands of this row's epilogue (residual / `+=` source / GroupNorm input of the backward statistics) into
      // L2 ONE TILE AHEAD: when the epilogue is as long as the main loop the accumulator of the next tile is ready the moment this
      // one is drained, and a prefetch issued at that point hides nothing (the first tile prefetches for itself as well).
      for (int ahead = (j == 0 ? 0 : 1); ahead < 2; ++ahead) {
        const int t2 = tile + ahead * n_pairs;
        if (t2 >= n_tiles) break;
        const int mp2 = t2 / n_ntiles, co2 = (t2 - mp2 * n_ntiles) * BN;
        int m2 = 2 * mp2 + (int)rank;
        if (m2 >= p.n_mtiles) continue;
        const int tw2 = m2 % p.tiles_w; m2 /= p.tiles_w;
        const int th2 = m2 % p.tiles_h; m2 /= p.tiles_h;
        const int w2 = tw2 * HALO_TW + ww, h2 = th2 * HALO_TH + hh;
        if (w2 >= p.W || h2 >= p.H || m2 >= p.B) continue;
        const size_t pix = ((size_t)m2 * p.H + h2) * p.W + w2;
        if (p.epi.res_mode == RES_SAME) {
          const float* q1 = p.epi.res + pix * p.epi.ldr + co2;
          for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(q1 + 32 * c));
        }
        if (p.epi.accumulate) {
          const float* q2 = p.epi.out + pix * p.epi.ldo + co2;
          for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(q2 + 32 * c));
        }
        if (p.epi.stat_mode == 2) {
          const float* q3 = p.epi.stat_x + pix * p.epi.stat_ldx + co2;
          for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(q3 + 32 * c));
        }
      }
      mbar_wait(tfull0 + 8 * acc, aph);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)(c * 32), r);
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        if (row_ok) conv_epilogue_chunk32_lean(p.epi, n, h, w, co0 + c * 32, p.Cout_p, r, st, ecoef_s + c * 32);
        if (p.epi.stat_mode && tile_ok)
          conv_epilogue_stat_flush(st, lane, p.epi.stat_cpg, p.epi.stat_partial + ((size_t)mtile * 4 + q) * 64,
                                   (co0 + c * 32) / p.epi.stat_cpg);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty0 + 8 * acc);
    }
  } else {
    // ===== transform warps 10..15 (both CTAs): thread tt < 180 owns halo pixel tt = shared-memory row tt of the raw boxes and of the
    //       operand slot and converts its 64 channels of every K block, 16 at a time.  The (a, b) pairs of a K block are staged in
    //       shared memory one K block ahead by the first transform warp (one float4 = two channels per lane). =====
    const int xw = warp - (2 + EPI_WARPS);
    const int tt = xw * 32 + lane;
    const bool have = tt < HALO_BW * HALO_BH;
    const int C = kpt * H16_BK;
    const int hy = tt / HALO_BW, hx = tt - hy * HALO_BW;
    const int sw = tt & 7;
    uint8_t* const smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    float4* const coef_s = reinterpret_cast<float4*>(smem_al + (size_t)NR * 2 * H16_RAW_BOX + (size_t)NA * H16_A_SLOT + (size_t)NB * H16_B_BYTES);
    const int mode = p.xf_coef ? (p.xf_silu ? 2 : 1) : 0;
    const int per_img = p.tiles_w * p.tiles_h;
    auto tile_img = [&](int tile_) {
      const int mtile_ = 2 * (tile_ / n_ntiles) + (int)rank;
      return mtile_ < p.n_mtiles ? mtile_ / per_img : 0;
    };
    if (mode && xw == 0 && pair < n_tiles) coef_s[lane] = __ldg(p.xf_coef + (((size_t)tile_img(pair) * C) >> 1) + lane);
    asm volatile("bar.sync 2, %0;" ::"n"(H16_XF_WARPS * 32) : "memory");
    uint32_t gA = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs) {
      const int mtile = 2 * (tile / n_ntiles) + (int)rank;
      int mt = mtile;
      const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
      const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
      const bool tile_ok = mtile < p.n_mtiles;
      const int w0 = tile_w * HALO_TW - 1, h0 = tile_h * HALO_TH - 1;
      // is this thread's pixel inside the image?  the others are the conv's zero padding (applied AFTER the activation)
      const bool in = tile_ok && have && (unsigned)(w0 + hx) < (unsigned)p.W && (unsigned)(h0 + hy) < (unsigned)p.H;
      for (int kc = 0; kc < kpt; ++kc, ++gA) {
        const uint32_t sr = gA % NR, phr = (gA / NR) & 1u;
        const uint32_t sa = gA % NA, pha = (gA / NA) & 1u;
        // (a, b) of the NEXT K block: in flight during this block's work
        float4 cnext = make_float4(0.f, 0.f, 0.f, 0.f);
        bool have_next = false;
        if (mode && xw == 0) {
          int nt = tile, nk = kc + 1;
          if (nk == kpt) { nk = 0; nt += n_pairs; }
          if (nt < n_tiles) { have_next = true; cnext = __ldg(p.xf_coef + (((size_t)tile_img(nt) * C + (size_t)nk * H16_BK) >> 1) + lane); }
        }
        const float4* cq = coef_s + (gA & 1u) * (H16_BK / 2);
        const uint8_t* rrow = smem_al + (size_t)sr * 2 * H16_RAW_BOX + (size_t)tt * 128;
        uint8_t* orow = smem_al + (size_t)NR * 2 * H16_RAW_BOX + (size_t)sa * H16_A_SLOT + (size_t)tt * 128;
        mbar_wait(fullR0 + 8 * sr, phr);           // the raw boxes have landed
        mbar_wait(emptyA0 + 8 * sa, pha ^ 1u);     // the MMAs that read the operand slot's previous contents are done
        if (have) {
#pragma unroll 2
          for (int hq = 0; hq < 4; ++hq) {         // 16 channels at a time: raw box hq / 2, quarter hq % 2
            uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
            if (in) {
              const uint8_t* rb = rrow + (hq >> 1) * H16_RAW_BOX;
              float4 r[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) r[c] = *reinterpret_cast<const float4*>(rb + ((((hq & 1) * 4 + c) ^ sw) << 4));
              if (mode == 2) h16_xf16<2>(r, cq + hq * 8, o0, o1);
              else if (mode == 1) h16_xf16<1>(r, cq + hq * 8, o0, o1);
              else h16_xf16<0>(r, cq, o0, o1);
            }
            *reinterpret_cast<uint4*>(orow + (((2 * hq) ^ sw) << 4)) = o0;
            *reinterpret_cast<uint4*>(orow + (((2 * hq + 1) ^ sw) << 4)) = o1;
          }
        }
        if (have_next) coef_s[((gA + 1u) & 1u) * (H16_BK / 2) + lane] = cnext;
        fence_proxy_async_smem();
        asm volatile("bar.sync 2, %0;" ::"n"(H16_XF_WARPS * 32) : "memory");
        if (tt == 0) {
          mbar_arrive(emptyR0 + 8 * sr);                // raw slot may be refilled
          mbar_arrive_leader(readyA0 + 8 * sa);         // this CTA's operand rows are in place
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent variant with 256-pixel x 256-channel tiles (split == 1, Cout_p % 256 == 0, enough tiles to fill the SMs).
// The main loop of every variant is bound by the bytes a CTA pulls into shared memory (~92 GB/s per SM measured), so
// this one shares each 32 KB weight tile between TWO 128-row MMAs: 64 KB per K block for 256x256x32 MACs instead of
// 2 x 48 KB (-33 % bytes per FLOP).  The two fp32 accumulators fill all 512 TMEM columns, so the epilogue cannot be
// double-buffered against the next tile's MMAs; eight epilogue warps (two per TMEM lane quadrant, one per
// accumulator) keep that exposed drain short, and the TMA ring keeps prefetching the next tile meanwhile.
//   warp 0: TMA producer   warp 1: MMA issuer   warps 2..9: epilogue
// ------------------------------------------------------------------------------------------------
template <int STAGES>
__global__ void __launch_bounds__(320, 1)
conv_tc_persist_m256_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
  constexpr int BN = 256;
  constexpr int A_BYTES = 2 * TC_A_BYTES;
  constexpr int B_BYTES = BN * TC_BK * 4;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
  const uint32_t tfull = smem_u32(&bars[2 * STAGES]), tempty = smem_u32(&bars[2 * STAGES + 1]);
  pdl_launch_dependents();

  if (threadIdx.x == 32) { prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); }
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();

  const int total_k = p.taps * p.kblocks_per_tap;
  const int n_ntiles = p.Cout_p / BN;
  const int n_tiles = p.n_mtiles * n_ntiles;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int mt = tile / n_ntiles;
        const int co0 = (tile - mt * n_ntiles) * BN;
        const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
        const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
        const int w0 = tile_w * p.tw, h0 = tile_h * p.th, n0 = mt * p.tn;
        for (int it = 0; it < total_k; ++it, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          const int tap = it / p.kblocks_per_tap, kc = it - tap * p.kblocks_per_tap;
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
          const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
          mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
          tma_load_4d(sa, &tmA, full0 + 8 * s, kc * TC_BK, w0 + dx, h0 + dy, n0);   // 256 pixel rows x 32 channels
          tma_load_3d(sb, &tmB, full0 + 8 * s, 0, co0, it);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(TC_BM, BN);
      uint32_t g = 0, j = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
        mbar_wait(tempty, (j & 1u) ^ 1u);   // all eight epilogue warps have drained the previous tile
        tcgen05_fence_after();
        for (int it = 0; it < total_k; ++it, ++g) {
          const uint32_t s = g % STAGES, ph = (g / STAGES) & 1u;
          mbar_wait(full0 + 8 * s, ph);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t bdesc = make_smem_desc(sb + k * 32);
            const uint32_t acc = (uint32_t)((it != 0) || (k != 0));
            mma_tf32(tmem_base, make_smem_desc(sa + k * 32), bdesc, idesc, acc);                      // pixel rows 0..127
            mma_tf32(tmem_base + BN, make_smem_desc(sa + TC_A_BYTES + k * 32), bdesc, idesc, acc);    // pixel rows 128..255
          }
          tcgen05_commit(empty0 + 8 * s);
        }
        tcgen05_commit(tfull);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 2..9: TMEM lane quadrant q = warp % 4; warps 2..5 drain accumulator 0, warps 6..9 accumulator 1 =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = half * TC_BM + q * 32 + lane;
    const int ww = row % p.tw, hh = (row / p.tw) % p.th, nn = row / (p.tw * p.th);
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      int mt = tile / n_ntiles;
      const int co0 = (tile - mt * n_ntiles) * BN;
      const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
      const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
      const int w = tile_w * p.tw + ww, h = tile_h * p.th + hh, n = mt * p.tn + nn;
      const bool row_ok = (w < p.W) && (h < p.H) && (n < p.B);
      mbar_wait(tfull, j & 1u);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * BN) + (uint32_t)(c * 32), r);
        if (row_ok) {
          float st[16];
          conv_epilogue_chunk32(p.epi, n, h, w, co0 + c * 32, p.Cout_p, r, st);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host: tensor maps + launch
// ------------------------------------------------------------------------------------------------
PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

int conv_check(const ConvArgs& a);

static void pick_tile(int rows, int H, int W, int* tw, int* th, int* tn) {
  int w = 1;
  while (w * 2 <= W && w * 2 <= 16) w *= 2;
  int h = 1;
  while (h * 2 <= H && w * h * 2 <= rows) h *= 2;
  *tw = w; *th = h; *tn = rows / (w * h);
}

// Clusters of `split` CTAs of the one-tile-per-CTA kernel that can be co-resident on this device (cached per variant).
template <int BN, int STAGES>
static int query_clusters(int split) {
  const size_t smem = (size_t)STAGES * (TC_A_BYTES + BN * TC_BK * 4) + 1024;
  if (cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  if (split > 8 && cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1, 1, (unsigned)split);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)split;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<BN, STAGES, 1>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
static int max_active_clusters(int BN, int split) {
  static int cache[3][5] = {{-1, -1, -1, -1, -1}, {-1, -1, -1, -1, -1}, {-1, -1, -1, -1, -1}};
  const int bi = BN == 256 ? 0 : (BN == 128 ? 1 : 2);
  const int si = split == 1 ? 0 : (split == 2 ? 1 : (split == 4 ? 2 : (split == 8 ? 3 : 4)));
  if (cache[bi][si] < 0) {
    int n = BN == 256 ? query_clusters<256, 4>(split) : (BN == 128 ? query_clusters<128, 6>(split) : query_clusters<64, 8>(split));
    // B200 values, used only if the query is unavailable; 16-CTA clusters are opt-in (non-portable): no query, no use
    if (n <= 0) n = split == 16 ? 0 : (split == 8 ? 15 : (split == 4 ? 32 : 72));
    cache[bi][si] = n;
  }
  return cache[bi][si];
}

// stream-K scratch: one raw partial tile per CTA + one counter per CTA, allocated once (at plan time, outside any capture)
static float* g_sk_ws = nullptr;
static unsigned int* g_sk_flags = nullptr;
// Cluster split-K scratch (conv_tc_kernel): one fp32 partial tile per CTA of the launch.  64 MB hold every split plan of the shipped
// configs (batch 1: <= 128 CTAs x 128 KB; batch 32: <= 512 x 64 KB); a launch whose partials would not fit reduces through DSMEM.
// The engine passes 64 MB of its own workspace (ConvArgs::splitk_ws); this process-wide one serves callers without a workspace
// (the osm_dbg_* entry points): launches that use it must be stream-ordered, like the stream-K scratch.
constexpr size_t SPLITK_WS_BYTES = (size_t)64 << 20;
static float* g_splitk_ws = nullptr;
static int splitk_scratch_ensure() {
  if (g_splitk_ws) return OSM_OK;
  OSM_CUDA_CHECK(cudaMalloc(&g_splitk_ws, SPLITK_WS_BYTES));
  return OSM_OK;
}
static int sk_scratch_ensure() {
  if (g_sk_ws) return OSM_OK;
  OSM_CUDA_CHECK(cudaMalloc(&g_sk_ws, (size_t)160 * TC_BM * 256 * sizeof(float)));
  OSM_CUDA_CHECK(cudaMalloc(&g_sk_flags, 160 * sizeof(unsigned int)));
  OSM_CUDA_CHECK(cudaMemset(g_sk_flags, 0, 160 * sizeof(unsigned int)));
  return OSM_OK;
}

int conv_tc_plan(const ConvArgs& a, ConvTcPlan* plan) {
  if (int e = conv_check(a)) return e;
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(OSM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (needs a CUDA 12 driver)");
  plan->a = a;
  pick_tile(TC_BM, a.H, a.W, &plan->tw, &plan->th, &plan->tn);
  plan->tiles_w = (a.W + plan->tw - 1) / plan->tw;
  plan->tiles_h = (a.H + plan->th - 1) / plan->th;
  plan->tiles_b = (a.B + plan->tn - 1) / plan->tn;
  plan->halo = 0;
  plan->f16 = 0;
  if (a.f16 && a.Cin_p % H16_BK != 0) return fail(OSM_ERR_INVALID, "conv_tc: fp16 operands need Cin % 64 == 0");
  const int bk = a.f16 ? H16_BK : TC_BK;      // elements per 128-byte K block
  if (a.f16 && !a.halo) plan->f16 = 2;        // non-halo kernels reading an fp16 activation tensor directly
  if (a.halo) {
    if (!(a.f16 ? conv_tc_halo16_ok(a.B, a.H, a.W, a.Cin_p, a.Cout_p, a.taps) : conv_tc_halo_ok(a.B, a.H, a.W, a.Cin_p, a.Cout_p, a.taps)))
      return fail(OSM_ERR_INVALID, "conv_tc: the halo kernel takes 3x3 convs with Cout % 256 == 0 (fp16: or Cout == 64), H % 16 == 0, W % 8 == 0");
    plan->halo = 1;
    plan->f16 = a.f16 ? 1 : 0;
    plan->tw = HALO_TW; plan->th = HALO_TH; plan->tn = 1;
    plan->tiles_w = a.W / HALO_TW; plan->tiles_h = a.H / HALO_TH; plan->tiles_b = a.B;
  } else if (a.xf_coef) {
    return fail(OSM_ERR_INVALID, "conv_tc: an operand transform (xf_coef) needs the halo kernel");
  }
  if (a.xf_coef && !a.f16 && a.Cin_p > HALO_XF_MAX_C)
    return fail(OSM_ERR_INVALID, "conv_tc: the operand transform takes at most 1536 input channels");
  const long mtiles = (long)plan->tiles_w * plan->tiles_h * plan->tiles_b;
  // Tile policy: pick (BN, split) by a small cost model fitted to measurements on B200 (profiles/r02_splitk_trace.md; the
  // round-1 fit is in profiles/r01_conv_policy.md).  Times in us, inside a CUDA graph, weights streamed from HBM:
  //   * a K block (four 128 x BN tcgen05.mma) takes ~0.4 us on the one CTA that runs it whatever BN is - these small-M layers
  //     are bound by the LENGTH of the per-CTA K loop, not by bytes: 0.385 + 0.0002 BN (persistent) / 0.42 + 0.00025 BN (split);
  //   * split == 1 runs the persistent kernel: ~4 us of launch + prologue + epilogue, ceil(tiles / SMs) waves of the full K loop;
  //   * split > 1 runs one tile per cluster with K / split blocks per CTA and a fixed cost of 4.9 us + 0.018 us x BN x (fraction
  //     of valid tile rows) for launch, staging the partials in L2, the cluster barrier and the reduction (10 us when the
  //     reduction went through distributed shared memory), and AT MOST cudaOccupancyMaxActiveClusters clusters at once (15
  //     clusters of 8 on a B200: a 16th cluster waits for a whole extra wave).
  const int total_k = a.taps * (a.Cin_p / bk);
  int BN = 256, split = 1, m256 = 0;
  double best = 1e30;   // modelled time (us) of the chosen single-CTA / cluster split-K variant
  if (!plan->halo) {
    int num_sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); }
    for (int bn = 256; bn >= 64; bn /= 2) {
      if (a.Cout_p % bn) continue;
      const long tiles = mtiles * (a.Cout_p / bn);
      static const int max_split = [] { const char* e = getenv("OSM_CONV_MAX_SPLIT"); return e ? atoi(e) : 16; }();
      // fraction of the 128 tile rows that are pixels (an 8x8 image at batch 1 fills half a tile): the reduction moves only those
      const double valid = (double)a.B * a.H * a.W / ((double)mtiles * TC_BM);
      for (int sp = 1; sp <= max_split; sp *= 2) {
        if (sp > 1 && (total_k / sp < 2 || bn / 4 < sp)) break;
        double t;
        if (sp == 1) {
          t = 4.1 + (double)((tiles + num_sms - 1) / num_sms) * total_k * (0.385 + 0.0002 * bn);
        } else {
          const int maxc = max_active_clusters(bn, sp);
          if (maxc <= 0) continue;
          // 16-CTA clusters (non-portable size, one per GPC): a 16-way reduction, ~0.5 us more than the 8-way one
          t = (double)((tiles + maxc - 1) / maxc) *
              (4.9 + 0.018 * bn * valid + (sp == 16 ? 0.5 : 0.0) + (double)((total_k + sp - 1) / sp) * (0.42 + 0.00025 * bn));
        }
        if (t < best * 0.97) { best = t; BN = bn; split = sp; }  // prefer the wider / less split variant on near-ties
      }
    }
    // 256-pixel x 256-channel persistent tiles: 64 KB per K block for twice the MACs, ~5 us of exposed epilogue per tile
    static const int allow_m256 = [] { const char* e = getenv("OSM_CONV_M256"); return e ? atoi(e) : 0; }();
    if (allow_m256 && !a.f16 && a.Cout_p % 256 == 0) {
      int tw2, th2, tn2;
      pick_tile(2 * TC_BM, a.H, a.W, &tw2, &th2, &tn2);
      if (tn2 <= 256) {
        const long mt2 = (long)((a.W + tw2 - 1) / tw2) * ((a.H + th2 - 1) / th2) * ((a.B + tn2 - 1) / tn2);
        const long tiles = mt2 * (a.Cout_p / 256);
        const double t = 6.0 + (double)((tiles + num_sms - 1) / num_sms) * (total_k * (65536.0 / 92e3) + 5.0);
        if (t < best * 0.97 || allow_m256 == 2) { best = t; BN = 256; split = 1; m256 = 1; }
      }
    }
    if (best > 1e29) {  // Cout_p not a multiple of 64 (e.g. the 32-channel output conv)
      BN = 256;
      while (BN > 32 && a.Cout_p % BN != 0) BN /= 2;
    }
  }
  if (const char* e = getenv("OSM_CONV_NO_SPLIT")) { if (e[0] == '1') split = 1; }
  if (const char* e = plan->halo ? nullptr : getenv("OSM_CONV_FORCE")) {  // development: "BN,split" for every layer (tools/time_conv.py sweeps)
    int fbn = 0, fsp = 0;
    if (sscanf(e, "%d,%d", &fbn, &fsp) == 2 && fbn >= 32 && a.Cout_p % fbn == 0 && fsp >= 1 && total_k / fsp >= 1) { BN = fbn; split = fsp; m256 = 0; }
  }
  if (m256) {
    pick_tile(2 * TC_BM, a.H, a.W, &plan->tw, &plan->th, &plan->tn);
    plan->tiles_w = (a.W + plan->tw - 1) / plan->tw;
    plan->tiles_h = (a.H + plan->th - 1) / plan->th;
    plan->tiles_b = (a.B + plan->tn - 1) / plan->tn;
  }
  int stages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  // Experiment (OSM_CONV_2CTA=1, off by default): two co-resident CTAs per SM with 128-wide tiles and 3 stages each
  // (2 x 97 KB shared memory, 2 x 128 TMEM columns) so one CTA's epilogue overlaps the other's main loop.  Measured on
  // B200: no gain (256->256@256x256: 512 vs 530 TFLOP/s at B=8) - the main loop is bound by TMA latency x bytes in flight,
  // and the narrower tile costs 33 % more L2 traffic.
  static const int two_cta = [] { const char* e = getenv("OSM_CONV_2CTA"); return e ? atoi(e) : 0; }();
  if (two_cta && !plan->halo && !m256 && split == 1 && BN >= 128 && a.Cout_p % 128 == 0 && mtiles * (a.Cout_p / 128) >= 2 * 148) { BN = 128; stages = 3; }
  // CTA-pair kernel (conv_tc_persist_2sm_kernel): the persistent 256-wide plan with at least one full wave of pair tiles
  // OSM_CONV_2SM: 0 = off, 1 (default) = when the pair tiles fill at least one wave of 74 pairs, 2 = wherever it applies
  // (tests).  Read per plan (plans are built at bind time, not per launch) so a test can switch it.
  // OSM_CONV_SK (default 0): stream-K on top - the 74 pairs split the tile-major list of K blocks evenly instead of walking
  // whole tiles, which removes the partial last wave (512 tiles = 3.46 waves at 256x256, B = 1) and lets layers with fewer
  // pair tiles than pairs use every SM; a tile shared by consecutive pairs is completed by its first pair from the raw
  // partials the others publish (fixed order, self-resetting counters, no atomics on data).  Correct (parity tests run it
  // with OSM_CONV_SK=2) but measured SLOWER on B200 despite the shorter K loops - 256x256 256->256: 140 vs 129 us,
  // 128x128 256->256: 76 vs 47 us, a constant ~28 us per launch: the pairs no longer sweep the weight K blocks in lockstep
  // and every CTA ends on a fix-up epilogue nothing overlaps.  Kept as a documented switch.
  const int two_sm = [] { const char* e = getenv("OSM_CONV_2SM"); return e ? atoi(e) : 1; }();
  const int sk_on = [] { const char* e = getenv("OSM_CONV_SK"); return e ? atoi(e) : 0; }();
  plan->two_sm = 0;
  if (plan->halo) {
    // 256-channel pair tiles.  128-channel tiles (a.halo == 128, tests only) quantise better into waves of 74 pairs (512 tiles
    // = 6.92 waves at 256x256, B = 1, against 3.46 -> 4 waves) but were measured 55 % SLOWER on B200: a tcgen05.mma of
    // 256 x 128 x 8 TF32 takes as long as one of 256 x 256 x 8 (~200 clk per instruction either way), so halving N halves the
    // work per instruction slot.
    plan->two_sm = 1;
    BN = a.f16 ? (a.Cout_p == 64 ? 64 : 256) : (a.halo == 128 ? 128 : 256);
    stages = a.f16 ? H16_NB : (BN == 256 ? HaloCfg<256>::NB : HaloCfg<128>::NB);
  } else if (two_sm && !m256 && a.Cout_p % 256 == 0) {
    const long ptiles = ((mtiles + 1) / 2) * (a.Cout_p / 256);
    const long share = ptiles * total_k / 74;                         // K blocks per pair under stream-K
    const bool sk_ok = sk_on && ptiles >= (sk_on == 2 ? 1 : 16) && share >= 8 && share * 4 >= total_k && (ptiles % 74) != 0;
    const bool plain_ok = split == 1 && BN == 256 && (mtiles / 2) * (a.Cout_p / 256) >= (two_sm == 2 ? 1 : 74);
    if (sk_ok && (plain_ok || sk_on == 2 || ptiles < 74)) {
      // below one wave of pair tiles the alternative is the cluster split-K kernel: take stream-K when its K-block count wins
      const double t_sk = 8.0 + (double)share * 0.455 + 3.0;
      if (plain_ok || sk_on == 2 || t_sk < best) { plan->two_sm = 2; BN = 256; split = 1; stages = 6; }
    } else if (plain_ok) {
      plan->two_sm = 1;
      static const int st2 = [] { const char* e = getenv("OSM_CONV_2SM_STAGES"); return e ? atoi(e) : 6; }();
      stages = st2 == 7 ? 7 : 6;   // 7 x 32 KB = 224 KB of the 227 KB: one more K block in flight per CTA
    }
    if (plan->two_sm == 2) { if (int e = sk_scratch_ensure()) return e; }
  }
  if (split > 1 && !a.splitk_ws) { if (int e = splitk_scratch_ensure()) return e; }   // at plan time: never inside a stream capture
  plan->BN = BN;
  plan->split = split;
  plan->stages = m256 ? 3 : stages;
  plan->m256 = m256;
  if (getenv("OSM_CONV_VERBOSE"))
    fprintf(stderr, "conv_tc_plan: B=%d %dx%d Cin=%d Cout=%d taps=%d -> mtiles=%ld BN=%d split=%d m256=%d\n", a.B, a.H, a.W, a.Cin_p,
            a.Cout_p, a.taps, mtiles, BN, split, m256);
  plan->smem_bytes = (size_t)plan->stages * ((m256 ? 2 : 1) * TC_A_BYTES + (plan->two_sm ? BN / 2 : BN) * TC_BK * 4) + 1024;
  if (plan->halo) plan->smem_bytes = plan->f16 ? H16_SMEM : (BN == 256 ? HaloCfg<256>::SMEM : HaloCfg<128>::SMEM);

  // A: NHWC view as a 4-D tensor {C, W, H, B}
  {
    const bool a16 = plan->f16 == 2;            // the activation tensor itself is fp16 (the halo fp16 kernel reads fp32 boxes)
    const cuuint64_t es = a16 ? 2 : 4;
    cuuint64_t dims[4] = {(cuuint64_t)a.Cin_p, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    cuuint64_t strides[3] = {(cuuint64_t)a.ldx * es, (cuuint64_t)a.W * a.ldx * es, (cuuint64_t)a.H * a.W * a.ldx * es};
    cuuint32_t box[4] = {(cuuint32_t)(a16 ? H16_BK : TC_BK), (cuuint32_t)plan->tw, (cuuint32_t)plan->th, (cuuint32_t)plan->tn};
    if (plan->halo) { box[1] = HALO_BW; box[2] = HALO_BH; box[3] = 1; }   // the halo box
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc((CUtensorMap*)plan->tmA, a16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)a.x, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OSM_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: code " + std::to_string((int)r));
  }
  // B: packed weights as {32 ci, Cout_p, taps * Cin_p/32 K blocks}; fp16 pack: {64 ci, Cout_p, taps * Cin_p/64 K blocks}
  if (plan->f16) {
    cuuint64_t dims[3] = {(cuuint64_t)H16_BK, (cuuint64_t)a.Cout_p, (cuuint64_t)a.taps * (a.Cin_p / H16_BK)};
    cuuint64_t strides[2] = {(cuuint64_t)H16_BK * 2, (cuuint64_t)a.Cout_p * H16_BK * 2};
    cuuint32_t box[3] = {H16_BK, (cuuint32_t)(plan->two_sm ? BN / 2 : BN), 1};   // the pair kernels load half a weight tile per CTA
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc((CUtensorMap*)plan->tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)a.w, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OSM_ERR_CUDA, "cuTensorMapEncodeTiled(B, fp16) failed: code " + std::to_string((int)r));
  } else {
    cuuint64_t dims[3] = {(cuuint64_t)TC_BK, (cuuint64_t)a.Cout_p, (cuuint64_t)a.taps * (a.Cin_p / TC_BK)};
    cuuint64_t strides[2] = {(cuuint64_t)TC_BK * 4, (cuuint64_t)a.Cout_p * TC_BK * 4};
    cuuint32_t box[3] = {TC_BK, (cuuint32_t)(plan->two_sm ? BN / 2 : BN), 1};   // the pair kernel loads half a weight tile per CTA
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc((CUtensorMap*)plan->tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)a.w, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OSM_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: code " + std::to_string((int)r));
  }
  return OSM_OK;
}

template <int BN, int STAGES, int MINB, int NW>
static int launch_t_nw(const ConvTcPlan& pl, const ConvTcParams& p, dim3 grid, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    OSM_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, MINB, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)pl.smem_bytes));
    OSM_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, MINB, NW>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        (int)cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  if (p.split > 8) {
    static bool np_set = false;
    if (!np_set) {
      OSM_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES, MINB, NW>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      np_set = true;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)p.split;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  OSM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, STAGES, MINB, NW>, *(const CUtensorMap*)pl.tmA, *(const CUtensorMap*)pl.tmB, p));
  return OSM_OK;
}
// eight warps for the split-K plans that reduce through the L2 scratch (OSM_CONV_SK_WARPS=4: the four-warp form everywhere)
template <int BN, int STAGES, int MINB>
static int launch_t(const ConvTcPlan& pl, const ConvTcParams& p, dim3 grid, cudaStream_t s) {
  const int wide = [] { const char* e = getenv("OSM_CONV_SK_WARPS"); return e ? atoi(e) : 8; }();
  if (MINB == 1 && p.split > 1 && p.sk_ws && wide == 8) return launch_t_nw<BN, STAGES, 1, 8>(pl, p, grid, s);
  return launch_t_nw<BN, STAGES, MINB, 4>(pl, p, grid, s);
}

template <int BN, int STAGES, int EPI_WARPS>
static int launch_persist_e(const ConvTcPlan& pl, const ConvTcParams& p, cudaStream_t s) {
  static bool attr_set = false;
  static int num_sms = 148;
  if (!attr_set) {
    OSM_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_persist_kernel<BN, STAGES, EPI_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)pl.smem_bytes));
    OSM_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_persist_kernel<BN, STAGES, EPI_WARPS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        (int)cudaSharedmemCarveoutMaxShared));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr_set = true;
  }
  const long n_tiles = (long)p.n_mtiles * (p.Cout_p / BN);
  const unsigned grid = (unsigned)(n_tiles < num_sms ? n_tiles : num_sms);
  OSM_LAUNCH_PDL("conv_tc_persist_kernel", (conv_tc_persist_kernel<BN, STAGES, EPI_WARPS>), dim3(grid), dim3((2 + EPI_WARPS) * 32),
                 pl.smem_bytes, s, *(const CUtensorMap*)pl.tmA, *(const CUtensorMap*)pl.tmB, p);
  return OSM_OK;
}
template <int BN, int STAGES>
static int launch_persist(const ConvTcPlan& pl, const ConvTcParams& p, cudaStream_t s) {
  static const int force = [] { const char* e = getenv("OSM_CONV_EPI_WARPS"); return e ? atoi(e) : 0; }();
  const bool wide = force ? force == 8 : p.epi.stat_mode == 2;
  return wide ? launch_persist_e<BN, STAGES, 8>(pl, p, s) : launch_persist_e<BN, STAGES, 4>(pl, p, s);
}

static int launch_persist_m256(const ConvTcPlan& pl, const ConvTcParams& p, cudaStream_t s) {
  static bool attr_set = false;
  static int num_sms = 148;
  if (!attr_set) {
    OSM_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_persist_m256_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    OSM_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_persist_m256_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        (int)cudaSharedmemCarveoutMaxShared));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr_set = true;
  }
  const long n_tiles = (long)p.n_mtiles * (p.Cout_p / 256);
  const unsigned grid = (unsigned)(n_tiles < num_sms ? n_tiles : num_sms);
  OSM_LAUNCH_PDL("conv_tc_persist_m256_kernel", conv_tc_persist_m256_kernel<3>, dim3(grid), dim3(320), pl.smem_bytes, s,
                 *(const CUtensorMap*)pl.tmA, *(const CUtensorMap*)pl.tmB, p);
  return OSM_OK;
}

template <int EPI_WARPS, bool SK, int STAGES2 = 6>
static int launch_persist_2sm(const ConvTcPlan& pl, ConvTcParams p, cudaStream_t s) {
  static bool attr_set = false;
  static int max_pairs = 74;
  auto kern = conv_tc_persist_2sm_kernel<STAGES2, EPI_WARPS, SK>;
  if (!attr_set) {
    OSM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    OSM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    int dev = 0, num_sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    max_pairs = num_sms / 2;
    // stream-K CTAs wait for each other: never launch more pairs than can be co-resident
    cudaLaunchConfig_t q{};
    q.gridDim = dim3(2 * max_pairs); q.blockDim = dim3((2 + EPI_WARPS) * 32); q.dynamicSmemBytes = pl.smem_bytes;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    q.attrs = qa; q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &q) == cudaSuccess && n > 0 && n < max_pairs) max_pairs = n;
    cudaGetLastError();
    attr_set = true;
  }
  const long n_tiles = (long)((p.n_mtiles + 1) / 2) * (p.Cout_p / 256);
  unsigned pairs = (unsigned)(n_tiles < max_pairs ? n_tiles : max_pairs);
  if (SK) {
    pairs = (unsigned)max_pairs;   // every pair gets an equal share of the K blocks
    p.sk_ws = g_sk_ws; p.sk_flags = g_sk_flags;
    if (!g_sk_ws || 2 * pairs > 160) return fail(OSM_ERR_STATE, "conv_tc: stream-K scratch missing");
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3((2 + EPI_WARPS) * 32);
  cfg.dynamicSmemBytes = pl.smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  OSM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, *(const CUtensorMap*)pl.tmA, *(const CUtensorMap*)pl.tmB, p));
  return OSM_OK;
}

// the fp16-operand halo kernel: 256-channel pair tiles, or ONE 64-channel tile for the narrow (4- / 8-channel, padded to 64) convs
bool conv_tc_halo16_ok(int B, int H, int W, int Cin_p, int Cout_p, int taps) {
  return taps == 9 && B >= 1 && (Cout_p % 256 == 0 || Cout_p == 64) && Cin_p % H16_BK == 0 && H % HALO_TH == 0 && W % HALO_TW == 0;
}
bool conv_tc_halo_ok(int B, int H, int W, int Cin_p, int Cout_p, int taps) {
  return taps == 9 && B >= 1 && Cout_p % 256 == 0 && Cin_p % TC_BK == 0 && H % HALO_TH == 0 && W % HALO_TW == 0;
}

template <int EPI_WARPS, bool XFORM, int BN>
static int launch_halo(const ConvTcPlan& pl, const ConvTcParams& p, cudaStream_t s) {
  static bool attr_set = false;
  static int max_pairs = 74;
  auto kern = conv_tc_halo_2sm_kernel<EPI_WARPS, XFORM, BN>;
  constexpr int SMEM = HaloCfg<BN>::SMEM + (XFORM ? HALO_XF_COEF_BYTES : 0);
  if (!attr_set) {
    OSM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    OSM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    int dev = 0, num_sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    max_pairs = num_sms / 2;
    attr_set = true;
  }
  const long n_tiles = (long)((p.n_mtiles + 1) / 2) * (p.Cout_p / BN);
  const unsigned pairs = (unsigned)(n_tiles < max_pairs ? n_tiles : max_pairs);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3((2 + EPI_WARPS + (XFORM ? 4 : 0)) * 32);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  OSM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, *(const CUtensorMap*)pl.tmA, *(const CUtensorMap*)pl.tmB, p));
  return OSM_OK;
}
template <int BN>
static int launch_halo16(const ConvTcPlan& pl, const ConvTcParams& p, cudaStream_t s) {
  static bool attr_set = false;
  static int max_pairs = 74;
  auto kern = conv_tc_halo16_2sm_kernel<BN>;
  if (!attr_set) {
    OSM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, H16_SMEM));
    OSM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    int dev = 0, num_sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    max_pairs = num_sms / 2;
    attr_set = true;
  }
  const long n_tiles = (long)((p.n_mtiles + 1) / 2) * (p.Cout_p / BN);
  const unsigned pairs = (unsigned)(n_tiles < max_pairs ? n_tiles : max_pairs);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(H16_THREADS);
  cfg.dynamicSmemBytes = H16_SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  OSM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, *(const CUtensorMap*)pl.tmA, *(const CUtensorMap*)pl.tmB, p));
  return OSM_OK;
}
template <int EPI_WARPS, bool XFORM>
static int launch_halo_p(const ConvTcPlan& pl, const ConvTcParams& p, cudaStream_t s) {
  return pl.BN == 256 ? launch_halo<EPI_WARPS, XFORM, 256>(pl, p, s) : launch_halo<EPI_WARPS, XFORM, 128>(pl, p, s);
}

// Fused GroupNorm statistics need the persistent 128-row kernel with every tile inside one image.
bool conv_tc_stats_capable(const ConvTcPlan& pl) { return pl.split == 1 && !pl.m256 && pl.stages != 3 && pl.tn == 1 && pl.BN >= 32; }
int conv_tc_stat_slots(const ConvTcPlan& pl) { return pl.tiles_w * pl.tiles_h * 4; }  // partial slots per image

int conv_tc_launch(const ConvTcPlan& pl, cudaStream_t s) {
  const ConvArgs& a = pl.a;
  ConvTcParams p;
  p.split = pl.split;
  p.taps = a.taps; p.kblocks_per_tap = a.Cin_p / (pl.f16 ? H16_BK : TC_BK);
  p.f16 = pl.f16 == 2; p.bk = pl.f16 ? H16_BK : TC_BK;
  p.tw = pl.tw; p.th = pl.th; p.tn = pl.tn; p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h;
  p.n_mtiles = pl.tiles_w * pl.tiles_h * pl.tiles_b;
  p.B = a.B; p.H = a.H; p.W = a.W; p.Cout_p = a.Cout_p;
  p.epi = EpiArgs{a.bias, a.res, a.ldr, a.res_mode, a.out, a.ldo, a.accumulate, a.H, a.W,
                  a.stat_mode, a.stat_cpg, a.stat_partial, a.stat_x, a.stat_ldx, (const float4*)a.stat_coef, a.stat_silu};
  p.sk_ws = nullptr; p.sk_flags = nullptr;
  p.xf_coef = (const float4*)a.xf_coef; p.xf_silu = a.xf_silu;
  if (pl.split > 1 && !pl.halo) {
    // OSM_CONV_SKRED=0: reduce through distributed shared memory (the earlier path; tests compare the two bit for bit)
    const int l2 = [] { const char* e = getenv("OSM_CONV_SKRED"); return e ? atoi(e) : 1; }();
    const size_t need = (size_t)pl.tiles_w * pl.tiles_h * pl.tiles_b * (a.Cout_p / pl.BN) * pl.split * TC_BM * pl.BN * sizeof(float);
    float* ws = a.splitk_ws ? a.splitk_ws : g_splitk_ws;
    const size_t ws_bytes = a.splitk_ws ? a.splitk_ws_bytes : SPLITK_WS_BYTES;
    if (l2 && ws && need <= ws_bytes && pl.BN / 4 >= pl.split) p.sk_ws = ws;
  }
  if (a.stat_mode && !conv_tc_stats_capable(pl)) return fail(OSM_ERR_STATE, "conv_tc: fused statistics requested on a non-capable plan");
  if (pl.halo) {
    const bool wide = p.epi.stat_mode == 2;
    if (pl.f16 == 1) return pl.BN == 64 ? launch_halo16<64>(pl, p, s) : launch_halo16<256>(pl, p, s);
    if (a.xf_coef) return wide ? launch_halo_p<8, true>(pl, p, s) : launch_halo_p<4, true>(pl, p, s);
    return wide ? launch_halo_p<8, false>(pl, p, s) : launch_halo_p<4, false>(pl, p, s);
  }
  dim3 grid((unsigned)((long)pl.tiles_w * pl.tiles_h * pl.tiles_b), (unsigned)(a.Cout_p / pl.BN), (unsigned)pl.split);
  if (pl.m256) return launch_persist_m256(pl, p, s);
  if (pl.two_sm) {
    static const int force = [] { const char* e = getenv("OSM_CONV_EPI_WARPS"); return e ? atoi(e) : 0; }();
    const bool wide = force ? force == 8 : p.epi.stat_mode == 2;
    if (pl.two_sm == 2) return wide ? launch_persist_2sm<8, true>(pl, p, s) : launch_persist_2sm<4, true>(pl, p, s);
    if (pl.stages == 7) return wide ? launch_persist_2sm<8, false, 7>(pl, p, s) : launch_persist_2sm<4, false, 7>(pl, p, s);
    return wide ? launch_persist_2sm<8, false>(pl, p, s) : launch_persist_2sm<4, false>(pl, p, s);
  }
  static const int persist = [] { const char* e = getenv("OSM_CONV_PERSIST"); return e ? atoi(e) : 1; }();
  if (persist && pl.split == 1 && pl.stages != 3) {
    switch (pl.BN) {
      case 256: return launch_persist<256, 4>(pl, p, s);
      case 128: return launch_persist<128, 6>(pl, p, s);
      case 64: return launch_persist<64, 8>(pl, p, s);
      case 32: return launch_persist<32, 8>(pl, p, s);
    }
  }
  switch (pl.BN) {
    case 256: return launch_t<256, 4, 1>(pl, p, grid, s);
    case 128: return pl.stages == 3 ? launch_t<128, 3, 2>(pl, p, grid, s) : launch_t<128, 6, 1>(pl, p, grid, s);
    case 64: return launch_t<64, 8, 1>(pl, p, grid, s);
    case 32: return launch_t<32, 8, 1>(pl, p, grid, s);
  }
  return fail(OSM_ERR_INVALID, "conv_tc: unsupported BN");
}

}  // namespace osm

#ifdef OSM_TRACE
extern "C" int osm_dbg_trace_read(unsigned long long* host_out, int n_words) {
  if (n_words > 256 * 16 * 2) n_words = 256 * 16 * 2;
  return (int)cudaMemcpyFromSymbol(host_out, osm::g_conv_trace, (size_t)n_words * sizeof(unsigned long long));
}
extern "C" int osm_dbg_trace_flags(int flags) { return (int)cudaMemcpyToSymbol(osm::g_trace_flags, &flags, sizeof(int)); }
extern "C" int osm_dbg_trace_clear() {
  static unsigned long long zeros[256 * 16 * 2];
  return (int)cudaMemcpyToSymbol(osm::g_conv_trace, zeros, sizeof(zeros));
}
#endif
