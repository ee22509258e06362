// Epilogue shared by the tensor-core and the CUDA-core conv kernels: bias, residual (identity skip of a
// ResBlock incl. the AvgPool2d / nearest-up applied to the skip path, unet.py:318-320 & :335), optional
// accumulation (gradient fan-in), 128-bit store into an NHWC view.
#pragma once
#include "common.cuh"

namespace osm {

struct EpiArgs {
  const float* bias;
  const float* res; int ldr; int res_mode;
  float* out; int ldo;
  int accumulate;
  int H, W;  // output spatial size
  // GroupNorm statistics of the values this conv stores, reduced in the epilogue (persistent tcgen05 kernel, tiles
  // inside one image): the consumer GroupNorm then needs no pass of its own over the tensor.
  //   mode 1 (forward):  per (image, group)  sum v, sum v^2                      -> mean / rstd        (nn.py:17-19)
  //   mode 2 (backward): v is dL/d(activation); with x the GroupNorm input,  d = v silu'(x a + b) e,
  //                      sum d, sum d x  -> the two means of the GroupNorm input gradient
  int stat_mode;
  int stat_cpg;                  // channels per group: 4, 8, 16 or 32
  float* stat_partial;           // [pixel tile][4 epilogue warps][32 groups][2], every entry written exactly once
  const float* stat_x; int stat_ldx;
  const float4* stat_coef;       // mode 2: [B][C] (a, b, e, 0)
  int stat_silu;
};

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// (b,h,w): output pixel; co: first of 4 consecutive output channels; v: accumulator values
__device__ __forceinline__ float4 conv_epilogue_store4(const EpiArgs& e, int b, int h, int w, int co, float4 v) {
  if (e.bias) v = f4_add(v, *reinterpret_cast<const float4*>(e.bias + co));
  if (e.res_mode == RES_SAME) {
    v = f4_add(v, *reinterpret_cast<const float4*>(e.res + (((size_t)b * e.H + h) * e.W + w) * e.ldr + co));
  } else if (e.res_mode == RES_AVGPOOL) {
    const int Hs = e.H * 2, Ws = e.W * 2;
    const float* base = e.res + (((size_t)b * Hs + 2 * h) * Ws + 2 * w) * e.ldr + co;
    const float4 r0 = *reinterpret_cast<const float4*>(base);
    const float4 r1 = *reinterpret_cast<const float4*>(base + e.ldr);
    const float4 r2 = *reinterpret_cast<const float4*>(base + (size_t)Ws * e.ldr);
    const float4 r3 = *reinterpret_cast<const float4*>(base + (size_t)Ws * e.ldr + e.ldr);
    const float4 s = f4_add(f4_add(r0, r1), f4_add(r2, r3));
    v = f4_add(v, make_float4(0.25f * s.x, 0.25f * s.y, 0.25f * s.z, 0.25f * s.w));
  } else if (e.res_mode == RES_NEAREST_UP) {
    const int Hs = e.H / 2, Ws = e.W / 2;
    v = f4_add(v, *reinterpret_cast<const float4*>(e.res + (((size_t)b * Hs + h / 2) * Ws + w / 2) * e.ldr + co));
  }
  float4* dst = reinterpret_cast<float4*>(e.out + (((size_t)b * e.H + h) * e.W + w) * e.ldo + co);
  if (e.accumulate) v = f4_add(v, *dst);
  *dst = v;
  return v;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// conv_epilogue_store4 in two phases, so that a caller can have the operands of several float4s in flight before the first store
// (stores to `out` may alias the loads as far as the compiler knows).  Same additions in the same order: same bits.
struct EpiOperands4 { float4 bias, res, prev; };
__device__ __forceinline__ EpiOperands4 conv_epilogue_load4(const EpiArgs& e, int b, int h, int w, int co) {
  EpiOperands4 o;
  o.bias = o.res = o.prev = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e.bias) o.bias = ld4(e.bias + co);
  if (e.res_mode == RES_SAME) {
    o.res = ld4(e.res + (((size_t)b * e.H + h) * e.W + w) * e.ldr + co);
  } else if (e.res_mode == RES_AVGPOOL) {
    const int Hs = e.H * 2, Ws = e.W * 2;
    const float* base = e.res + (((size_t)b * Hs + 2 * h) * Ws + 2 * w) * e.ldr + co;
    const float4 r0 = ld4(base), r1 = ld4(base + e.ldr), r2 = ld4(base + (size_t)Ws * e.ldr), r3 = ld4(base + (size_t)Ws * e.ldr + e.ldr);
    const float4 s = f4_add(f4_add(r0, r1), f4_add(r2, r3));
    o.res = make_float4(0.25f * s.x, 0.25f * s.y, 0.25f * s.z, 0.25f * s.w);
  } else if (e.res_mode == RES_NEAREST_UP) {
    const int Hs = e.H / 2, Ws = e.W / 2;
    o.res = ld4(e.res + (((size_t)b * Hs + h / 2) * Ws + w / 2) * e.ldr + co);
  }
  if (e.accumulate) o.prev = ld4(e.out + (((size_t)b * e.H + h) * e.W + w) * e.ldo + co);
  return o;
}
__device__ __forceinline__ void conv_epilogue_apply_store4(const EpiArgs& e, int b, int h, int w, int co, float4 v, const EpiOperands4& o) {
  if (e.bias) v = f4_add(v, o.bias);
  if (e.res_mode != RES_NONE) v = f4_add(v, o.res);
  if (e.accumulate) v = f4_add(v, o.prev);
  *reinterpret_cast<float4*>(e.out + (((size_t)b * e.H + h) * e.W + w) * e.ldo + co) = v;
}

// One pixel x 32 consecutive output channels (one tcgen05.ld chunk `r`): bias / residual / accumulate / store, and the
// per-slot partial sums st[16] = {s, q} x 8 four-channel slots of the fused GroupNorm statistics (untouched when off).
// Written as straight-line phases - all loads of a phase are issued before their first use - because the per-float4
// form (load, add, store, repeat) serialises eight global-memory round trips per chunk.
__device__ __forceinline__ void conv_epilogue_chunk32(const EpiArgs& e, int b, int h, int w, int co, int Cout_p, const uint32_t (&r)[32],
                                                      float (&st)[16]) {
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    v[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
  const size_t pix = ((size_t)b * e.H + h) * e.W + w;
  float4* dst = reinterpret_cast<float4*>(e.out + pix * e.ldo + co);
  float4 acc[8];
  if (e.accumulate) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = dst[i];
  }
  if (e.res_mode == RES_SAME) {
    const float* rp = e.res + pix * e.ldr + co;
    float4 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = ld4(rp + 4 * i);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = f4_add(v[i], t[i]);
  } else if (e.res_mode == RES_AVGPOOL) {
    const int Ws = e.W * 2;
    const float* base = e.res + (((size_t)b * e.H * 2 + 2 * h) * Ws + 2 * w) * e.ldr + co;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float4 t[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float* q = base + 4 * (4 * hf + i);
        t[i][0] = ld4(q); t[i][1] = ld4(q + e.ldr); t[i][2] = ld4(q + (size_t)Ws * e.ldr); t[i][3] = ld4(q + (size_t)Ws * e.ldr + e.ldr);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 s4 = f4_add(f4_add(t[i][0], t[i][1]), f4_add(t[i][2], t[i][3]));
        v[4 * hf + i] = f4_add(v[4 * hf + i], make_float4(0.25f * s4.x, 0.25f * s4.y, 0.25f * s4.z, 0.25f * s4.w));
      }
    }
  } else if (e.res_mode == RES_NEAREST_UP) {
    const float* rp = e.res + (((size_t)b * (e.H / 2) + h / 2) * (e.W / 2) + w / 2) * e.ldr + co;
    float4 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = ld4(rp + 4 * i);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = f4_add(v[i], t[i]);
  }
  if (e.bias) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = f4_add(v[i], __ldg(reinterpret_cast<const float4*>(e.bias + co) + i));
  }
  if (e.accumulate) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = f4_add(v[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) dst[i] = v[i];
  if (e.stat_mode == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      st[2 * i] = (v[i].x + v[i].y) + (v[i].z + v[i].w);
      st[2 * i + 1] = (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  } else if (e.stat_mode == 2) {
    const float* xp = e.stat_x + pix * e.stat_ldx + co;
    float4 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __ldcs(reinterpret_cast<const float4*>(xp) + i);  // streamed once: evict-first, keep L2 for the A tiles
    const float4* cf = e.stat_coef + (size_t)b * Cout_p + co;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 c0 = __ldg(cf + 4 * i), c1 = __ldg(cf + 4 * i + 1), c2 = __ldg(cf + 4 * i + 2), c3 = __ldg(cf + 4 * i + 3);
      const float xs[4] = {x[i].x, x[i].y, x[i].z, x[i].w}, gs[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      const float ca[4] = {c0.x, c1.x, c2.x, c3.x}, cb[4] = {c0.y, c1.y, c2.y, c3.y}, ce[4] = {c0.z, c1.z, c2.z, c3.z};
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float d = gs[k] * ce[k];
        if (e.stat_silu) {
          const float u = xs[k] * ca[k] + cb[k];
          const float sg = __fdividef(1.0f, 1.0f + __expf(-u));
          d *= sg * (1.0f + u * (1.0f - sg));
        }
        s0 += d;
        s1 += d * xs[k];
      }
      st[2 * i] = s0;
      st[2 * i + 1] = s1;
    }
  }
}

// The same contract as conv_epilogue_chunk32 with a smaller live set, for the 16-warp fp16 kernel (128 registers per thread): the
// accumulator registers are updated in place and every phase (residual, bias, `+=`, statistics) loads at most 16 floats at a
// time, so nothing spills (a spilled scalar costs an L2 round trip there).  The phases of a chunk serialise a few more L2 round
// trips than the wide version; eight epilogue warps and the L2 prefetches issued before the accumulator is ready hide them.
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one instruction per 8 floats of a pixel row.  A thread owns a row here, so
// every access of a warp instruction touches its own 128-byte line (32 L1 wavefronts per instruction whatever its width): twice the
// width is half the load / store instructions and half the LSU time of the epilogue.  Addresses are 32-byte aligned (pixel strides
// and channel offsets are multiples of 8 floats).
struct float8 { float4 lo, hi; };
__device__ __forceinline__ float8 ld8(const float* p) {
  float8 v;
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v.lo.x), "=f"(v.lo.y), "=f"(v.lo.z), "=f"(v.lo.w), "=f"(v.hi.x), "=f"(v.hi.y), "=f"(v.hi.z), "=f"(v.hi.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float8 ld8_stream(const float* p) {   // read once: evict-first
  float8 v;
  asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v.lo.x), "=f"(v.lo.y), "=f"(v.lo.z), "=f"(v.lo.w), "=f"(v.hi.x), "=f"(v.hi.y), "=f"(v.hi.z), "=f"(v.hi.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st8(float* p, float4 a, float4 b) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y),
               "f"(b.z), "f"(b.w) : "memory");
}

// coef_sm != null (mode 2): the (a, b, e, 0) coefficients of this chunk's 32 channels staged in shared memory by the caller
// (read back as warp-wide broadcasts) instead of 32 global loads per chunk and thread.
__device__ __forceinline__ void conv_epilogue_stat2_quad(const float4 x, const float4 v, const float4 c0, const float4 c1, const float4 c2,
                                                         const float4 c3, int silu, float& s0, float& s1) {
  const float xs[4] = {x.x, x.y, x.z, x.w}, gs[4] = {v.x, v.y, v.z, v.w};
  const float ca[4] = {c0.x, c1.x, c2.x, c3.x}, cb[4] = {c0.y, c1.y, c2.y, c3.y}, ce[4] = {c0.z, c1.z, c2.z, c3.z};
  s0 = 0.f; s1 = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float d = gs[k] * ce[k];
    if (silu) {
      const float u = xs[k] * ca[k] + cb[k];
      const float sg = __fdividef(1.0f, 1.0f + __expf(-u));
      d *= sg * (1.0f + u * (1.0f - sg));
    }
    s0 += d;
    s1 += d * xs[k];
  }
}

// xs (mode 2): the GroupNorm input of this row's first 16 channels, loaded by the caller BEFORE it waits for the accumulator (so that
// the L2 / DRAM round trip overlaps the tcgen05.ld and the stores); null: loaded here.
__device__ __forceinline__ void conv_epilogue_chunk32_lean(const EpiArgs& e, int b, int h, int w, int co, int Cout_p, uint32_t (&r)[32],
                                                           float (&st)[16], const float4* coef_sm = nullptr, const float8* xs = nullptr) {
  const size_t pix = ((size_t)b * e.H + h) * e.W + w;
  float* dst = e.out + pix * e.ldo + co;
#define OSM_V(i) make_float4(__uint_as_float(r[4 * (i)]), __uint_as_float(r[4 * (i) + 1]), __uint_as_float(r[4 * (i) + 2]), __uint_as_float(r[4 * (i) + 3]))
#define OSM_SETV(i, q) do { r[4 * (i)] = __float_as_uint((q).x); r[4 * (i) + 1] = __float_as_uint((q).y); r[4 * (i) + 2] = __float_as_uint((q).z); r[4 * (i) + 3] = __float_as_uint((q).w); } while (0)
#define OSM_ADD8(i2, t8) do { const float4 qa = f4_add(OSM_V(2 * (i2)), (t8).lo), qb = f4_add(OSM_V(2 * (i2) + 1), (t8).hi); OSM_SETV(2 * (i2), qa); OSM_SETV(2 * (i2) + 1, qb); } while (0)
  // mode 2: the first half of the GroupNorm input row is requested before anything else (streamed once: evict-first)
  const float* xp = e.stat_mode == 2 ? e.stat_x + pix * e.stat_ldx + co : nullptr;
  float8 x0[2];
  if (e.stat_mode == 2 && !xs) {
#pragma unroll
    for (int i = 0; i < 2; ++i) x0[i] = ld8_stream(xp + 8 * i);
  }
  if (e.res_mode == RES_SAME || e.res_mode == RES_NEAREST_UP) {
    const float* rp = e.res_mode == RES_SAME ? e.res + pix * e.ldr + co
                                             : e.res + (((size_t)b * (e.H / 2) + h / 2) * (e.W / 2) + w / 2) * e.ldr + co;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float8 t[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) t[i] = ld8(rp + 16 * hf + 8 * i);
#pragma unroll
      for (int i = 0; i < 2; ++i) OSM_ADD8(2 * hf + i, t[i]);
    }
  } else if (e.res_mode == RES_AVGPOOL) {
    const int Ws = e.W * 2;
    const float* base = e.res + (((size_t)b * e.H * 2 + 2 * h) * Ws + 2 * w) * e.ldr + co;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* q = base + 8 * i;
      float4 sa, sb;
      {
        const float8 t0 = ld8(q), t1 = ld8(q + e.ldr);
        sa = f4_add(t0.lo, t1.lo); sb = f4_add(t0.hi, t1.hi);
      }
      {
        const float8 t2 = ld8(q + (size_t)Ws * e.ldr), t3 = ld8(q + (size_t)Ws * e.ldr + e.ldr);
        sa = f4_add(sa, f4_add(t2.lo, t3.lo)); sb = f4_add(sb, f4_add(t2.hi, t3.hi));
      }
      float8 t;
      t.lo = make_float4(0.25f * sa.x, 0.25f * sa.y, 0.25f * sa.z, 0.25f * sa.w);
      t.hi = make_float4(0.25f * sb.x, 0.25f * sb.y, 0.25f * sb.z, 0.25f * sb.w);
      OSM_ADD8(i, t);
    }
  }
  if (e.bias) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float4 t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) t[i] = __ldg(reinterpret_cast<const float4*>(e.bias + co) + 4 * hf + i);
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float4 q = f4_add(OSM_V(4 * hf + i), t[i]); OSM_SETV(4 * hf + i, q); }
    }
  }
  if (e.accumulate) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float8 t[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) t[i] = ld8(dst + 16 * hf + 8 * i);
#pragma unroll
      for (int i = 0; i < 2; ++i) OSM_ADD8(2 * hf + i, t[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) st8(dst + 8 * i, OSM_V(2 * i), OSM_V(2 * i + 1));
  if (e.stat_mode == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 v = OSM_V(i);
      st[2 * i] = (v.x + v.y) + (v.z + v.w);
      st[2 * i + 1] = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
  } else if (e.stat_mode == 2) {
    float8 x1[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) x1[i] = ld8_stream(xp + 16 + 8 * i);   // second half: in flight while the first is reduced
    if (xs) {
#pragma unroll
      for (int i = 0; i < 2; ++i) x0[i] = xs[i];
    }
    const float4* cf = e.stat_coef + (size_t)b * Cout_p + co;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 c0, c1, c2, c3;
      if (coef_sm) { c0 = coef_sm[4 * i]; c1 = coef_sm[4 * i + 1]; c2 = coef_sm[4 * i + 2]; c3 = coef_sm[4 * i + 3]; }
      else { c0 = __ldg(cf + 4 * i); c1 = __ldg(cf + 4 * i + 1); c2 = __ldg(cf + 4 * i + 2); c3 = __ldg(cf + 4 * i + 3); }
      const float8& xx = i < 4 ? x0[(i & 3) >> 1] : x1[(i & 3) >> 1];
      conv_epilogue_stat2_quad((i & 1) ? xx.hi : xx.lo, OSM_V(i), c0, c1, c2, c3, e.stat_silu, st[2 * i], st[2 * i + 1]);
    }
  }
#undef OSM_V
#undef OSM_SETV
#undef OSM_ADD8
}

// Warp total of 16 per-thread values in 16 shuffles (halving exchange): afterwards EVERY lane L holds the total of
// element L >> 1.  Fixed order -> bit-reproducible.
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 16;
    const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 8;
    const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 4;
    const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const bool up = lane & 2;
    const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Folds the per-thread slot sums st[16] = {s, q} x 8 four-channel slots of one 32-channel chunk over the warp's 32 pixel
// rows and writes the per-group totals of this (tile, warp): dst -> [32 groups][2], first group of the chunk = g0.
__device__ __forceinline__ void conv_epilogue_stat_flush(float (&st)[16], int lane, int cpg, float* dst, int g0) {
  float t = warp_reduce16(st, lane);  // lane L: element L >> 1 = 2 * slot + component
  const int spg = cpg >> 2;           // 4-channel slots per group: 1, 2, 4 or 8
  for (int d = 1; d < spg; d <<= 1) t += __shfl_xor_sync(0xffffffffu, t, d * 4);
  const int slot = lane >> 2, comp = (lane >> 1) & 1;
  if ((lane & 1) == 0 && (slot & (spg - 1)) == 0) dst[(g0 + slot / spg) * 2 + comp] = t;
}

}  // namespace osm
