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
};

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// (b,h,w): output pixel; co: first of 4 consecutive output channels; v: accumulator values
__device__ __forceinline__ void conv_epilogue_store4(const EpiArgs& e, int b, int h, int w, int co, float4 v) {
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
}

}  // namespace osm
