// fp32 CUDA-core implicit-GEMM convolution (3x3 pad 1 / 1x1, stride 1) on NHWC views.
// This is the EXACT-fp32 mode of the engine (conv_mode = 1): same packed weights, same epilogue and the
// same call sites as the tcgen05 kernel, used by the parity tests to separate logic errors from TF32
// rounding and as the on-device cross-check of the tensor-core path.  It is not the product path.
#include "conv_epilogue.cuh"

namespace osm {

constexpr int CS_BM = 64, CS_BN = 64, CS_BK = 16;

__global__ void __launch_bounds__(256) conv_simt_kernel(ConvArgs a) {
  __shared__ float As[CS_BK][CS_BM + 4];
  __shared__ float Bs[CS_BK][CS_BN + 4];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const long M = (long)a.B * a.H * a.W;
  const long m0 = (long)blockIdx.x * CS_BM;
  const int n0 = blockIdx.y * CS_BN;

  // the pixel / weight row this thread stages
  const int lpx = tid / 4, lcv = tid % 4;
  const long lm = m0 + lpx;
  const bool lm_ok = lm < M;
  int lb = 0, lh = 0, lw = 0;
  if (lm_ok) {
    lw = (int)(lm % a.W);
    lh = (int)((lm / a.W) % a.H);
    lb = (int)(lm / ((long)a.W * a.H));
  }
  const int lco = n0 + lpx;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < a.taps; ++tap) {
    const int dy = a.taps == 9 ? tap / 3 - 1 : 0, dx = a.taps == 9 ? tap % 3 - 1 : 0;
    const int sh = lh + dy, sw = lw + dx;
    const bool px_ok = lm_ok && sh >= 0 && sh < a.H && sw >= 0 && sw < a.W;
    const float* xrow = a.x + (((size_t)lb * a.H + sh) * a.W + sw) * a.ldx;
    const int kc_n = a.Cin_p / 32;  // packed weights: [tap][ci/32][co][ci%32]
    for (int c0 = 0; c0 < a.Cin_p; c0 += CS_BK) {
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
      if (px_ok) av = *reinterpret_cast<const float4*>(xrow + c0 + lcv * 4);
      if (lco < a.Cout_p)
        bv = *reinterpret_cast<const float4*>(a.w + (((size_t)tap * kc_n + c0 / 32) * a.Cout_p + lco) * 32 + (c0 & 31) + lcv * 4);
      As[lcv * 4 + 0][lpx] = av.x; As[lcv * 4 + 1][lpx] = av.y; As[lcv * 4 + 2][lpx] = av.z; As[lcv * 4 + 3][lpx] = av.w;
      Bs[lcv * 4 + 0][lpx] = bv.x; Bs[lcv * 4 + 1][lpx] = bv.y; Bs[lcv * 4 + 2][lpx] = bv.z; Bs[lcv * 4 + 3][lpx] = bv.w;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < CS_BK; ++k) {
        const float4 p = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 q = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float pv[4] = {p.x, p.y, p.z, p.w}, qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(pv[i], qv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  EpiArgs e{a.bias, a.res, a.ldr, a.res_mode, a.out, a.ldo, a.accumulate, a.H, a.W};
  const int co = n0 + tx * 4;
  if (co >= a.Cout_p) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int w = (int)(m % a.W), h = (int)((m / a.W) % a.H), b = (int)(m / ((long)a.W * a.H));
    conv_epilogue_store4(e, b, h, w, co, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
  }
}

int conv_check(const ConvArgs& a) {
  if (a.taps != 1 && a.taps != 9) return fail(OSM_ERR_INVALID, "conv: taps must be 1 or 9");
  if (a.Cin_p % 32 || a.Cout_p % 32) return fail(OSM_ERR_INVALID, "conv: padded channel counts must be multiples of 32");
  if (a.ldx % 4 || a.ldo % 4 || (a.res_mode != RES_NONE && a.ldr % 4)) return fail(OSM_ERR_INVALID, "conv: ld must be a multiple of 4");
  if (a.res_mode == RES_NEAREST_UP && ((a.H | a.W) & 1)) return fail(OSM_ERR_INVALID, "conv: odd size with upsampled residual");
  return OSM_OK;
}

int conv_simt_launch(const ConvArgs& a, cudaStream_t s) {
  if (int e = conv_check(a)) return e;
  const long M = (long)a.B * a.H * a.W;
  dim3 grid((unsigned)((M + CS_BM - 1) / CS_BM), (a.Cout_p + CS_BN - 1) / CS_BN);
  conv_simt_kernel<<<grid, 256, 0, s>>>(a);
  OSM_LAUNCH_CHECK("conv_simt_kernel");
  return OSM_OK;
}

}  // namespace osm
