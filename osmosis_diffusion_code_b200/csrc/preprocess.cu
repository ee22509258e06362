// Input pipeline on the device.  Replaces, per image, the torchvision transform chain the reference
// builds in osmosis_sampling.py:46-49
//     ToTensor() -> Resize(256) -> CenterCrop([256, 256]) -> Normalize(0.5, 0.5)
// and the optional de-gamma of the measurement (osmosis_sampling.py:170-175), for a decoded uint8 HWC image that is already
// in device memory.  Resize on a tensor is F.interpolate(mode='bilinear', align_corners=False, antialias=True): a separable
// triangle filter whose support grows with the down-scaling factor, horizontal pass first.  Only the pixels the centre crop
// keeps are computed.
//
//   preprocess_h_kernel   u8 [H][W][3] -> f32 tmp [3][H][S]   (x / 255, horizontal taps, crop columns only)
//   preprocess_v_kernel   tmp -> f32 out [3][S][S]             (vertical taps, crop rows only, (v - 0.5) / 0.5, de-gamma)
//
// Arithmetic mirrors ATen's CPU kernel op by op (float variables, double literals, one fused multiply-add per tap in tap
// order), so the result is bit-identical to oracle.preprocess_image (numpy), which is pinned to torchvision's own output
// (bit-exact for the golden images with short filters, <= 2.4e-7 for the 1080p one where ATen vectorises the tap loop).
// HBM-bound byte work: the source image is read once (taps of neighbouring outputs overlap in L1), 3*H*S floats of
// intermediate are written and read once.
#include <math.h>
#include <stdint.h>

#include "common.cuh"

namespace osm {

namespace {

struct AxisP {
  float scale;     // in / out
  float support;   // scale if down-scaling else 1
  float invscale;  // 1 / scale if down-scaling else 1
  int in_size;
  int offset;      // first kept output index (centre crop)
};

// weights of output index i: taps [xmin, xmin + xsize), w_j = tri((j + xmin - center + 0.5) * invscale) / sum
__device__ __forceinline__ void axis_window(const AxisP& a, int i, int* xmin, int* xsize, float* center) {
  // ATen (UpSampleKernel.cpp, _compute_indices_min_size_weights_aa with scalar_t = float) mixes float variables with
  // double literals: the products / sums below are formed in double and rounded once where ATen stores a float.
  const float c = (float)((double)a.scale * ((double)i + 0.5));
  const int lo = max((int)((double)__fsub_rn(c, a.support) + 0.5), 0);
  const int hi = min((int)((double)__fadd_rn(c, a.support) + 0.5), a.in_size);
  *xmin = lo;
  *xsize = hi - lo;
  *center = c;
}
__device__ __forceinline__ float tri(const AxisP& a, int j, float center) {
  const float x = (float)(((double)__fsub_rn((float)j, center) + 0.5) * (double)a.invscale);
  return fmaxf(0.0f, __fsub_rn(1.0f, fabsf(x)));
}

constexpr int PRE_THREADS = 256;

// grid (ceil(S / 256), H, 3)
__global__ void __launch_bounds__(PRE_THREADS)
preprocess_h_kernel(const uint8_t* __restrict__ src, int pitch, int Cs, AxisP ax, float* __restrict__ tmp, int H, int S) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y, c = blockIdx.z;
  if (x >= S) return;
  int xmin, xsize;
  float center;
  axis_window(ax, x + ax.offset, &xmin, &xsize, &center);
  float total = 0.0f;
  for (int j = 0; j < xsize; ++j) total = __fadd_rn(total, tri(ax, j + xmin, center));
  const uint8_t* row = src + (size_t)h * pitch + (Cs == 1 ? 0 : c);
  float acc = 0.0f;
  for (int j = 0; j < xsize; ++j) {
    const float w = __fdiv_rn(tri(ax, j + xmin, center), total);
    const float v = __fdiv_rn((float)row[(size_t)(xmin + j) * Cs], 255.0f);   // ToTensor: float(u8) / 255
    acc = (j == 0) ? __fmul_rn(v, w) : __fmaf_rn(v, w, acc);
  }
  tmp[((size_t)c * H + h) * S + x] = acc;
}

// grid (ceil(S / 256), S, 3)
__global__ void __launch_bounds__(PRE_THREADS)
preprocess_v_kernel(const float* __restrict__ tmp, AxisP ay, float* __restrict__ out, int H, int S, int degamma) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, c = blockIdx.z;
  if (x >= S) return;
  int ymin, ysize;
  float center;
  axis_window(ay, y + ay.offset, &ymin, &ysize, &center);
  float total = 0.0f;
  for (int j = 0; j < ysize; ++j) total = __fadd_rn(total, tri(ay, j + ymin, center));
  const float* col = tmp + (size_t)c * H * S + x;
  float acc = 0.0f;
  for (int j = 0; j < ysize; ++j) {
    const float w = __fdiv_rn(tri(ay, j + ymin, center), total);
    const float v = col[(size_t)(ymin + j) * S];
    acc = (j == 0) ? __fmul_rn(v, w) : __fmaf_rn(v, w, acc);
  }
  float v = __fdiv_rn(__fsub_rn(acc, 0.5f), 0.5f);                              // Normalize(0.5, 0.5)
  if (degamma) v = __fsub_rn(__fmul_rn(2.0f, powf(__fmul_rn(0.5f, __fadd_rn(v, 1.0f)), 2.2f)), 1.0f);
  out[((size_t)c * S + y) * S + x] = v;
}

__global__ void degamma_kernel(const float* __restrict__ y, float* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __fsub_rn(__fmul_rn(2.0f, powf(__fmul_rn(0.5f, __fadd_rn(y[i], 1.0f)), 2.2f)), 1.0f);
}

AxisP make_axis(int in_size, int out_size, int offset) {
  AxisP a;
  a.scale = (float)in_size / (float)out_size;
  a.support = a.scale >= 1.0f ? a.scale : 1.0f;
  a.invscale = a.scale >= 1.0f ? (float)(1.0 / (double)a.scale) : 1.0f;
  a.in_size = in_size;
  a.offset = offset;
  return a;
}

}  // namespace

// torchvision's sizes: the short side becomes S, the long side int(S * long / short); the centre crop starts at
// round((n - S) / 2) with Python's round-half-to-even.
int preprocess_launch(const uint8_t* src, int H, int W, int Cs, int pitch, float* tmp, float* out, int S, int degamma,
                      cudaStream_t s) {
  if (H <= 0 || W <= 0 || S <= 0 || (Cs != 1 && Cs != 3)) return fail(OSM_ERR_INVALID, "preprocess: bad image shape");
  if (pitch < W * Cs) return fail(OSM_ERR_INVALID, "preprocess: row pitch smaller than a row");
  int nh, nw;
  if (H <= W) { nh = S; nw = (int)((long long)S * W / H); } else { nw = S; nh = (int)((long long)S * H / W); }
  auto half_even = [](int d) { return (d % 2 == 0) ? d / 2 : ((d / 2) % 2 == 0 ? d / 2 : d / 2 + 1); };
  const int top = half_even(nh - S), left = half_even(nw - S);
  const AxisP ax = make_axis(W, nw, left), ay = make_axis(H, nh, top);
  dim3 gh((S + PRE_THREADS - 1) / PRE_THREADS, H, 3), gv((S + PRE_THREADS - 1) / PRE_THREADS, S, 3);
  if (H > 65535) return fail(OSM_ERR_INVALID, "preprocess: image taller than 65535 rows");
  preprocess_h_kernel<<<gh, PRE_THREADS, 0, s>>>(src, pitch, Cs, ax, tmp, H, S);
  preprocess_v_kernel<<<gv, PRE_THREADS, 0, s>>>(tmp, ay, out, H, S, degamma);
  OSM_CUDA_CHECK(cudaGetLastError());
  return OSM_OK;
}

int degamma_launch(const float* y, float* out, size_t n, cudaStream_t s) {
  if (n == 0) return OSM_OK;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
  degamma_kernel<<<blocks, 256, 0, s>>>(y, out, n);
  OSM_CUDA_CHECK(cudaGetLastError());
  return OSM_OK;
}

}  // namespace osm
