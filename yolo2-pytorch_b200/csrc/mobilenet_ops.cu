// MobileNet plugin kernels (BASELINE configs[4], "C5"; /root/reference model/mobilenet.py:25-85), inference only.
//   mb_conv0   conv_bn(3, 32, stride 2): nn.Conv2d(3,32,3,2,1) + BatchNorm2d + ReLU  (:25-30), fp32 NCHW image in,
//              fp16 NHWC out -- the layout boundary of this backbone.
//   dwconv3x3  conv_dw: depthwise nn.Conv2d(C,C,3,stride,1,groups=C) + BatchNorm2d + ReLU (:33-38) on fp16 NHWC.
//              HBM-bound: one thread = 8 channels (16 B) of one output pixel, 9 vector loads, fp32 FMA, 16 B store.
// The pointwise convs (conv_pw, :41-46) and the 1x1 head reuse the tcgen05 implicit-GEMM kernel (slope = 0 -> ReLU).
#include "yb_common.h"
#include <cuda_fp16.h>
#include <stdint.h>

namespace yb {

__global__ void __launch_bounds__(256) mb_conv0_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, __half* __restrict__ y, int batch, int height, int width) {
  __shared__ __align__(16) float ws[27][32];
  __shared__ float sc[32], sh[32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) ws[i / 32][i % 32] = w[(i % 32) * 27 + i / 32];
  if (threadIdx.x < 32) { sc[threadIdx.x] = scale[threadIdx.x]; sh[threadIdx.x] = shift[threadIdx.x]; }
  __syncthreads();
  const int oh = height >> 1, ow = width >> 1;
  const long long total = static_cast<long long>(batch) * oh * ow;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int px = static_cast<int>(idx % ow);
  const long long t = idx / ow;
  const int py = static_cast<int>(t % oh);
  const int img = static_cast<int>(t / oh);
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = 2 * py - 1 + r;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int ix = 2 * px - 1 + s;
        const float v = (iy >= 0 && iy < height && ix >= 0 && ix < width) ? __ldg(x + ((static_cast<long long>(img) * 3 + c) * height + iy) * width + ix) : 0.f;
        const float4* wp = reinterpret_cast<const float4*>(&ws[c * 9 + r * 3 + s][0]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 wv = wp[q];
          acc[4 * q] = fmaf(v, wv.x, acc[4 * q]); acc[4 * q + 1] = fmaf(v, wv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v, wv.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v, wv.w, acc[4 * q + 3]);
        }
      }
    }
  uint4* dst = reinterpret_cast<uint4*>(y + idx * 32);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 pk;
    __half2* h = reinterpret_cast<__half2*>(&pk);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = q * 8 + 2 * e;
      h[e] = __floats2half2_rn(fmaxf(acc[c] * sc[c] + sh[c], 0.f), fmaxf(acc[c + 1] * sc[c + 1] + sh[c + 1], 0.f));
    }
    dst[q] = pk;
  }
}

int mb_conv0(const float* x, const float* w, const float* scale, const float* shift, void* y, int batch, int height, int width, cudaStream_t stream) {
  YB_REQUIRE(x && w && scale && shift && y && batch > 0 && height % 2 == 0 && width % 2 == 0, "mb_conv0: bad argument");
  const long long total = static_cast<long long>(batch) * (height / 2) * (width / 2);
  mb_conv0_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(x, w, scale, shift, reinterpret_cast<__half*>(y), batch, height, width);
  return check_launch("mb_conv0_kernel");
}

// w: fp32 [C][9] (the [C,1,3,3] depthwise weight), scale/shift: folded BN
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const __half* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, __half* __restrict__ y, int batch, int height, int width,
                                                        int channels, int stride) {
  const int c8 = channels >> 3;
  const int oh = height / stride, ow = width / stride;
  const long long total = static_cast<long long>(batch) * oh * ow * c8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % c8);
  long long t = idx / c8;
  const int px = static_cast<int>(t % ow); t /= ow;
  const int py = static_cast<int>(t % oh);
  const int img = static_cast<int>(t / oh);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const float* wc = w + static_cast<long long>(cg) * 8 * 9;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int iy = py * stride - 1 + r;
    if (iy < 0 || iy >= height) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int ix = px * stride - 1 + s;
      if (ix < 0 || ix >= width) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<long long>(img) * height + iy) * width + ix) * channels + cg * 8));
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        acc[2 * e] = fmaf(f.x, __ldg(wc + (2 * e) * 9 + r * 3 + s), acc[2 * e]);
        acc[2 * e + 1] = fmaf(f.y, __ldg(wc + (2 * e + 1) * 9 + r * 3 + s), acc[2 * e + 1]);
      }
    }
  }
  uint4 pk;
  __half2* h = reinterpret_cast<__half2*>(&pk);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = cg * 8 + 2 * e;
    h[e] = __floats2half2_rn(fmaxf(acc[2 * e] * __ldg(scale + c) + __ldg(shift + c), 0.f),
                             fmaxf(acc[2 * e + 1] * __ldg(scale + c + 1) + __ldg(shift + c + 1), 0.f));
  }
  reinterpret_cast<uint4*>(y)[idx] = pk;
}

int dwconv3x3(const void* x, const float* w, const float* scale, const float* shift, void* y, int batch, int height, int width, int channels,
              int stride, cudaStream_t stream) {
  YB_REQUIRE(x && w && scale && shift && y && batch > 0 && channels % 8 == 0 && (stride == 1 || stride == 2) && height % stride == 0 &&
                 width % stride == 0,
             "dwconv3x3: bad argument");
  const long long total = static_cast<long long>(batch) * (height / stride) * (width / stride) * (channels / 8);
  dwconv3x3_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(x), w, scale, shift,
                                                                                   reinterpret_cast<__half*>(y), batch, height, width, channels, stride);
  return check_launch("dwconv3x3_kernel");
}

}  // namespace yb
