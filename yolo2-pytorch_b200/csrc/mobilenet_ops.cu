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

// w: fp32 [C][9] (the [C,1,3,3] depthwise weight), scale/shift: folded BN.
// One thread = 8 channels (16 B) of TX consecutive output pixels of one row.  Its 72 weights are 288 contiguous bytes (18 x LDG.128,
// once), the 3 x 3 input window slides along the row in registers (3 new 16 B loads per output pixel at stride 1, 6 at stride 2), so
// the kernel issues ~1/4 of the load instructions of the one-pixel-per-thread form (which ncu showed issue-bound at 0.10 of HBM peak).
template <int STRIDE, int TX>
__global__ void __launch_bounds__(128) dwconv3x3_kernel(const __half* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, __half* __restrict__ y, int batch, int height, int width,
                                                        int channels) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int c8 = channels >> 3;
  const int oh = height / STRIDE, ow = width / STRIDE;
  const int strips = (ow + TX - 1) / TX;
  const long long total = static_cast<long long>(batch) * oh * strips * c8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % c8);
  long long t = idx / c8;
  const int strip = static_cast<int>(t % strips); t /= strips;
  const int py = static_cast<int>(t % oh);
  const int img = static_cast<int>(t / oh);
  float wr[72];
  {
    const float4* wp = reinterpret_cast<const float4*>(w + static_cast<long long>(cg) * 72);
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      const float4 v = __ldg(wp + i);
      wr[4 * i] = v.x; wr[4 * i + 1] = v.y; wr[4 * i + 2] = v.z; wr[4 * i + 3] = v.w;
    }
  }
  float sc[8], sh[8];
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(scale + cg * 8)), a1 = __ldg(reinterpret_cast<const float4*>(scale + cg * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + cg * 8)), b1 = __ldg(reinterpret_cast<const float4*>(shift + cg * 8) + 1);
    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
    sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
  }
  const int px0 = strip * TX;
  const int iy0 = py * STRIDE - 1;
  const __half* xrow[3];
  bool rok[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int iy = iy0 + r;
    rok[r] = iy >= 0 && iy < height;
    xrow[r] = x + ((static_cast<long long>(img) * height + (rok[r] ? iy : 0)) * width) * channels + cg * 8;
  }
  auto load_col = [&](int ix, uint4 (&col)[3]) {
    const bool cok = ix >= 0 && ix < width;
#pragma unroll
    for (int r = 0; r < 3; ++r)
      col[r] = (cok && rok[r]) ? __ldg(reinterpret_cast<const uint4*>(xrow[r] + static_cast<long long>(ix) * channels)) : make_uint4(0u, 0u, 0u, 0u);
  };
  uint4 win[3][3];                                   // [column s][row r]
  load_col(px0 * STRIDE - 1, win[0]);
  if (STRIDE == 1) load_col(px0 * STRIDE, win[1]);
#pragma unroll
  for (int j = 0; j < TX; ++j) {
    const int px = px0 + j;
    if (px >= ow) break;
    if (STRIDE == 1) {
      load_col(px + 1, win[2]);
    } else {
      load_col(2 * px, win[1]);
      load_col(2 * px + 1, win[2]);
    }
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const __half2* h = reinterpret_cast<const __half2*>(&win[s][r]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          acc[2 * e] = fmaf(f.x, wr[(2 * e) * 9 + r * 3 + s], acc[2 * e]);
          acc[2 * e + 1] = fmaf(f.y, wr[(2 * e + 1) * 9 + r * 3 + s], acc[2 * e + 1]);
        }
      }
    uint4 pk;
    __half2* ho = reinterpret_cast<__half2*>(&pk);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      ho[e] = __floats2half2_rn(fmaxf(acc[2 * e] * sc[2 * e] + sh[2 * e], 0.f), fmaxf(acc[2 * e + 1] * sc[2 * e + 1] + sh[2 * e + 1], 0.f));
    *reinterpret_cast<uint4*>(y + ((static_cast<long long>(img) * oh + py) * ow + px) * channels + cg * 8) = pk;
    // slide: stride 1 keeps two columns, stride 2 keeps one
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (STRIDE == 1) { win[0][r] = win[1][r]; win[1][r] = win[2][r]; }
      else win[0][r] = win[2][r];
    }
  }
}

int dwconv3x3(const void* x, const float* w, const float* scale, const float* shift, void* y, int batch, int height, int width, int channels,
              int stride, cudaStream_t stream) {
  YB_REQUIRE(x && w && scale && shift && y && batch > 0 && channels % 8 == 0 && (stride == 1 || stride == 2) && height % stride == 0 &&
                 width % stride == 0,
             "dwconv3x3: bad argument");
  YB_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(scale) & 15) == 0 && (reinterpret_cast<uintptr_t>(shift) & 15) == 0,
             "dwconv3x3: weights / scale / shift must be 16 B aligned");
  const int oh = height / stride, ow = width / stride;
  // pixels per thread: long strips amortise the 72-weight preload, but the grid must still cover the SMs on the 13 x 13 layers
  const int tx = ow >= 52 ? 8 : 4;
  const long long total = static_cast<long long>(batch) * oh * ((ow + tx - 1) / tx) * (channels / 8);
  const unsigned grid = static_cast<unsigned>((total + 127) / 128);
  const __half* xp = reinterpret_cast<const __half*>(x);
  __half* yp = reinterpret_cast<__half*>(y);
  if (stride == 1) {
    if (tx == 8) dwconv3x3_kernel<1, 8><<<grid, 128, 0, stream>>>(xp, w, scale, shift, yp, batch, height, width, channels);
    else dwconv3x3_kernel<1, 4><<<grid, 128, 0, stream>>>(xp, w, scale, shift, yp, batch, height, width, channels);
  } else {
    if (tx == 8) dwconv3x3_kernel<2, 8><<<grid, 128, 0, stream>>>(xp, w, scale, shift, yp, batch, height, width, channels);
    else dwconv3x3_kernel<2, 4><<<grid, 128, 0, stream>>>(xp, w, scale, shift, yp, batch, height, width, channels);
  }
  return check_launch("dwconv3x3_kernel");
}

}  // namespace yb
