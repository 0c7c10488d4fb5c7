// HBM-bound layout / pointwise kernels of the Darknet-19 path (coalesced, 16-byte vectorised):
//   pack_weight      fp32 OIHW -> fp16 KRSC (the K-major B operand of the implicit GEMM)
//   bn_fold          eval-mode BatchNorm2d -> per-channel fp32 (scale, shift)      model/yolo2.py:58
//   conv0            layer "layers1.0": 3x3 Cin=3 conv + BN + leaky + 2x2 max-pool, reading the
//                    caller's fp32 NCHW image and writing fp16 NHWC              model/yolo2.py:78-79
//   maxpool2x2       nn.MaxPool2d(2) on fp16 NHWC                                 model/yolo2.py:79,86,97
//   reorg            space-to-depth, offset-major channel order                   model/yolo2.py:33-46
#include "yb_common.h"
#include <cuda_fp16.h>
#include <stdint.h>

namespace yb {

// ------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, __half* __restrict__ out, int cout, int cin, int k, int mode) {
  const long long total = static_cast<long long>(cout) * cin * k * k;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  if (mode == 0) {
    // out[co][r][s][ci] = w[co][ci][r][s]
    const int ci = static_cast<int>(idx % cin);
    long long t = idx / cin;
    const int s = static_cast<int>(t % k); t /= k;
    const int r = static_cast<int>(t % k);
    const int co = static_cast<int>(t / k);
    out[idx] = __float2half_rn(w[((static_cast<long long>(co) * cin + ci) * k + r) * k + s]);
  } else {
    // data-gradient operand: out[ci][r][s][co] = w[co][ci][k-1-r][k-1-s]  (rotated, in/out swapped)
    const int co = static_cast<int>(idx % cout);
    long long t = idx / cout;
    const int s = static_cast<int>(t % k); t /= k;
    const int r = static_cast<int>(t % k);
    const int ci = static_cast<int>(t / k);
    out[idx] = __float2half_rn(w[((static_cast<long long>(co) * cin + ci) * k + (k - 1 - r)) * k + (k - 1 - s)]);
  }
}

int pack_weight(const float* w, void* out, int cout, int cin, int k, int mode, cudaStream_t stream) {
  YB_REQUIRE(w && out && cout > 0 && cin > 0 && (k == 1 || k == 3) && (mode == 0 || mode == 1), "pack_weight: bad argument");
  const long long total = static_cast<long long>(cout) * cin * k * k;
  pack_weight_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(w, reinterpret_cast<__half*>(out), cout, cin, k, mode);
  return check_launch("pack_weight_kernel");
}

// ------------------------------------------------------------------------------------------
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, float* __restrict__ scale, float* __restrict__ shift, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float s = gamma[i] / sqrtf(var[i] + eps);
  scale[i] = s;
  shift[i] = beta[i] - mean[i] * s;
}

int bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale, float* shift, int c,
            cudaStream_t stream) {
  YB_REQUIRE(gamma && beta && mean && var && scale && shift && c > 0, "bn_fold: bad argument");
  bn_fold_kernel<<<(c + 127) / 128, 128, 0, stream>>>(gamma, beta, mean, var, eps, scale, shift, c);
  return check_launch("bn_fold_kernel");
}

// ------------------------------------------------------------------------------------------
// conv0: block = 128 pooled pixels (along x) x 2 channel halves (16 each); fp32 FMA, weights
// broadcast from shared memory as float4, the 4x4x3 input patch of a pool window in registers.
constexpr int kC0 = 32;

__global__ void __launch_bounds__(256) conv0_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                    const float* __restrict__ scale, const float* __restrict__ shift, float slope,
                                                    __half* __restrict__ y, int height, int width) {
  __shared__ __align__(16) float ws[27][kC0];  // [ci*9 + r*3 + s][co]
  __shared__ float sc[kC0], sh[kC0];
  for (int i = threadIdx.x; i < 27 * kC0; i += blockDim.x) {
    const int co = i % kC0, kk = i / kC0;     // w is OIHW: [co][ci][r][s] -> flat co*27 + kk
    ws[kk][co] = w[co * 27 + kk];
  }
  if (threadIdx.x < kC0) { sc[threadIdx.x] = scale[threadIdx.x]; sh[threadIdx.x] = shift[threadIdx.x]; }
  __syncthreads();
  const int ow = width >> 1, oh = height >> 1;
  const int half = threadIdx.x >> 7;                 // warp-uniform channel half
  const int px = blockIdx.x * 128 + (threadIdx.x & 127);
  const int py = blockIdx.y;
  const int img = blockIdx.z;
  if (px >= ow) return;
  float in[3][4][4];
  const float* xb = x + static_cast<long long>(img) * 3 * height * width;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int dy = 0; dy < 4; ++dy) {
      const int iy = 2 * py - 1 + dy;
      const bool yok = iy >= 0 && iy < height;
      const float* row = xb + (static_cast<long long>(c) * height + (yok ? iy : 0)) * width;
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        const int ix = 2 * px - 1 + dx;
        in[c][dy][dx] = (yok && ix >= 0 && ix < width) ? __ldg(row + ix) : 0.f;
      }
    }
  }
  float acc[4][16];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[p][j] = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const float4* wp = reinterpret_cast<const float4*>(&ws[c * 9 + r * 3 + s][half * 16]);
        float wv[16];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 t = wp[v];
          wv[4 * v] = t.x; wv[4 * v + 1] = t.y; wv[4 * v + 2] = t.z; wv[4 * v + 3] = t.w;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float xv = in[c][(p >> 1) + r][(p & 1) + s];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[p][j] = fmaf(xv, wv[j], acc[p][j]);
        }
      }
    }
  }
  uint32_t packed[8];
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    float o[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int co = half * 16 + j + e;
      float m = -INFINITY;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float v = acc[p][j + e] * sc[co] + sh[co];
        v = v > 0.f ? v : v * slope;
        m = fmaxf(m, v);
      }
      o[e] = m;
    }
    const __half2 h = __floats2half2_rn(o[0], o[1]);
    packed[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
  }
  uint4* dst = reinterpret_cast<uint4*>(y + ((static_cast<long long>(img) * oh + py) * ow + px) * kC0 + half * 16);
  dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
}

int conv0_forward(const float* x, const float* w, const float* scale, const float* shift, float slope, void* y, int batch, int height,
                  int width, int cout, cudaStream_t stream) {
  YB_REQUIRE(x && w && scale && shift && y, "conv0: null pointer");
  YB_REQUIRE(cout == kC0, "conv0: Cout=%d unsupported (32)", cout);
  YB_REQUIRE(batch > 0 && height > 0 && width > 0 && height % 2 == 0 && width % 2 == 0, "conv0: bad shape %dx%dx%d", batch, height, width);
  YB_REQUIRE(batch <= 65535 && height / 2 <= 65535, "conv0: grid too large");
  dim3 grid((width / 2 + 127) / 128, height / 2, batch);
  conv0_kernel<<<grid, 256, 0, stream>>>(x, w, scale, shift, slope, reinterpret_cast<__half*>(y), height, width);
  return check_launch("conv0_kernel");
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
  uint4 r;
  const __half2* pa = reinterpret_cast<const __half2*>(&a);
  const __half2* pb = reinterpret_cast<const __half2*>(&b);
  __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

// thread = 8 channels of one output pixel
__global__ void maxpool2x2_kernel(const __half* __restrict__ x, __half* __restrict__ y, int batch, int height, int width, int channels,
                                  int x_ld) {
  const int c8 = channels >> 3;
  const int oh = height >> 1, ow = width >> 1;
  const long long total = static_cast<long long>(batch) * oh * ow * c8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % c8);
  long long t = idx / c8;
  const int px = static_cast<int>(t % ow); t /= ow;
  const int py = static_cast<int>(t % oh);
  const int img = static_cast<int>(t / oh);
  const __half* p00 = x + ((static_cast<long long>(img) * height + 2 * py) * width + 2 * px) * x_ld + cg * 8;
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p00));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(p00 + x_ld));
  const uint4 c = __ldg(reinterpret_cast<const uint4*>(p00 + static_cast<long long>(width) * x_ld));
  const uint4 d = __ldg(reinterpret_cast<const uint4*>(p00 + static_cast<long long>(width) * x_ld + x_ld));
  reinterpret_cast<uint4*>(y)[idx] = hmax8(hmax8(a, b), hmax8(c, d));
}

int maxpool2x2(const void* x, void* y, int batch, int height, int width, int channels, int x_ld, cudaStream_t stream) {
  YB_REQUIRE(x && y && batch > 0 && height % 2 == 0 && width % 2 == 0 && channels % 8 == 0 && x_ld % 8 == 0 && x_ld >= channels,
             "maxpool2x2: bad argument");
  const long long total = static_cast<long long>(batch) * (height / 2) * (width / 2) * (channels / 8);
  maxpool2x2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), batch, height, width, channels, x_ld);
  return check_launch("maxpool2x2_kernel");
}

// ------------------------------------------------------------------------------------------
// reorg on fp16 NHWC: out[b, h', w', y_ch_off + (sh*2+sw)*C + c] = in[b, 2h'+sh, 2w'+sw, c]
__global__ void reorg_nhwc_kernel(const __half* __restrict__ x, __half* __restrict__ y, int batch, int height, int width, int channels,
                                  int x_ld, int y_ld, int y_ch_off) {
  const int c8 = channels >> 3;
  const int oh = height >> 1, ow = width >> 1;
  const long long total = static_cast<long long>(batch) * oh * ow * 4 * c8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % c8);
  long long t = idx / c8;
  const int off = static_cast<int>(t % 4); t /= 4;   // sh*2+sw
  const int px = static_cast<int>(t % ow); t /= ow;
  const int py = static_cast<int>(t % oh);
  const int img = static_cast<int>(t / oh);
  const int sh = off >> 1, sw = off & 1;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(
      x + ((static_cast<long long>(img) * height + 2 * py + sh) * width + 2 * px + sw) * x_ld + cg * 8));
  *reinterpret_cast<uint4*>(y + ((static_cast<long long>(img) * oh + py) * ow + px) * y_ld + y_ch_off + off * channels + cg * 8) = v;
}

int reorg_nhwc(const void* x, void* y, int batch, int height, int width, int channels, int x_ld, int y_ld, int y_ch_off,
               cudaStream_t stream) {
  YB_REQUIRE(x && y && batch > 0 && height % 2 == 0 && width % 2 == 0 && channels % 8 == 0 && x_ld % 8 == 0 && y_ld % 8 == 0 &&
                 y_ch_off % 8 == 0 && y_ld >= y_ch_off + 4 * channels,
             "reorg_nhwc: bad argument");
  const long long total = static_cast<long long>(batch) * (height / 2) * (width / 2) * 4 * (channels / 8);
  reorg_nhwc_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), batch, height, width, channels, x_ld, y_ld, y_ch_off);
  return check_launch("reorg_nhwc_kernel");
}

// reorg on the caller's fp32 NCHW tensors (the public model.yolo2.reorg):
// out[b, (sh*stride_w+sw)*C + c, h', w'] = x[b, c, h'*stride_h+sh, w'*stride_w+sw]
__global__ void reorg_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int batch, int channels, int height, int width,
                                  int stride_h, int stride_w) {
  const int oh = height / stride_h, ow = width / stride_w;
  const long long total = static_cast<long long>(batch) * channels * stride_h * stride_w * oh * ow;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int px = static_cast<int>(idx % ow);
  long long t = idx / ow;
  const int py = static_cast<int>(t % oh); t /= oh;
  const int oc = static_cast<int>(t % (channels * stride_h * stride_w));
  const int img = static_cast<int>(t / (channels * stride_h * stride_w));
  const int c = oc % channels;
  const int off = oc / channels;
  const int sh = off / stride_w, sw = off % stride_w;
  y[idx] = __ldg(x + ((static_cast<long long>(img) * channels + c) * height + py * stride_h + sh) * width + px * stride_w + sw);
}

int reorg_nchw(const float* x, float* y, int batch, int channels, int height, int width, int stride_h, int stride_w, cudaStream_t stream) {
  YB_REQUIRE(x && y && batch > 0 && channels > 0 && stride_h > 0 && stride_w > 0 && height % stride_h == 0 && width % stride_w == 0,
             "reorg_nchw: bad argument");
  const long long total = static_cast<long long>(batch) * channels * height * width;
  reorg_nchw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(x, y, batch, channels, height, width, stride_h, stride_w);
  return check_launch("reorg_nchw_kernel");
}

}  // namespace yb
