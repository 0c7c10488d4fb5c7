// HBM-bound layout / pointwise kernels of the Darknet-19 path (coalesced, 16-byte vectorised):
//   pack_weight      fp32 OIHW -> fp16 KRSC (the K-major B operand of the implicit GEMM)
//   bn_fold          eval-mode BatchNorm2d -> per-channel fp32 (scale, shift)      model/yolo2.py:58
//   maxpool2x2       nn.MaxPool2d(2) on fp16 NHWC                                 model/yolo2.py:79,86,97
//   reorg            space-to-depth, offset-major channel order                   model/yolo2.py:33-46
#include "yb_common.h"
#include <cuda_fp16.h>
#include <stdint.h>

namespace yb {

// ------------------------------------------------------------------------------------------
// mode 0: out[co][r][s][ci] = w[co][ci][r][s].  One block per (Cout row, 128-channel chunk): the [128][k*k] input run is read
// coalesced into shared memory and written back as k*k coalesced fp16 rows.
__global__ void __launch_bounds__(128) pack_weight_fwd_kernel(const float* __restrict__ w, __half* __restrict__ out, int cin, int k) {
  __shared__ float tile[128 * 9 + 8];
  const int k2 = k * k;
  const int co = blockIdx.y, ci0 = blockIdx.x * 128, t = threadIdx.x;
  const int nci = cin - ci0 < 128 ? cin - ci0 : 128;
  const float* src = w + (static_cast<long long>(co) * cin + ci0) * k2;
  for (int j = t; j < nci * k2; j += 128) tile[j] = src[j];
  __syncthreads();
  if (t < nci)
    for (int tap = 0; tap < k2; ++tap) out[(static_cast<long long>(co) * k2 + tap) * cin + ci0 + t] = __float2half_rn(tile[t * k2 + tap]);
}

// mode 1 (data-gradient operand): out[ci][r][s][co] = w[co][ci][k-1-r][k-1-s]  (rotated, in/out swapped; the reduction
// dimension Cout may be zero-padded to cout_pad).  32 x 32 tiles of the [Cout] x [Cin*k*k] matrix through shared memory:
// rows of w are read along (ci, tap), rows of out are written along co.
__global__ void __launch_bounds__(256) pack_weight_dgrad_kernel(const float* __restrict__ w, __half* __restrict__ out, int cout, int cin, int k, int cout_pad) {
  __shared__ float tile[32][33];
  const int k2 = k * k;
  const long long cols = static_cast<long long>(cin) * k2;
  const long long j0 = static_cast<long long>(blockIdx.x) * 32;
  const int co0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int co = co0 + r;
    const long long j = j0 + tx;
    tile[r][tx] = (co < cout && j < cols) ? w[static_cast<long long>(co) * cols + j] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const long long j = j0 + r;
    const int co = co0 + tx;
    if (j < cols && co < cout_pad) {
      const long long ci = j / k2;
      const int tap = static_cast<int>(j - ci * k2);
      out[(ci * k2 + (k2 - 1 - tap)) * cout_pad + co] = __float2half_rn(tile[tx][r]);
    }
  }
}

int pack_weight(const float* w, void* out, int cout, int cin, int k, int mode, int cout_pad, cudaStream_t stream) {
  YB_REQUIRE(w && out && cout > 0 && cin > 0 && (k == 1 || k == 3) && (mode == 0 || mode == 1), "pack_weight: bad argument");
  if (cout_pad < cout) cout_pad = cout;
  if (mode == 0) {
    pack_weight_fwd_kernel<<<dim3((cin + 127) / 128, cout), 128, 0, stream>>>(w, reinterpret_cast<__half*>(out), cin, k);
  } else {
    const long long cols = static_cast<long long>(cin) * k * k;
    pack_weight_dgrad_kernel<<<dim3(static_cast<unsigned>((cols + 31) / 32), (cout_pad + 31) / 32), 256, 0, stream>>>(
        w, reinterpret_cast<__half*>(out), cout, cin, k, cout_pad);
  }
  return check_launch("pack_weight_kernel");
}

// Both training operands of MANY units in one launch (yb_pack_weights_batch): every block converts one 64-filter x (32 channels x 9 taps | 256
// channels) tile of one unit -- found from the unit table's running block offsets -- reads the fp32 OIHW rows coalesced ONCE, and writes the
// forward operand out_f[co][tap][ci] and the data-gradient operand out_d[ci][k*k-1-tap][co] (rotated, filters zero-padded to cout_pad) as
// 4-byte pairs along their contiguous dimension.  Replaces 2 launches per unit (44 per Darknet-19 step, most of them latency-bound).
struct PackUnit {
  const float* w;
  __half* out_f;      // may be null
  __half* out_d;      // may be null
  int cout, cin, k, cout_pad, block0, ci_blocks;
};
constexpr int kPackCo = 64, kPackPitch = 290;      // halves per staged filter row (288 + 2: odd word pitch)

__global__ void __launch_bounds__(256) pack_weights_batch_kernel(const PackUnit* __restrict__ units, int num_units) {
  __shared__ __half tile[kPackCo * kPackPitch];
  __shared__ int s_unit;
  if (threadIdx.x == 0) {
    int u = 0;
    while (u + 1 < num_units && static_cast<int>(blockIdx.x) >= units[u + 1].block0) ++u;
    s_unit = u;
  }
  __syncthreads();
  const PackUnit pu = units[s_unit];
  const int k2 = pu.k * pu.k;
  const int chunk = pu.k == 3 ? 32 : 256;                         // input channels per tile (<= 288 staged columns)
  const int local = static_cast<int>(blockIdx.x) - pu.block0;
  const int co0 = (local / pu.ci_blocks) * kPackCo, ci0 = (local % pu.ci_blocks) * chunk;
  const int nci = pu.cin - ci0 < chunk ? pu.cin - ci0 : chunk;     // even (cin % 2 == 0 is required)
  const int ncols = nci * k2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int r = warp; r < kPackCo; r += 8) {
    const int co = co0 + r;
    const float* src = pu.w + (static_cast<long long>(co) * pu.cin + ci0) * k2;
    for (int c = lane; c < ncols; c += 32) tile[r * kPackPitch + c] = __float2half_rn(co < pu.cout ? __ldg(src + c) : 0.f);
  }
  __syncthreads();
  if (pu.out_f != nullptr) {
    const int half_ci = nci >> 1;
    const int items = kPackCo * k2 * half_ci;
    for (int i = tid; i < items; i += 256) {
      const int cp = i % half_ci, rt = i / half_ci;
      const int tap = rt % k2, r = rt / k2;
      const int co = co0 + r;
      if (co >= pu.cout) break;                                     // rows ascend with i
      const __half2 v = __halves2half2(tile[r * kPackPitch + (2 * cp) * k2 + tap], tile[r * kPackPitch + (2 * cp + 1) * k2 + tap]);
      *reinterpret_cast<__half2*>(pu.out_f + (static_cast<long long>(co) * k2 + tap) * pu.cin + ci0 + 2 * cp) = v;
    }
  }
  if (pu.out_d != nullptr) {
    const int items = ncols * (kPackCo / 2);
    for (int i = tid; i < items; i += 256) {
      const int cp = i & (kPackCo / 2 - 1), col = i / (kPackCo / 2);
      const int co = co0 + 2 * cp;
      if (co >= pu.cout_pad) continue;
      const int ci = col / k2, tap = col - ci * k2;
      const __half2 v = __halves2half2(tile[(2 * cp) * kPackPitch + col], tile[(2 * cp + 1) * kPackPitch + col]);
      *reinterpret_cast<__half2*>(pu.out_d + (static_cast<long long>(ci0 + ci) * k2 + (k2 - 1 - tap)) * pu.cout_pad + co) = v;
    }
  }
}

int pack_weights_batch(const void* units_dev, int num_units, int total_blocks, cudaStream_t stream) {
  YB_REQUIRE(units_dev && num_units > 0 && total_blocks > 0, "pack_weights_batch: bad argument");
  static_assert(sizeof(PackUnit) == 48, "PackUnit must match yb_pack_unit");
  pack_weights_batch_kernel<<<total_blocks, 256, 0, stream>>>(static_cast<const PackUnit*>(units_dev), num_units);
  return check_launch("pack_weights_batch_kernel");
}

// Split-precision B operand (strict mode, see ConvParams::a_wrap in conv_igemm.cu): out[co][tap][s*Cin + ci] for s in [0, segments),
// segment s holding w_hi = fp16(w) or, where bit s of lo_mask is set, w_lo = fp16(w - fp32(w_hi)).  Same tiling as mode 0.
__global__ void __launch_bounds__(128) pack_weight_split_kernel(const float* __restrict__ w, __half* __restrict__ out, int cin, int k, int segments,
                                                                int lo_mask) {
  __shared__ float tile[128 * 9 + 8];
  const int k2 = k * k;
  const int co = blockIdx.y, ci0 = blockIdx.x * 128, t = threadIdx.x;
  const int nci = cin - ci0 < 128 ? cin - ci0 : 128;
  const float* src = w + (static_cast<long long>(co) * cin + ci0) * k2;
  for (int j = t; j < nci * k2; j += 128) tile[j] = src[j];
  __syncthreads();
  if (t < nci) {
    const long long kw = static_cast<long long>(segments) * cin;
    for (int tap = 0; tap < k2; ++tap) {
      const float v = tile[t * k2 + tap];
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      for (int s = 0; s < segments; ++s)
        out[(static_cast<long long>(co) * k2 + tap) * kw + static_cast<long long>(s) * cin + ci0 + t] = ((lo_mask >> s) & 1) ? lo : hi;
    }
  }
}

int pack_weight_split(const float* w, void* out, int cout, int cin, int k, int segments, int lo_mask, cudaStream_t stream) {
  YB_REQUIRE(w && out && cout > 0 && cin > 0 && (k == 1 || k == 3) && segments >= 1 && segments <= 3, "pack_weight_split: bad argument");
  pack_weight_split_kernel<<<dim3((cin + 127) / 128, cout), 128, 0, stream>>>(w, reinterpret_cast<__half*>(out), cin, k, segments, lo_mask);
  return check_launch("pack_weight_split_kernel");
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
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: the conv that follows may start its prologue now
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

// nn.MaxPool2d(2) on split-precision activations (hi at channel c, lo at c + lo_off of the same pixel): the window element with the
// largest hi + lo (fp32) wins and its (hi, lo) pair is copied, so the pooled value is exactly the maximum of the represented values.
__global__ void maxpool2x2_split_kernel(const __half* __restrict__ x, __half* __restrict__ y, int batch, int height, int width, int channels,
                                        int x_ld, int x_lo_off, int y_ld, int y_lo_off) {
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
  const long long offs[4] = {0, x_ld, static_cast<long long>(width) * x_ld, static_cast<long long>(width) * x_ld + x_ld};
  uint4 hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    hi[j] = __ldg(reinterpret_cast<const uint4*>(p00 + offs[j]));
    lo[j] = __ldg(reinterpret_cast<const uint4*>(p00 + offs[j] + x_lo_off));
  }
  uint4 oh4, ol4;
  __half* ph = reinterpret_cast<__half*>(&oh4);
  __half* pl = reinterpret_cast<__half*>(&ol4);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float best = -INFINITY;
    __half bh = __float2half_rn(0.f), bl = bh;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half h = reinterpret_cast<const __half*>(&hi[j])[e], l = reinterpret_cast<const __half*>(&lo[j])[e];
      const float v = __half2float(h) + __half2float(l);
      if (v > best) { best = v; bh = h; bl = l; }
    }
    ph[e] = bh; pl[e] = bl;
  }
  __half* dst = y + ((static_cast<long long>(img) * oh + py) * ow + px) * y_ld + cg * 8;
  *reinterpret_cast<uint4*>(dst) = oh4;
  *reinterpret_cast<uint4*>(dst + y_lo_off) = ol4;
}

int maxpool2x2_split(const void* x, void* y, int batch, int height, int width, int channels, int x_ld, int x_lo_off, int y_ld, int y_lo_off,
                     cudaStream_t stream) {
  YB_REQUIRE(x && y && batch > 0 && height % 2 == 0 && width % 2 == 0 && channels % 8 == 0 && x_ld % 8 == 0 && y_ld % 8 == 0 && x_lo_off % 8 == 0 &&
                 y_lo_off % 8 == 0 && x_lo_off >= channels && y_lo_off >= channels && x_ld >= x_lo_off + channels && y_ld >= y_lo_off + channels,
             "maxpool2x2_split: bad argument");
  const long long total = static_cast<long long>(batch) * (height / 2) * (width / 2) * (channels / 8);
  maxpool2x2_split_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), batch, height, width, channels, x_ld, x_lo_off, y_ld, y_lo_off);
  return check_launch("maxpool2x2_split_kernel");
}

// ------------------------------------------------------------------------------------------
// model.yolo2.Tiny's ConstantPad2d((0, 1, 0, 1), float32 min) + MaxPool2d(2, stride=1)  (model/yolo2.py:150-151): same-size output,
// out[y, x] = max over the in-range pixels of {y, y+1} x {x, x+1} (the pad value never wins).  thread = 8 channels of one pixel
__global__ void maxpool2x2_s1_kernel(const __half* __restrict__ x, __half* __restrict__ y, int batch, int height, int width, int channels,
                                     int x_ld) {
  const int c8 = channels >> 3;
  const long long total = static_cast<long long>(batch) * height * width * c8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % c8);
  long long t = idx / c8;
  const int px = static_cast<int>(t % width); t /= width;
  const int py = static_cast<int>(t % height);
  const __half* p00 = x + (t * width + px) * x_ld + cg * 8;       // t = img * height + py
  uint4 m = __ldg(reinterpret_cast<const uint4*>(p00));
  const bool right = px + 1 < width, down = py + 1 < height;
  if (right) m = hmax8(m, __ldg(reinterpret_cast<const uint4*>(p00 + x_ld)));
  if (down) m = hmax8(m, __ldg(reinterpret_cast<const uint4*>(p00 + static_cast<long long>(width) * x_ld)));
  if (right && down) m = hmax8(m, __ldg(reinterpret_cast<const uint4*>(p00 + static_cast<long long>(width) * x_ld + x_ld)));
  reinterpret_cast<uint4*>(y)[idx] = m;
}

int maxpool2x2_s1(const void* x, void* y, int batch, int height, int width, int channels, int x_ld, cudaStream_t stream) {
  YB_REQUIRE(x && y && batch > 0 && height > 0 && width > 0 && channels % 8 == 0 && x_ld % 8 == 0 && x_ld >= channels, "maxpool2x2_s1: bad argument");
  const long long total = static_cast<long long>(batch) * height * width * (channels / 8);
  maxpool2x2_s1_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), batch, height, width, channels, x_ld);
  return check_launch("maxpool2x2_s1_kernel");
}

// Backward of the same pad + stride-1 pooling (training of model.yolo2.Tiny): dx[p] = sum of dy over the (up to four) windows that contain
// p and whose FIRST maximum (scan order (0,0), (0,1), (1,0), (1,1), as torch's max_pool2d records it) is p.  thread = 8 channels of one pixel.
__global__ void maxpool2x2_s1_bwd_kernel(const __half* __restrict__ x, const __half* __restrict__ dy, __half* __restrict__ dx, int batch, int height,
                                         int width, int channels) {
  const int c8 = channels >> 3;
  const long long total = static_cast<long long>(batch) * height * width * c8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % c8);
  long long t = idx / c8;
  const int px = static_cast<int>(t % width); t /= width;
  const int py = static_cast<int>(t % height);
  const long long img = t / height;
  const __half* xb = x + img * height * width * channels + cg * 8;
  const __half* gb = dy + img * height * width * channels + cg * 8;
  // the 3 x 3 neighbourhood of p (values outside the frame = -inf, they never win)
  float v[3][3][8];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int yy = py - 1 + r, xx = px - 1 + c;
      if (yy >= 0 && yy < height && xx >= 0 && xx < width) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(xb + (static_cast<long long>(yy) * width + xx) * channels));
        const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h[e]); v[r][c][2 * e] = f.x; v[r][c][2 * e + 1] = f.y; }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[r][c][e] = -INFINITY;
      }
    }
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  // window with top-left (py - dy, px - dx): p sits at position (dy, dx) of it
#pragma unroll
  for (int dyy = 0; dyy < 2; ++dyy)
#pragma unroll
    for (int dxx = 0; dxx < 2; ++dxx) {
      const int wy = py - dyy, wx = px - dxx;
      if (wy < 0 || wx < 0) continue;
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(gb + (static_cast<long long>(wy) * width + wx) * channels));
      const __half2* h = reinterpret_cast<const __half2*>(&q);
      float g[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h[e]); g[2 * e] = f.x; g[2 * e + 1] = f.y; }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        // neighbourhood coordinates of the window's four elements: rows (1 - dyy) + {0, 1}, columns (1 - dxx) + {0, 1}
        float best = -INFINITY;
        int arg = -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float val = v[1 - dyy + (k >> 1)][1 - dxx + (k & 1)][e];
          if (arg < 0 || val > best) { best = val; arg = k; }      // first maximum wins
        }
        if (arg == dyy * 2 + dxx) acc[e] += g[e];
      }
    }
  uint4 out;
  __half2* ho = reinterpret_cast<__half2*>(&out);
#pragma unroll
  for (int e = 0; e < 4; ++e) ho[e] = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
  reinterpret_cast<uint4*>(dx)[idx] = out;
}

int maxpool2x2_s1_bwd(const void* x, const void* dy, void* dx, int batch, int height, int width, int channels, cudaStream_t stream) {
  YB_REQUIRE(x && dy && dx && batch > 0 && height > 0 && width > 0 && channels % 8 == 0, "maxpool2x2_s1_bwd: bad argument");
  const long long total = static_cast<long long>(batch) * height * width * (channels / 8);
  maxpool2x2_s1_bwd_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, stream>>>(
      reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(dy), reinterpret_cast<__half*>(dx), batch, height, width, channels);
  return check_launch("maxpool2x2_s1_bwd_kernel");
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
