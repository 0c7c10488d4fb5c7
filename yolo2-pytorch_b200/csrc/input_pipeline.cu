// GPU input pipeline: the step immediately upstream of the first conv kernel.
// Replaces, for a whole batch of decoded frames in ONE launch, what the reference does per image on the CPU:
//   transform/resize/image.py:23-24 / label.py:25-31  `rescale`: cv2.resize(image, (width, height)) (INTER_LINEAR, uint8) and the
//                                                      box scaling yx *= (height / _height, width / _width)
//   transform/image.py:27-29                          BGR2RGB
// (ToTensor's 1/255 and the HWC -> network layout change are already fused into yb_conv0_u8_bn_leaky_pool_fwd.)
//
// Exactness contract: bit-identical to cv2.resize (OpenCV 4.x, 8-bit, INTER_LINEAR) -- the arithmetic lives in the
// reference's third-party dependency, so it is restated here from its published algorithm (modules/imgproc resize.cpp,
// the 8-bit fixed-point path) and pinned by fixtures generated with cv2 itself:
//   fx = float((dx + 0.5) * (src_w / dst_w) - 0.5) (double arithmetic, one rounding to float); sx = floor(fx); fx -= sx;
//   horizontally sx < 0 -> (sx, fx) = (0, 0), sx >= src_w - 1 -> (src_w - 1, 0); vertically the two rows are clamped
//   to the image but the fraction is kept; coefficients are rounded to 11-bit fixed point with round-half-even
//   (saturate_cast<short>(c * 2048)); the row pass is S[sx] * a0 + S[sx + 1] * a1 in int, the column pass
//   (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
#include "yb_common.h"
#include <stdint.h>

namespace yb {

struct ResizeCoef { int s0, s1, a0, a1; };

__device__ __forceinline__ ResizeCoef resize_coef(int d, int n_dst, int n_src, bool clamp_fraction) {
  const double scale = static_cast<double>(n_src) / static_cast<double>(n_dst);
  float f = static_cast<float>((static_cast<double>(d) + 0.5) * scale - 0.5);
  int s = static_cast<int>(floorf(f));
  f -= static_cast<float>(s);
  ResizeCoef c;
  if (clamp_fraction) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= n_src - 1) { f = 0.f; s = n_src - 1; }
    c.s0 = s;
    c.s1 = s + 1 < n_src ? s + 1 : n_src - 1;
  } else {
    c.s0 = s < 0 ? 0 : (s < n_src ? s : n_src - 1);
    c.s1 = s + 1 < 0 ? 0 : (s + 1 < n_src ? s + 1 : n_src - 1);
  }
  c.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  c.a1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return c;
}

// one thread per output pixel (3 channels); blockIdx.y = image
__global__ void __launch_bounds__(256) resize_u8_kernel(const uint8_t* __restrict__ src, const long long* __restrict__ src_off,
                                                        const int* __restrict__ src_hw, uint8_t* __restrict__ dst, int height, int width,
                                                        int swap_rb) {
  const int img = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= height * width) return;
  const int dy = idx / width, dx = idx - dy * width;
  const int sh = src_hw[2 * img], sw = src_hw[2 * img + 1];
  const uint8_t* s = src + src_off[img];
  const ResizeCoef cx = resize_coef(dx, width, sw, true);
  const ResizeCoef cy = resize_coef(dy, height, sh, false);
  const uint8_t* r0 = s + static_cast<long long>(cy.s0) * sw * 3;
  const uint8_t* r1 = s + static_cast<long long>(cy.s1) * sw * 3;
  uint8_t out[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int h0 = static_cast<int>(r0[cx.s0 * 3 + c]) * cx.a0 + static_cast<int>(r0[cx.s1 * 3 + c]) * cx.a1;
    const int h1 = static_cast<int>(r1[cx.s0 * 3 + c]) * cx.a0 + static_cast<int>(r1[cx.s1 * 3 + c]) * cx.a1;
    int v = (((cy.a0 * (h0 >> 4)) >> 16) + ((cy.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    out[c] = static_cast<uint8_t>(v);
  }
  uint8_t* d = dst + (static_cast<long long>(img) * height * width + idx) * 3;
  d[0] = swap_rb ? out[2] : out[0];
  d[1] = out[1];
  d[2] = swap_rb ? out[0] : out[2];
}

// boxes [B, G, 2] (y, x): yx *= (height / src_h, width / src_w) in float32 exactly as numpy does it (label.py:27-30)
__global__ void rescale_boxes_kernel(float* __restrict__ yx_min, float* __restrict__ yx_max, const int* __restrict__ src_hw, int batch, int slots,
                                     int height, int width) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * slots) return;
  const int img = i / slots;
  // np.array([height / _height, width / _width], np.float32): a double division rounded once more to float32
  const float sy = static_cast<float>(static_cast<double>(height) / static_cast<double>(src_hw[2 * img]));
  const float sx = static_cast<float>(static_cast<double>(width) / static_cast<double>(src_hw[2 * img + 1]));
  yx_min[2 * i] = __fmul_rn(yx_min[2 * i], sy); yx_min[2 * i + 1] = __fmul_rn(yx_min[2 * i + 1], sx);
  yx_max[2 * i] = __fmul_rn(yx_max[2 * i], sy); yx_max[2 * i + 1] = __fmul_rn(yx_max[2 * i + 1], sx);
}

// torchvision.transforms.ToTensor on a batch (transform/image.py + the `transform_tensor` step of utils/data.py:120-121):
// uint8 NHWC [B,H,W,3] -> fp32 NCHW [B,3,H,W], value / 255 (IEEE division, as `.div(255)` does).  The inference path never
// needs this (the first conv kernel reads the uint8 frames directly); the training forward and its first-layer weight
// gradient read the fp32 NCHW image the reference hands them.
__global__ void totensor_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int batch, int hw) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;     // one thread per pixel
  if (idx >= static_cast<long long>(batch) * hw) return;
  const long long img = idx / hw, pix = idx - img * hw;
  const uint8_t* s = src + idx * 3;
  float* d = dst + img * 3 * hw + pix;
#pragma unroll
  for (int c = 0; c < 3; ++c) d[static_cast<long long>(c) * hw] = __fdiv_rn(static_cast<float>(s[c]), 255.f);
}

// Augmented form (the default training resize `resize_train = transform.resize.label.RandomCrop`, config.ini:48, after the optional
// `transform.augmentation.RandomFlipHorizontally`, config.ini:47): out = cv2.resize(crop(flip(image))).  Both are pure index transforms on
// the SOURCE of the same bit-exact resize: a flipped image's column j is column (w - 1 - j) (cv2.flip(image, 1), augmentation.py:87-95),
// the crop image[y0:y1, x0:x1] (resize/label.py:73) an offset window.  crop: int[B][4] = (y0, x0, y1, x1) in the (flipped) frame or NULL;
// flip: uint8[B] or NULL.
__global__ void __launch_bounds__(256) resize_aug_u8_kernel(const uint8_t* __restrict__ src, const long long* __restrict__ src_off,
                                                            const int* __restrict__ src_hw, const int* __restrict__ crop,
                                                            const uint8_t* __restrict__ flip, uint8_t* __restrict__ dst, int height, int width,
                                                            int swap_rb) {
  const int img = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= height * width) return;
  const int dy = idx / width, dx = idx - dy * width;
  const int sw = src_hw[2 * img + 1];
  int y0 = 0, x0 = 0, eh = src_hw[2 * img], ew = sw;
  if (crop != nullptr) { y0 = crop[4 * img]; x0 = crop[4 * img + 1]; eh = crop[4 * img + 2] - y0; ew = crop[4 * img + 3] - x0; }
  const bool fl = flip != nullptr && flip[img] != 0;
  const uint8_t* s = src + src_off[img];
  const ResizeCoef cx = resize_coef(dx, width, ew, true);
  const ResizeCoef cy = resize_coef(dy, height, eh, false);
  const uint8_t* r0 = s + static_cast<long long>(y0 + cy.s0) * sw * 3;
  const uint8_t* r1 = s + static_cast<long long>(y0 + cy.s1) * sw * 3;
  const int c0 = fl ? sw - 1 - (x0 + cx.s0) : x0 + cx.s0;
  const int c1 = fl ? sw - 1 - (x0 + cx.s1) : x0 + cx.s1;
  uint8_t out[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int h0 = static_cast<int>(r0[c0 * 3 + c]) * cx.a0 + static_cast<int>(r0[c1 * 3 + c]) * cx.a1;
    const int h1 = static_cast<int>(r1[c0 * 3 + c]) * cx.a0 + static_cast<int>(r1[c1 * 3 + c]) * cx.a1;
    int v = (((cy.a0 * (h0 >> 4)) >> 16) + ((cy.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    out[c] = static_cast<uint8_t>(v);
  }
  uint8_t* d = dst + (static_cast<long long>(img) * height * width + idx) * 3;
  d[0] = swap_rb ? out[2] : out[0];
  d[1] = out[1];
  d[2] = swap_rb ? out[0] : out[2];
}

// boxes, in the reference's order and float32 arithmetic: flip (x' = w - x, min / max swapped; augmentation.py:91-94), crop
// (yx -= margin, the UN-truncated float32 margin; resize/label.py:74), scale by (height / crop_h, width / crop_w) (resize/label.py:27-30)
__global__ void augment_boxes_kernel(float* __restrict__ yx_min, float* __restrict__ yx_max, const int* __restrict__ src_hw, const int* __restrict__ crop,
                                     const float* __restrict__ margin, const uint8_t* __restrict__ flip, int batch, int slots, int height, int width) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * slots) return;
  const int img = i / slots;
  float ymin = yx_min[2 * i], xmin = yx_min[2 * i + 1], ymax = yx_max[2 * i], xmax = yx_max[2 * i + 1];
  const int sw = src_hw[2 * img + 1];
  int eh = src_hw[2 * img], ew = sw;
  if (flip != nullptr && flip[img] != 0) {
    const float w = static_cast<float>(sw);
    const float t = __fsub_rn(w, xmin);
    xmin = __fsub_rn(w, xmax);
    xmax = t;
  }
  if (crop != nullptr) {
    eh = crop[4 * img + 2] - crop[4 * img]; ew = crop[4 * img + 3] - crop[4 * img + 1];
    const float my = margin[2 * img], mx = margin[2 * img + 1];
    ymin = __fsub_rn(ymin, my); xmin = __fsub_rn(xmin, mx); ymax = __fsub_rn(ymax, my); xmax = __fsub_rn(xmax, mx);
  }
  const float sy = static_cast<float>(static_cast<double>(height) / static_cast<double>(eh));
  const float sx = static_cast<float>(static_cast<double>(width) / static_cast<double>(ew));
  yx_min[2 * i] = __fmul_rn(ymin, sy); yx_min[2 * i + 1] = __fmul_rn(xmin, sx);
  yx_max[2 * i] = __fmul_rn(ymax, sy); yx_max[2 * i + 1] = __fmul_rn(xmax, sx);
}

int resize_aug_batch_u8(const void* src, const long long* src_off, const int* src_hw, const int* crop, const float* margin, const unsigned char* flip,
                        void* dst, int batch, int height, int width, int swap_rb, float* yx_min, float* yx_max, int slots, cudaStream_t stream) {
  YB_REQUIRE(src && src_off && src_hw && dst && batch > 0 && height > 0 && width > 0, "resize_aug_batch_u8: bad argument");
  YB_REQUIRE((yx_min == nullptr) == (yx_max == nullptr) && slots >= 0, "resize_aug_batch_u8: boxes come as a (yx_min, yx_max) pair");
  YB_REQUIRE(crop == nullptr || margin != nullptr || yx_min == nullptr, "resize_aug_batch_u8: cropping boxes needs the float margins");
  const int pixels = height * width;
  resize_aug_u8_kernel<<<dim3((pixels + 255) / 256, batch), 256, 0, stream>>>(static_cast<const uint8_t*>(src), src_off, src_hw, crop, flip,
                                                                              static_cast<uint8_t*>(dst), height, width, swap_rb);
  int rc = check_launch("resize_aug_u8_kernel");
  if (rc) return rc;
  if (yx_min != nullptr && slots > 0) {
    augment_boxes_kernel<<<(batch * slots + 127) / 128, 128, 0, stream>>>(yx_min, yx_max, src_hw, crop, margin, flip, batch, slots, height, width);
    rc = check_launch("augment_boxes_kernel");
  }
  return rc;
}

// ------------------------------------------------------------------------------------------------
// cv2.warpAffine(image, M, (dw, dh), flags=INTER_LINEAR, borderMode=BORDER_CONSTANT, borderValue=fill) for uint8 HWC frames, bit-exact:
// the arithmetic of the reference's `Rotator.__call__` (transform/augmentation.py:46-49, used by `random_rotate` :61-76) and of
// `transform.resize.image.fixed` (transform/resize/image.py:36-46; INTER_AREA is INTER_LINEAR inside warpAffine).  It lives in the
// reference's third-party dependency, so it is restated from OpenCV's published algorithm (imgproc imgwarp.cpp: warpAffine + remapBilinear,
// 8-bit fixed-point path) and pinned by fixtures made with cv2 itself:
//   the caller passes the INVERTED matrix (double); X0 = cvRound((m01*y + m02)*1024) + 16, adelta = cvRound(m00*x*1024) (likewise Y);
//   X = (X0 + adelta) >> 5; source pixel sx = X >> 5 with the 5-bit fraction X & 31; bilinear weights (32-fy)(32-fx)*32 ... as int16
//   (saturated to 32767, the remainder added to the last tap so they sum to 32768); out = (sum + 2^14) >> 15; taps outside the
//   frame read `fill`.
__global__ void __launch_bounds__(256) warp_affine_u8_kernel(const uint8_t* __restrict__ src, int sh, int sw, uint8_t* __restrict__ dst, int dh, int dw,
                                                             double m00, double m01, double m02, double m10, double m11, double m12, int fill0,
                                                             int fill1, int fill2) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= dh * dw) return;
  const int y = idx / dw, x = idx - y * dw;
  auto sat = [](double v) -> long long {
    const long long r = __double2ll_rn(v);
    return r < -2147483648ll ? -2147483648ll : (r > 2147483647ll ? 2147483647ll : r);
  };
  const long long X0 = sat((m01 * static_cast<double>(y) + m02) * 1024.0) + 16;
  const long long Y0 = sat((m11 * static_cast<double>(y) + m12) * 1024.0) + 16;
  const long long X = (X0 + sat(m00 * static_cast<double>(x) * 1024.0)) >> 5;
  const long long Y = (Y0 + sat(m10 * static_cast<double>(x) * 1024.0)) >> 5;
  long long sxl = X >> 5, syl = Y >> 5;
  sxl = sxl < -32768 ? -32768 : (sxl > 32767 ? 32767 : sxl);
  syl = syl < -32768 ? -32768 : (syl > 32767 ? 32767 : syl);
  const int sx = static_cast<int>(sxl), sy = static_cast<int>(syl);
  const int fx = static_cast<int>(X & 31), fy = static_cast<int>(Y & 31);
  int w[4] = {(32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32};
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = w[k] > 32767 ? 32767 : w[k];
  w[3] += 32768 - (w[0] + w[1] + w[2] + w[3]);
  const int fill[3] = {fill0, fill1, fill2};
  int acc[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int yy = sy + (k >> 1), xx = sx + (k & 1);
    const bool inside = yy >= 0 && yy < sh && xx >= 0 && xx < sw;
    const uint8_t* sp = src + (static_cast<long long>(inside ? yy : 0) * sw + (inside ? xx : 0)) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += (inside ? static_cast<int>(sp[c]) : fill[c]) * w[k];
  }
  uint8_t* d = dst + static_cast<long long>(idx) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int v = (acc[c] + (1 << 14)) >> 15;
    d[c] = static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
  }
}

int warp_affine_u8(const void* src, int sh, int sw, void* dst, int dh, int dw, const double* minv, const int* fill, cudaStream_t stream) {
  YB_REQUIRE(src && dst && minv && fill && sh > 0 && sw > 0 && dh > 0 && dw > 0 && static_cast<long long>(dh) * dw < (1ll << 31), "warp_affine_u8: bad argument");
  const int pixels = dh * dw;
  warp_affine_u8_kernel<<<(pixels + 255) / 256, 256, 0, stream>>>(static_cast<const uint8_t*>(src), sh, sw, static_cast<uint8_t*>(dst), dh, dw, minv[0], minv[1],
                                                                 minv[2], minv[3], minv[4], minv[5], fill[0], fill[1], fill[2]);
  return check_launch("warp_affine_u8_kernel");
}

int totensor_u8(const void* src, float* dst, int batch, int height, int width, cudaStream_t stream) {
  YB_REQUIRE(src && dst && batch > 0 && height > 0 && width > 0, "totensor_u8: bad argument");
  const long long total = static_cast<long long>(batch) * height * width;
  totensor_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(static_cast<const uint8_t*>(src), dst, batch, height * width);
  return check_launch("totensor_u8_kernel");
}

int resize_batch_u8(const void* src, const long long* src_off, const int* src_hw, void* dst, int batch, int height, int width, int swap_rb,
                    float* yx_min, float* yx_max, int slots, cudaStream_t stream) {
  YB_REQUIRE(src && src_off && src_hw && dst && batch > 0 && height > 0 && width > 0, "resize_batch_u8: bad argument");
  YB_REQUIRE((yx_min == nullptr) == (yx_max == nullptr) && slots >= 0, "resize_batch_u8: boxes come as a (yx_min, yx_max) pair");
  const int pixels = height * width;
  resize_u8_kernel<<<dim3((pixels + 255) / 256, batch), 256, 0, stream>>>(static_cast<const uint8_t*>(src), src_off, src_hw, static_cast<uint8_t*>(dst),
                                                                          height, width, swap_rb);
  int rc = check_launch("resize_u8_kernel");
  if (rc) return rc;
  if (yx_min != nullptr && slots > 0) {
    rescale_boxes_kernel<<<(batch * slots + 127) / 128, 128, 0, stream>>>(yx_min, yx_max, src_hw, batch, slots, height, width);
    rc = check_launch("rescale_boxes_kernel");
  }
  return rc;
}

}  // namespace yb
