// K5: anchor-box decode + class softmax in one pass over the head's output.
// Replaces model.Inference.forward (/root/reference model/__init__.py:117-135: permute+contiguous,
// sigmoid, meshgrid, exp * anchors, +-size/2) and F.softmax(logits, -1) (detect.py:152).
//
// feature is the reference-visible fp32 NCHW tensor [B, A*(5+C), rows, cols]; per-anchor channel
// order is (iou, y, x, h, w, cls...).  A block stages a [A*(5+C)] x 32-cell slab in shared memory
// with coalesced 128-byte row reads, then threads emit outputs in [cell][anchor] order so every
// output array is written contiguously.
#include "yb_common.h"
#include <stdint.h>

namespace yb {

constexpr int kCellsPerBlock = 32;

__device__ __forceinline__ float sigmoidf_acc(float v) { return 1.f / (1.f + expf(-v)); }

__global__ void decode_kernel(const float* __restrict__ feature, const float* __restrict__ anchors, float* __restrict__ iou,
                              float* __restrict__ center_offset, float* __restrict__ size_norm, float* __restrict__ yx_min,
                              float* __restrict__ yx_max, float* __restrict__ logits, float* __restrict__ prob, int rows, int cols,
                              int num_anchors, int num_cls) {
  extern __shared__ float slab[];  // [channels][kCellsPerBlock + 1]
  const int per = 5 + num_cls;
  const int channels = num_anchors * per;
  const int cells = rows * cols;
  const int img = blockIdx.y;
  const int cell0 = blockIdx.x * kCellsPerBlock;
  const int ncell = min(kCellsPerBlock, cells - cell0);
  const float* fb = feature + static_cast<long long>(img) * channels * cells;
  for (int i = threadIdx.x; i < channels * kCellsPerBlock; i += blockDim.x) {
    const int ch = i / kCellsPerBlock, cl = i % kCellsPerBlock;
    slab[ch * (kCellsPerBlock + 1) + cl] = (cl < ncell) ? __ldg(fb + static_cast<long long>(ch) * cells + cell0 + cl) : 0.f;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ncell * num_anchors; t += blockDim.x) {
    const int cl = t / num_anchors, a = t % num_anchors;
    const int cell = cell0 + cl;
    const float* f = slab + (a * per) * (kCellsPerBlock + 1) + cl;
#define YB_F(j) f[(j) * (kCellsPerBlock + 1)]
    const long long box = (static_cast<long long>(img) * cells + cell) * num_anchors + a;
    const float s0 = sigmoidf_acc(YB_F(0)), s1 = sigmoidf_acc(YB_F(1)), s2 = sigmoidf_acc(YB_F(2));
    const float n0 = YB_F(3), n1 = YB_F(4);
    // meshgrid (model/__init__.py:53-56): (y, x) = (k / rows, k % rows)
    const float gy = static_cast<float>(cell / rows), gx = static_cast<float>(cell % rows);
    const float cy = __fadd_rn(gy, s1), cx = __fadd_rn(gx, s2);
    const float hh = __fdiv_rn(__fmul_rn(expf(n0), __ldg(anchors + 2 * a)), 2.f);
    const float hw = __fdiv_rn(__fmul_rn(expf(n1), __ldg(anchors + 2 * a + 1)), 2.f);
    iou[box] = s0;
    reinterpret_cast<float2*>(center_offset)[box] = make_float2(s1, s2);
    reinterpret_cast<float2*>(size_norm)[box] = make_float2(n0, n1);
    reinterpret_cast<float2*>(yx_min)[box] = make_float2(__fsub_rn(cy, hh), __fsub_rn(cx, hw));
    reinterpret_cast<float2*>(yx_max)[box] = make_float2(__fadd_rn(cy, hh), __fadd_rn(cx, hw));
    if (num_cls > 1) {
      float mx = -INFINITY;
      for (int c = 0; c < num_cls; ++c) mx = fmaxf(mx, YB_F(5 + c));
      float sum = 0.f;
      for (int c = 0; c < num_cls; ++c) sum += expf(YB_F(5 + c) - mx);
      float* lg = logits + box * num_cls;
      float* pb = prob ? prob + box * num_cls : nullptr;
      for (int c = 0; c < num_cls; ++c) {
        const float v = YB_F(5 + c);
        lg[c] = v;
        if (pb) pb[c] = expf(v - mx) / sum;
      }
    } else if (prob) {
      prob[box] = 1.f;  // detect.py:47-48
    }
#undef YB_F
  }
}

int decode_forward(const float* feature, const float* anchors, float* iou, float* center_offset, float* size_norm, float* yx_min,
                   float* yx_max, float* logits, float* prob, int batch, int rows, int cols, int num_anchors, int num_cls,
                   cudaStream_t stream) {
  YB_REQUIRE(feature && anchors && iou && center_offset && size_norm && yx_min && yx_max, "decode: null pointer");
  YB_REQUIRE(batch > 0 && rows > 0 && cols > 0 && num_anchors > 0 && num_cls >= 1, "decode: bad shape");
  YB_REQUIRE(num_cls == 1 || logits != nullptr, "decode: logits buffer required when num_cls > 1");
  YB_REQUIRE(batch <= 65535, "decode: batch too large");
  const int per = 5 + (num_cls > 1 ? num_cls : 0);
  const int channels = num_anchors * per;
  const size_t smem = static_cast<size_t>(channels) * (kCellsPerBlock + 1) * sizeof(float);
  YB_REQUIRE(smem <= 200 * 1024, "decode: %d channels exceed the shared-memory slab", channels);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    YB_CUDA(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    smem_set = smem;
  }
  dim3 grid((rows * cols + kCellsPerBlock - 1) / kCellsPerBlock, batch);
  decode_kernel<<<grid, 160, smem, stream>>>(feature, anchors, iou, center_offset, size_norm, yx_min, yx_max, logits, prob, rows, cols,
                                             num_anchors, num_cls > 1 ? num_cls : 0);
  return check_launch("decode_kernel");
}

}  // namespace yb
