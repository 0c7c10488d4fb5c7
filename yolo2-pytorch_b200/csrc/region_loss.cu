// K8 + K9: YOLOv2 region loss, forward values and closed-form gradient w.r.t. the head feature map.
// Replaces model.loss and its helpers iou_match / fit_positive / fill_norm
// (/root/reference model/__init__.py:59-107,138-167), which in the reference are dozens of small
// torch kernels, boolean-mask gathers (host syncs) and per-image Python loops.
//
// Semantics restated from the reference (torch 0.3.1 behaviour, see oracle/yolo2_oracle.py: loss):
//   * predictions are decoded exactly like model.Inference.forward (:117-135);
//   * iou_match (:59-73): for every (cell, anchor) the best IoU against ALL G ground-truth slots of the
//     image, zero-padded slots included, ties -> lowest slot;  IoU op order of utils/iou/torch.py;
//   * fit_positive (:76-95): every valid GT (yx_min < yx_max) marks (cell of its centre, anchor with
//     the best centred IoU) positive; duplicates collapse;
//   * negative = !positive & best_iou < threshold (:145);
//   * regression / class targets come from the IoU-MATCHED GT, not from the GT that made the cell
//     positive (:142,146); centre target = frac(centre), size target = log(size / anchor);
//   * terms: foreground  sum_pos (iou - best_iou)^2, background sum_neg iou^2, center, size (sums of
//     squares over positives), cls = mean over positives of cross-entropy (train/cross_entropy = 1) or
//     sum_pos |softmax - onehot|^2; every term is divided by cnt = B * cells * A (:164-166).
// Gradient (SURVEY section 8a derived spec), f = feature viewed [B, cells, A, 5 + C], s = sigmoid:
//   dfg/df0 = 2 (s0 - t_iou) s0 (1 - s0) [pos] / cnt        dbg/df0 = 2 s0 * s0 (1 - s0) [neg] / cnt
//   dcenter/df1,2 = 2 (s - t_c) s (1 - s) [pos] / cnt         dsize/df3,4 = 2 (f - t_s) [pos] / cnt
//   dcls/df5.. = (softmax - onehot) [pos] / (Npos * cnt)      (cross-entropy form)
// The forward pass stores these per-term gradients UNWEIGHTED (grad: feature layout, channel 0 holds the
// foreground part; grad_bg: the background part of channel 0); region_loss_backward combines them with the
// five upstream weights (hparam * grad_output) read from device memory, so no host sync is needed.
#include "yb_common.h"
#include <stdint.h>

namespace yb {

__device__ __forceinline__ float rl_sigmoid(float v) { return 1.f / (1.f + expf(-v)); }

__device__ __forceinline__ float rl_iou(float ymin1, float xmin1, float ymax1, float xmax1, float ymin2, float xmin2, float ymax2,
                                        float xmax2) {
  const float h = fmaxf(__fsub_rn(fminf(ymax1, ymax2), fmaxf(ymin1, ymin2)), 0.f);
  const float w = fmaxf(__fsub_rn(fminf(xmax1, xmax2), fmaxf(xmin1, xmin2)), 0.f);
  const float inter = __fmul_rn(h, w);
  const float a1 = __fmul_rn(__fsub_rn(ymax1, ymin1), __fsub_rn(xmax1, xmin1));
  const float a2 = __fmul_rn(__fsub_rn(ymax2, ymin2), __fsub_rn(xmax2, xmin2));
  const float uni = fmaxf(__fsub_rn(__fadd_rn(a1, a2), inter), 1.1920928955078125e-07f);
  return __fdiv_rn(inter, uni);
}

struct LossParams {
  const float* feature;   // [B, A*(5+C), rows, cols]
  const float* anchors;   // [A, 2] (h, w)
  const float* gt_min;    // [B, G, 2] grid units
  const float* gt_max;    // [B, G, 2]
  const long long* gt_cls;  // [B, G]
  int batch, rows, cols, num_anchors, num_cls, num_gt;
  float threshold;
  int cross_entropy;
  unsigned char* positive;  // [B, cells, A]
  unsigned char* negative;  // [B, cells, A]
  float* best_iou;          // [B, cells, A]
  int* pos_count;           // [B]
  float* partial;           // [B, 5]
  float* losses;            // [5]
  float* grad;              // [B, A*(5+C), rows, cols] unweighted per-term gradients
  float* grad_bg;           // [B, A, cells]
};

// ---- pass 1: positive mask (fit_positive) + per-image positive count ------------------------------
__global__ void region_assign_kernel(const LossParams p) {
  const int img = blockIdx.x;
  const int cells = p.rows * p.cols;
  const int nbox = cells * p.num_anchors;
  unsigned char* pos = p.positive + static_cast<long long>(img) * nbox;
  for (int i = threadIdx.x; i < nbox; i += blockDim.x) pos[i] = 0;
  __syncthreads();
  for (int g = threadIdx.x; g < p.num_gt; g += blockDim.x) {
    const float2 a = reinterpret_cast<const float2*>(p.gt_min)[static_cast<long long>(img) * p.num_gt + g];
    const float2 b = reinterpret_cast<const float2*>(p.gt_max)[static_cast<long long>(img) * p.num_gt + g];
    if (!(a.x < b.x && a.y < b.y)) continue;                    // padded / degenerate slot
    const float cy = __fdiv_rn(__fadd_rn(a.x, b.x), 2.f), cx = __fdiv_rn(__fadd_rn(a.y, b.y), 2.f);
    const int i = static_cast<int>(floorf(cy)), j = static_cast<int>(floorf(cx));
    if (i < 0 || i >= p.rows || j < 0 || j >= p.cols) continue;  // the reference would index out of range here
    int best_a = 0;
    float best = -1.f;
    for (int k = 0; k < p.num_anchors; ++k) {
      const float ah = __fdiv_rn(p.anchors[2 * k], 2.f), aw = __fdiv_rn(p.anchors[2 * k + 1], 2.f);
      const float v = rl_iou(__fsub_rn(a.x, cy), __fsub_rn(a.y, cx), __fsub_rn(b.x, cy), __fsub_rn(b.y, cx), -ah, -aw, ah, aw);
      if (v > best) { best = v; best_a = k; }
    }
    pos[(i * p.cols + j) * p.num_anchors + best_a] = 1;
  }
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < nbox; i += blockDim.x) c += pos[i];
  __shared__ int red[32];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += red[w];
    p.pos_count[img] = t;
  }
}

// ---- pass 2: matching, masks, loss partial sums, unweighted gradients ------------------------------
constexpr int kLossThreads = 256;
constexpr int kMaxGt = 128;

__global__ void __launch_bounds__(kLossThreads) region_main_kernel(const LossParams p) {
  __shared__ float4 gt[kMaxGt];       // (ymin, xmin, ymax, xmax)
  __shared__ int gcls[kMaxGt];
  __shared__ float red[5][kLossThreads / 32];
  const int img = blockIdx.x;
  const int cells = p.rows * p.cols;
  const int per = 5 + p.num_cls;
  const int nbox = cells * p.num_anchors;
  for (int g = threadIdx.x; g < p.num_gt; g += blockDim.x) {
    const float2 a = reinterpret_cast<const float2*>(p.gt_min)[static_cast<long long>(img) * p.num_gt + g];
    const float2 b = reinterpret_cast<const float2*>(p.gt_max)[static_cast<long long>(img) * p.num_gt + g];
    gt[g] = make_float4(a.x, a.y, b.x, b.y);
    // single-class heads (A*5 channels, model.output_channels) carry no class targets: gt_cls may be NULL.  Ids outside
    // [0, C) (corrupt labels) are clamped so the logit indexing below stays inside the feature map.
    int cls_id = 0;
    if (p.num_cls > 0 && p.gt_cls != nullptr) {
      const long long raw = p.gt_cls[static_cast<long long>(img) * p.num_gt + g];
      cls_id = raw < 0 ? 0 : (raw >= p.num_cls ? p.num_cls - 1 : static_cast<int>(raw));
    }
    gcls[g] = cls_id;
  }
  int npos = 0;
  for (int b = 0; b < p.batch; ++b) npos += p.pos_count[b];
  __syncthreads();
  const float cnt = static_cast<float>(p.batch) * cells * p.num_anchors;
  const float inv_cnt = 1.f / cnt;
  const float* fb = p.feature + static_cast<long long>(img) * p.num_anchors * per * cells;
  float* gb = p.grad + static_cast<long long>(img) * p.num_anchors * per * cells;
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  // box index = a * cells + cell (cell fastest) so feature/grad accesses are coalesced
  for (int idx = threadIdx.x; idx < nbox; idx += blockDim.x) {
    const int a = idx / cells, cell = idx - a * cells;
    const float* f = fb + static_cast<long long>(a) * per * cells + cell;
    float* gq = gb + static_cast<long long>(a) * per * cells + cell;
    const float f0 = f[0], f1 = f[cells], f2 = f[2 * cells], f3 = f[3 * cells], f4 = f[4 * cells];
    const float s0 = rl_sigmoid(f0), s1 = rl_sigmoid(f1), s2 = rl_sigmoid(f2);
    const float gy = static_cast<float>(cell / p.rows), gx = static_cast<float>(cell % p.rows);   // meshgrid quirk, :53-56
    const float ah = p.anchors[2 * a], aw = p.anchors[2 * a + 1];
    const float cy = __fadd_rn(gy, s1), cx = __fadd_rn(gx, s2);
    const float hh = __fdiv_rn(__fmul_rn(expf(f3), ah), 2.f), hw = __fdiv_rn(__fmul_rn(expf(f4), aw), 2.f);
    const float ymin = __fsub_rn(cy, hh), xmin = __fsub_rn(cx, hw), ymax = __fadd_rn(cy, hh), xmax = __fadd_rn(cx, hw);
    float best = -1.f;
    int bi = 0;
    for (int g = 0; g < p.num_gt; ++g) {
      const float4 q = gt[g];
      const float v = rl_iou(ymin, xmin, ymax, xmax, q.x, q.y, q.z, q.w);
      if (v > best) { best = v; bi = g; }
    }
    const long long box = (static_cast<long long>(img) * cells + cell) * p.num_anchors + a;
    const bool pos = p.positive[box] != 0;
    const bool neg = !pos && (best < p.threshold);
    p.negative[box] = neg ? 1 : 0;
    p.best_iou[box] = best;
    float g0 = 0.f, gbg = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f, g4 = 0.f;
    const float ds0 = s0 * (1.f - s0);
    if (neg) {
      acc[1] += s0 * s0;
      gbg = 2.f * s0 * ds0 * inv_cnt;
    }
    if (pos) {
      const float4 m = gt[bi];
      const float mcy = __fdiv_rn(__fadd_rn(m.x, m.z), 2.f), mcx = __fdiv_rn(__fadd_rn(m.y, m.w), 2.f);
      const float t1 = __fsub_rn(mcy, floorf(mcy)), t2 = __fsub_rn(mcx, floorf(mcx));
      const float t3 = logf(__fdiv_rn(__fsub_rn(m.z, m.x), ah)), t4 = logf(__fdiv_rn(__fsub_rn(m.w, m.y), aw));
      const float d0 = s0 - best, d1 = s1 - t1, d2 = s2 - t2, d3 = f3 - t3, d4 = f4 - t4;
      acc[0] += d0 * d0;
      acc[2] += d1 * d1 + d2 * d2;
      acc[3] += d3 * d3 + d4 * d4;
      g0 = 2.f * d0 * ds0 * inv_cnt;
      g1 = 2.f * d1 * s1 * (1.f - s1) * inv_cnt;
      g2 = 2.f * d2 * s2 * (1.f - s2) * inv_cnt;
      g3 = 2.f * d3 * inv_cnt;
      g4 = 2.f * d4 * inv_cnt;
    }
    gq[0] = g0;
    gq[cells] = g1;
    gq[2 * cells] = g2;
    gq[3 * cells] = g3;
    gq[4 * cells] = g4;
    p.grad_bg[(static_cast<long long>(img) * p.num_anchors + a) * cells + cell] = gbg;
    if (p.num_cls > 0) {
      if (pos) {
        const int y = gcls[bi];
        float mx = -INFINITY;
        for (int c = 0; c < p.num_cls; ++c) mx = fmaxf(mx, f[(5 + c) * cells]);
        float sum = 0.f;
        for (int c = 0; c < p.num_cls; ++c) sum += expf(f[(5 + c) * cells] - mx);
        if (p.cross_entropy) {
          acc[4] += -(f[(5 + y) * cells] - mx - logf(sum));
          const float sc = inv_cnt / static_cast<float>(npos > 0 ? npos : 1);
          for (int c = 0; c < p.num_cls; ++c) {
            const float pr = expf(f[(5 + c) * cells] - mx) / sum;
            gq[(5 + c) * cells] = (pr - (c == y ? 1.f : 0.f)) * sc;
          }
        } else {
          // sum_pos |softmax - onehot|^2 ; d/dlogit_k = 2 p_k (d_k - sum_c d_c p_c)
          float dot = 0.f;
          for (int c = 0; c < p.num_cls; ++c) {
            const float pr = expf(f[(5 + c) * cells] - mx) / sum;
            const float d = pr - (c == y ? 1.f : 0.f);
            acc[4] += d * d;
            dot += d * pr;
          }
          for (int c = 0; c < p.num_cls; ++c) {
            const float pr = expf(f[(5 + c) * cells] - mx) / sum;
            const float d = pr - (c == y ? 1.f : 0.f);
            gq[(5 + c) * cells] = 2.f * pr * (d - dot) * inv_cnt;
          }
        }
      } else {
        for (int c = 0; c < p.num_cls; ++c) gq[(5 + c) * cells] = 0.f;
      }
    }
  }
  // deterministic per-image partial sums
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    float v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float t = 0.f;
    for (int w = 0; w < kLossThreads / 32; ++w) t += red[threadIdx.x][w];
    p.partial[img * 5 + threadIdx.x] = t;
  }
}

// ---- pass 3: finalize the five scalars (sequential over images: deterministic) ----------------------
__global__ void region_finalize_kernel(const LossParams p) {
  const int k = threadIdx.x;
  if (k >= 5) return;
  const int cells = p.rows * p.cols;
  const float cnt = static_cast<float>(p.batch) * cells * p.num_anchors;
  float t = 0.f;
  for (int b = 0; b < p.batch; ++b) t += p.partial[b * 5 + k];
  if (k == 4 && p.cross_entropy) {
    int npos = 0;
    for (int b = 0; b < p.batch; ++b) npos += p.pos_count[b];
    t = npos > 0 ? t / static_cast<float>(npos) : 0.f;      // F.cross_entropy: mean over the selected rows
  }
  p.losses[k] = t / cnt;
}

// dfeature = w_fg * grad[ch 0] + w_bg * grad_bg (ch 0), w_center * grad[ch 1,2], w_size * grad[ch 3,4], w_cls * grad[ch 5..]
__global__ void region_backward_kernel(const float* __restrict__ grad, const float* __restrict__ grad_bg, const float* __restrict__ weights,
                                       float* __restrict__ dfeature, int batch, int num_anchors, int per, int cells) {
  const long long total = static_cast<long long>(batch) * num_anchors * per * cells;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cell = static_cast<int>(idx % cells);
  const long long t = idx / cells;
  const int j = static_cast<int>(t % per);
  const long long ba = t / per;                       // img * A + a
  float v;
  if (j == 0) v = weights[0] * grad[idx] + weights[1] * grad_bg[ba * cells + cell];
  else if (j < 3) v = weights[2] * grad[idx];
  else if (j < 5) v = weights[3] * grad[idx];
  else v = weights[4] * grad[idx];
  dfeature[idx] = v;
}

int region_loss_forward(const float* feature, const float* anchors, const float* gt_min, const float* gt_max, const long long* gt_cls,
                        int batch, int rows, int cols, int num_anchors, int num_cls, int num_gt, float threshold, int cross_entropy,
                        float* losses, unsigned char* positive, unsigned char* negative, float* best_iou, int* pos_count, float* partial,
                        float* grad, float* grad_bg, cudaStream_t stream) {
  YB_REQUIRE(feature && anchors && gt_min && gt_max && losses && positive && negative && best_iou && pos_count && partial && grad && grad_bg,
             "region_loss: null pointer");
  YB_REQUIRE(num_cls <= 1 || gt_cls != nullptr, "region_loss: class targets required");
  YB_REQUIRE(batch > 0 && rows > 0 && cols > 0 && num_anchors > 0 && num_gt > 0 && num_gt <= kMaxGt, "region_loss: bad shape (G <= %d)", kMaxGt);
  LossParams p;
  p.feature = feature; p.anchors = anchors; p.gt_min = gt_min; p.gt_max = gt_max; p.gt_cls = gt_cls;
  p.batch = batch; p.rows = rows; p.cols = cols; p.num_anchors = num_anchors; p.num_cls = num_cls > 1 ? num_cls : 0; p.num_gt = num_gt;
  p.threshold = threshold; p.cross_entropy = cross_entropy;
  p.positive = positive; p.negative = negative; p.best_iou = best_iou; p.pos_count = pos_count; p.partial = partial; p.losses = losses;
  p.grad = grad; p.grad_bg = grad_bg;
  region_assign_kernel<<<batch, 128, 0, stream>>>(p);
  int rc = check_launch("region_assign_kernel");
  if (rc) return rc;
  region_main_kernel<<<batch, kLossThreads, 0, stream>>>(p);
  rc = check_launch("region_main_kernel");
  if (rc) return rc;
  region_finalize_kernel<<<1, 32, 0, stream>>>(p);
  return check_launch("region_finalize_kernel");
}

int region_loss_backward(const float* grad, const float* grad_bg, const float* weights, float* dfeature, int batch, int rows, int cols,
                         int num_anchors, int num_cls, cudaStream_t stream) {
  YB_REQUIRE(grad && grad_bg && weights && dfeature && batch > 0, "region_loss_backward: bad argument");
  const int per = 5 + (num_cls > 1 ? num_cls : 0);
  const long long total = static_cast<long long>(batch) * num_anchors * per * rows * cols;
  region_backward_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(grad, grad_bg, weights, dfeature, batch, num_anchors,
                                                                                         per, rows * cols);
  return check_launch("region_backward_kernel");
}

}  // namespace yb
