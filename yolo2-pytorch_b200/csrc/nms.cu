// K6 + K7: per-image threshold filter, stable descending sort, greedy NMS and per-class expansion
// in ONE launch with one CTA per image -- no host round trips.
// Replaces detect.filter_visible / detect.postprocess (/root/reference detect.py:51-80) and the
// Python `while` loop of utils.postprocess.nms (utils/postprocess.py:23-49, ~10 kernel launches and
// 2 host syncs per kept box in the reference).
//
// Exactness contract: survivor indices are bit-identical to the reference on identical fp32
// inputs.  IoU follows utils/iou/torch.py:24-61 operation by operation with round-to-nearest
// intrinsics (no FMA contraction):  h = max(min(ymax1,ymax2) - max(ymin1,ymin2), 0), same for w,
// inter = h*w, area = (ymax-ymin)*(xmax-xmin), union = max(area1 + area2 - inter, eps32),
// keep iff inter/union <= overlap.  Sorting is by (score descending, index ascending).
#include "yb_common.h"
#include <stdint.h>

namespace yb {

constexpr int kNmsThreads = 1024;   // one CTA per image; the phases are latency-bound, so go wide
constexpr int kMaxLimit = 1024;

__device__ __forceinline__ uint32_t float_orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending numeric order as unsigned
}

__device__ __forceinline__ float iou_exact(float ymin1, float xmin1, float ymax1, float xmax1, float area1, float ymin2, float xmin2,
                                           float ymax2, float xmax2, float area2, float eps) {
  const float h = fmaxf(__fsub_rn(fminf(ymax1, ymax2), fmaxf(ymin1, ymin2)), 0.f);
  const float w = fmaxf(__fsub_rn(fminf(xmax1, xmax2), fmaxf(xmin1, xmin2)), 0.f);
  const float inter = __fmul_rn(h, w);
  const float uni = fmaxf(__fsub_rn(__fadd_rn(area1, area2), inter), eps);
  return __fdiv_rn(inter, uni);
}

struct NmsParams {
  const float* score;   // [B, N]  (objectness `iou`)
  const float* yx_min;  // [B, N, 2]
  const float* yx_max;  // [B, N, 2]
  const float* prob;    // [B, N, C] or null (mode 2)
  int n, num_cls;
  int mode;             // 0: score > threshold; 1 ("fix"): score * max_c prob > threshold_cls; 2: no filter
  float threshold, threshold_cls, overlap;
  int limit;
  int n_pad;            // power of two >= n
  int det_cap;          // per-image capacity of the det_* arrays
  int* n_filtered;      // [B]
  int* n_keep;          // [B]
  int* keep_idx;        // [B, limit]  index into the FILTERED arrays (what utils.postprocess.nms returns)
  int* keep_box;        // [B, limit]  index into the original N boxes
  int* n_det;           // [B]      (nullable) per-class expansion, detect.py:72-77
  int* det_keep;        // [B, det_cap] rank in the keep list
  int* det_cls;         // [B, det_cap]
  float* det_score;     // [B, det_cap]
  int* filt_box;        // [B, n] (nullable) filtered rank -> input box (ascending: detect.py:57-62 order)
  int* best_cls;        // [B, n] (nullable) argmax_c prob   (detect.py:52)
  float* best_prob;     // [B, n] (nullable) max_c prob
};

__global__ void __launch_bounds__(kNmsThreads) filter_nms_kernel(const NmsParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);             // [n_pad]
  int* filt = reinterpret_cast<int*>(keys + p.n_pad);                                  // [n_pad] filtered rank -> box
  float4* cbox = reinterpret_cast<float4*>(filt + p.n_pad);                            // [limit] (ymin,xmin,ymax,xmax)
  float* carea = reinterpret_cast<float*>(cbox + p.limit);                             // [limit]
  int* cfilt = reinterpret_cast<int*>(carea + p.limit);                                // [limit] filtered rank of candidate
  uint32_t* sup = reinterpret_cast<uint32_t*>(cfilt + p.limit);                        // [limit][words]
  const int words = (p.limit + 31) >> 5;
  int* keep_rank = reinterpret_cast<int*>(sup + p.limit * words);                      // [limit] candidate rank of each kept
  __shared__ int warp_tot[kNmsThreads / 32];
  __shared__ int s_count, s_keep;

  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* score = p.score + static_cast<long long>(img) * p.n;
  const float2* bmin = reinterpret_cast<const float2*>(p.yx_min) + static_cast<long long>(img) * p.n;
  const float2* bmax = reinterpret_cast<const float2*>(p.yx_max) + static_cast<long long>(img) * p.n;
  const float* prob = p.prob ? p.prob + static_cast<long long>(img) * p.n * p.num_cls : nullptr;

  // ---- 1. filter + order-preserving compaction (detect.py:51-63) ----
  if (tid == 0) s_count = 0;
  __syncthreads();
  for (int base = 0; base < p.n; base += kNmsThreads) {
    const int i = base + tid;
    bool pass = false;
    float sc = 0.f;
    if (i < p.n) {
      sc = score[i];
      float mx = -INFINITY;
      if (prob != nullptr && (p.mode == 1 || p.best_cls != nullptr || p.best_prob != nullptr)) {
        int arg = 0;
        const float* pr = prob + static_cast<long long>(i) * p.num_cls;
        if ((p.num_cls & 3) == 0) {
          // 16-byte loads, all issued before the first use (one exposed L2 latency instead of num_cls)
          for (int c0 = 0; c0 < p.num_cls; c0 += 20) {
            float4 q[5];
#pragma unroll
            for (int u = 0; u < 5; ++u)
              if (c0 + 4 * u < p.num_cls) q[u] = __ldg(reinterpret_cast<const float4*>(pr + c0) + u);
#pragma unroll
            for (int u = 0; u < 5; ++u) {
              if (c0 + 4 * u < p.num_cls) {
                const float e[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                for (int z = 0; z < 4; ++z)
                  if (e[z] > mx) { mx = e[z]; arg = c0 + 4 * u + z; }   // first maximum wins, as torch.max does
              }
            }
          }
        } else {
          for (int c = 0; c < p.num_cls; ++c) {
            const float v = pr[c];
            if (v > mx) { mx = v; arg = c; }
          }
        }
        if (p.best_cls) p.best_cls[static_cast<long long>(img) * p.n + i] = arg;
        if (p.best_prob) p.best_prob[static_cast<long long>(img) * p.n + i] = mx;
      }
      if (p.mode == 2) pass = true;
      else if (p.mode == 0) pass = sc > p.threshold;
      else pass = __fmul_rn(sc, mx) > p.threshold_cls;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    int off = s_count;
    for (int w = 0; w < wid; ++w) off += warp_tot[w];
    if (pass) {
      const int rank = off + __popc(bal & ((1u << lane) - 1u));
      filt[rank] = i;
      if (p.filt_box) p.filt_box[static_cast<long long>(img) * p.n + rank] = i;
      // descending score, ascending filtered rank
      keys[rank] = (static_cast<unsigned long long>(~float_orderable(sc)) << 32) | static_cast<unsigned>(rank);
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < kNmsThreads / 32; ++w) tot += warp_tot[w];
      s_count += tot;
    }
    __syncthreads();
  }
  const int nf = s_count;
  if (tid == 0) p.n_filtered[img] = nf;
  if (nf == 0) {
    if (tid == 0) { p.n_keep[img] = 0; if (p.n_det) p.n_det[img] = 0; }
    return;
  }
  // ---- 2. bitonic sort of the filtered keys (pad to a power of two with +inf keys) ----
  int n_sort = 1;
  while (n_sort < nf) n_sort <<= 1;
  for (int i = nf + tid; i < n_sort; i += kNmsThreads) keys[i] = ~0ull;
  __syncthreads();
  for (int k = 2; k <= n_sort; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n_sort; i += kNmsThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  // ---- 3. top-`limit` candidates (utils/postprocess.py:37-38) ----
  const int nc = nf < p.limit ? nf : p.limit;
  for (int i = tid; i < nc; i += kNmsThreads) {
    const int rank = static_cast<int>(keys[i] & 0xffffffffu);
    const int box = filt[rank];
    const float2 a = bmin[box], b = bmax[box];
    cbox[i] = make_float4(a.x, a.y, b.x, b.y);
    carea[i] = __fmul_rn(__fsub_rn(b.x, a.x), __fsub_rn(b.y, a.y));
    cfilt[i] = rank;
  }
  __syncthreads();
  // ---- 4. suppression bit-matrix: sup[i] bit j (j > i) set iff NOT (iou(i, j) <= overlap) ----
  const float eps = 1.1920928955078125e-07f;  // float32 eps, utils/iou/torch.py:47
  for (int e = tid; e < nc * words; e += kNmsThreads) {
    const int i = e / words, w = e % words;
    uint32_t bits = 0;
    const float4 bi = cbox[i];
    const float ai = carea[i];
    const int j0 = w * 32;
    for (int b = 0; b < 32; ++b) {
      const int j = j0 + b;
      if (j > i && j < nc) {
        const float4 bj = cbox[j];
        const float v = iou_exact(bi.x, bi.y, bi.z, bi.w, ai, bj.x, bj.y, bj.z, bj.w, carea[j], eps);
        if (!(v <= p.overlap)) bits |= 1u << b;
      }
    }
    sup[e] = bits;
  }
  __syncthreads();
  // ---- 5. greedy scan in score order (one warp; lane w owns word w of the removed set) ----
  if (wid == 0) {
    uint32_t removed = 0;  // lanes >= words unused
    int nk = 0;
    for (int i = 0; i < nc; ++i) {
      const uint32_t wv = __shfl_sync(0xffffffffu, removed, i >> 5);
      if (!((wv >> (i & 31)) & 1u)) {
        if (lane == 0) keep_rank[nk] = i;
        ++nk;
        if (lane < words) removed |= sup[i * words + lane];
      }
    }
    if (lane == 0) { s_keep = nk; p.n_keep[img] = nk; }
  }
  __syncthreads();
  const int nk = s_keep;
  for (int k = tid; k < nk; k += kNmsThreads) {
    const int c = keep_rank[k];
    p.keep_idx[static_cast<long long>(img) * p.limit + k] = cfilt[c];
    p.keep_box[static_cast<long long>(img) * p.limit + k] = filt[cfilt[c]];
  }
  // ---- 6. per-class expansion (detect.py:72-77): (kept, cls) pairs with iou*prob > threshold_cls,
  //         in mask.nonzero() order (kept rank major, class minor) ----
  if (p.n_det == nullptr) return;
  if (p.mode != 1) { if (tid == 0) p.n_det[img] = 0; return; }
  __syncthreads();
  int* cnt = reinterpret_cast<int*>(keys);  // reuse: [nk + 1] exclusive offsets
  uint32_t* cmask = sup;                    // reuse: passing-class bitmask per kept box (num_cls <= 32 fast path)
  const bool fast = (p.num_cls <= 32) && ((p.num_cls & 3) == 0);
  for (int k = tid; k < nk; k += kNmsThreads) {
    const int box = filt[cfilt[keep_rank[k]]];
    const float sc = score[box];
    const float* pr = prob + static_cast<long long>(box) * p.num_cls;
    int c_pass = 0;
    if (fast) {
      uint32_t m = 0;
#pragma unroll 5
      for (int c0 = 0; c0 < p.num_cls; c0 += 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(pr + c0));
        m |= (__fmul_rn(sc, q.x) > p.threshold_cls ? 1u : 0u) << c0;
        m |= (__fmul_rn(sc, q.y) > p.threshold_cls ? 2u : 0u) << c0;
        m |= (__fmul_rn(sc, q.z) > p.threshold_cls ? 4u : 0u) << c0;
        m |= (__fmul_rn(sc, q.w) > p.threshold_cls ? 8u : 0u) << c0;
      }
      cmask[k] = m;
      c_pass = __popc(m);
    } else {
      for (int c = 0; c < p.num_cls; ++c) c_pass += (__fmul_rn(sc, pr[c]) > p.threshold_cls) ? 1 : 0;
    }
    cnt[k + 1] = c_pass;
  }
  if (tid == 0) cnt[0] = 0;
  __syncthreads();
  if (wid == 0) {
    // inclusive warp scan over cnt[1..nk] in chunks of 32
    int carry = 0;
    for (int base = 0; base < nk; base += 32) {
      const int k = base + lane;
      int v = (k < nk) ? cnt[k + 1] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      if (k < nk) cnt[k + 1] = v + carry;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
    if (lane == 0) p.n_det[img] = carry < p.det_cap ? carry : p.det_cap;
  }
  __syncthreads();
  for (int k = tid; k < nk; k += kNmsThreads) {
    const int box = filt[cfilt[keep_rank[k]]];
    const float sc = score[box];
    const float* pr = prob + static_cast<long long>(box) * p.num_cls;
    int o = cnt[k];
    if (fast) {
      uint32_t m = cmask[k];
      while (m) {
        const int c = __ffs(m) - 1;
        m &= m - 1;
        if (o < p.det_cap) {
          const long long d = static_cast<long long>(img) * p.det_cap + o;
          p.det_keep[d] = k;
          p.det_cls[d] = c;
          p.det_score[d] = __fmul_rn(sc, __ldg(pr + c));
        }
        ++o;
      }
    } else {
      for (int c = 0; c < p.num_cls; ++c) {
        const float v = __fmul_rn(sc, pr[c]);
        if (v > p.threshold_cls) {
          if (o < p.det_cap) {
            const long long d = static_cast<long long>(img) * p.det_cap + o;
            p.det_keep[d] = k;
            p.det_cls[d] = c;
            p.det_score[d] = v;
          }
          ++o;
        }
      }
    }
  }
}

int filter_nms(const float* score, const float* yx_min, const float* yx_max, const float* prob, int batch, int n, int num_cls, int mode,
               float threshold, float threshold_cls, float overlap, int limit, int* n_filtered, int* n_keep, int* keep_idx,
               int* keep_box, int* n_det, int* det_keep, int* det_cls, float* det_score, int det_cap, int* filt_box, int* best_cls, float* best_prob,
               cudaStream_t stream) {
  YB_REQUIRE(score && yx_min && yx_max && n_filtered && n_keep && keep_idx && keep_box, "filter_nms: null pointer");
  YB_REQUIRE(mode == 0 || mode == 1 || mode == 2, "filter_nms: mode %d", mode);
  YB_REQUIRE(mode != 1 || (prob != nullptr && num_cls >= 1), "filter_nms: fix mode needs prob");
  YB_REQUIRE(batch > 0 && n >= 0 && limit > 0 && limit <= kMaxLimit, "filter_nms: bad shape (limit <= %d)", kMaxLimit);
  YB_REQUIRE(n <= 16384, "filter_nms: n=%d exceeds 16384 boxes per image", n);
  YB_REQUIRE(n_det == nullptr || (det_keep && det_cls && det_score && det_cap > 0), "filter_nms: det buffers");
  if (n == 0) {
    YB_CUDA(cudaMemsetAsync(n_filtered, 0, sizeof(int) * batch, stream));
    YB_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int) * batch, stream));
    if (n_det) YB_CUDA(cudaMemsetAsync(n_det, 0, sizeof(int) * batch, stream));
    return 0;
  }
  NmsParams p;
  p.score = score; p.yx_min = yx_min; p.yx_max = yx_max; p.prob = prob;
  p.n = n; p.num_cls = num_cls; p.mode = mode; p.threshold = threshold; p.threshold_cls = threshold_cls; p.overlap = overlap;
  p.limit = limit;
  int n_pad = 32;
  while (n_pad < n) n_pad <<= 1;
  if (n_pad < limit + 1) { while (n_pad < limit + 1) n_pad <<= 1; }
  p.n_pad = n_pad; p.det_cap = det_cap;
  p.n_filtered = n_filtered; p.n_keep = n_keep; p.keep_idx = keep_idx; p.keep_box = keep_box;
  p.n_det = n_det; p.det_keep = det_keep; p.det_cls = det_cls; p.det_score = det_score;
  p.filt_box = filt_box; p.best_cls = best_cls; p.best_prob = best_prob;
  const int words = (limit + 31) / 32;
  const size_t smem = static_cast<size_t>(n_pad) * 12 + static_cast<size_t>(limit) * (16 + 4 + 4 + 4 * words + 4) + 64;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    YB_REQUIRE(smem <= 220 * 1024, "filter_nms: shared memory %zu too large", smem);
    YB_CUDA(cudaFuncSetAttribute(filter_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    smem_set = smem;
  }
  filter_nms_kernel<<<batch, kNmsThreads, smem, stream>>>(p);
  return check_launch("filter_nms_kernel");
}

// ------------------------------------------------------------------------------------------
// IoU matrices (utils/iou/torch.py:47-61 iou_matrix, :139-153 batch_iou_matrix): [B, N1, N2]
__global__ void iou_matrix_kernel(const float2* __restrict__ min1, const float2* __restrict__ max1, const float2* __restrict__ min2,
                                  const float2* __restrict__ max2, float* __restrict__ out, int batch, int n1, int n2, float eps) {
  const long long total = static_cast<long long>(batch) * n1 * n2;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = static_cast<int>(idx % n2);
  const long long t = idx / n2;
  const int i = static_cast<int>(t % n1);
  const int b = static_cast<int>(t / n1);
  const float2 a0 = min1[static_cast<long long>(b) * n1 + i], a1 = max1[static_cast<long long>(b) * n1 + i];
  const float2 b0 = min2[static_cast<long long>(b) * n2 + j], b1 = max2[static_cast<long long>(b) * n2 + j];
  const float area1 = __fmul_rn(__fsub_rn(a1.x, a0.x), __fsub_rn(a1.y, a0.y));
  const float area2 = __fmul_rn(__fsub_rn(b1.x, b0.x), __fsub_rn(b1.y, b0.y));
  out[idx] = iou_exact(a0.x, a0.y, a1.x, a1.y, area1, b0.x, b0.y, b1.x, b1.y, area2, eps);
}

int iou_matrix(const float* yx_min1, const float* yx_max1, const float* yx_min2, const float* yx_max2, float* out, int batch, int n1,
               int n2, float min_union, cudaStream_t stream) {
  YB_REQUIRE(batch >= 0 && n1 >= 0 && n2 >= 0, "iou_matrix: bad shape");
  const long long total = static_cast<long long>(batch) * n1 * n2;
  if (total == 0) return 0;
  YB_REQUIRE(yx_min1 && yx_max1 && yx_min2 && yx_max2 && out, "iou_matrix: null pointer");
  iou_matrix_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float2*>(yx_min1), reinterpret_cast<const float2*>(yx_max1), reinterpret_cast<const float2*>(yx_min2),
      reinterpret_cast<const float2*>(yx_max2), out, batch, n1, n2, min_union);
  return check_launch("iou_matrix_kernel");
}

// ------------------------------------------------------------------------------------------------
// Evaluation matching (reference eval.py:57-75 `_matching` / `matching`, called per image and per class from
// eval.py:210-216 `filter_cls`): a detection of class c is a true positive iff its best-IoU ground-truth box of class c
// (ties -> lowest index, torch.max) has IoU > threshold and was not already claimed by an earlier detection of the
// image (detections arrive in descending-score order).  One warp per (image, class): the lanes scan the ground truth
// for each detection in turn, the claimed set is a per-warp bitmap.  Segmented (ragged) inputs: image i owns
// detections [det_off[i], det_off[i+1]) and ground-truth boxes [gt_off[i], gt_off[i+1]).
constexpr int kMatchMaxGt = 1024;     // ground-truth boxes per image (bitmap in shared memory)

__global__ void __launch_bounds__(128) eval_match_kernel(const float2* __restrict__ det_min, const float2* __restrict__ det_max,
                                                         const int* __restrict__ det_cls, const int* __restrict__ det_off,
                                                         const float2* __restrict__ gt_min, const float2* __restrict__ gt_max,
                                                         const int* __restrict__ gt_cls, const int* __restrict__ gt_off, int num_cls, float threshold,
                                                         float eps, unsigned char* __restrict__ tp) {
  __shared__ unsigned claimed[4][kMatchMaxGt / 32];
  const int img = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d0 = det_off[img], d1 = det_off[img + 1];
  const int g0 = gt_off[img], g1 = gt_off[img + 1];
  const int ng = g1 - g0;
  for (int c = blockIdx.y * 4 + warp; c < num_cls; c += gridDim.y * 4) {
    for (int i = lane; i < (ng + 31) / 32; i += 32) claimed[warp][i] = 0u;
    __syncwarp();
    for (int d = d0; d < d1; ++d) {
      if (det_cls[d] != c) continue;                       // warp-uniform
      const float2 a0 = det_min[d], a1 = det_max[d];
      const float area_a = __fmul_rn(__fsub_rn(a1.x, a0.x), __fsub_rn(a1.y, a0.y));
      float best = -1.f;
      int best_j = 0x7fffffff;                             // rank among the class-c ground truth is not needed: identity suffices
      for (int j = lane; j < ng; j += 32) {
        if (gt_cls[g0 + j] != c) continue;
        const float2 b0 = gt_min[g0 + j], b1 = gt_max[g0 + j];
        const float area_b = __fmul_rn(__fsub_rn(b1.x, b0.x), __fsub_rn(b1.y, b0.y));
        const float v = iou_exact(a0.x, a0.y, a1.x, a1.y, area_a, b0.x, b0.y, b1.x, b1.y, area_b, eps);
        if (v > best) { best = v; best_j = j; }            // strict: the first (lowest-index) maximum of this lane's stride
      }
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        if (ob > best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
      }
      if (lane == 0) {
        unsigned char hit = 0;
        if (best_j != 0x7fffffff && best > threshold) {
          const unsigned bit = 1u << (best_j & 31);
          if (!(claimed[warp][best_j >> 5] & bit)) { claimed[warp][best_j >> 5] |= bit; hit = 1; }
        }
        tp[d] = hit;
      }
      __syncwarp();
    }
    __syncwarp();
  }
}

int eval_match(const float* det_yx_min, const float* det_yx_max, const int* det_cls, const int* det_off, const float* gt_yx_min,
               const float* gt_yx_max, const int* gt_cls, const int* gt_off, int batch, int num_cls, int max_gt, float threshold, float min_union,
               unsigned char* tp, cudaStream_t stream) {
  YB_REQUIRE(det_off && gt_off && tp && batch > 0 && num_cls > 0, "eval_match: bad argument");
  YB_REQUIRE(max_gt >= 0 && max_gt <= kMatchMaxGt, "eval_match: at most %d ground-truth boxes per image (got %d)", kMatchMaxGt, max_gt);
  const int gy = (num_cls + 3) / 4 < 8 ? (num_cls + 3) / 4 : 8;
  eval_match_kernel<<<dim3(batch, gy), 128, 0, stream>>>(reinterpret_cast<const float2*>(det_yx_min), reinterpret_cast<const float2*>(det_yx_max), det_cls,
                                                         det_off, reinterpret_cast<const float2*>(gt_yx_min), reinterpret_cast<const float2*>(gt_yx_max),
                                                         gt_cls, gt_off, num_cls, threshold, min_union, tp);
  return check_launch("eval_match_kernel");
}

}  // namespace yb
