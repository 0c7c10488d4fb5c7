// K1: convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05), sm_100a only.
//
// Replaces the reference's  nn.Conv2d -> nn.BatchNorm2d(eval) -> nn.LeakyReLU(0.1)  unit
// (/root/reference model/yolo2.py:49-65) for k in {1,3}, stride 1, pad (k-1)/2.
//
//   D[M = B*H*W pixels, N = Cout] = A[M, K = k*k*Cin] * W[N, K]^T
//
//   * A is never materialised: each K-block (one filter tap, BK input channels) of a 128-pixel
//     M-tile is fetched by ONE im2col-mode TMA (cp.async.bulk.tensor.4d...im2col) straight from
//     the NHWC fp16 activation into 128B- (or 64B-) swizzled shared memory; halo pixels are
//     zero-filled by the TMA unit.  W (KRSC fp16, K-major) comes in by a tiled 2D TMA.
//   * one elected thread issues tcgen05.mma (M=128, N=BN, K=16) with the fp32 accumulator in TMEM
//     (two accumulator stages so the epilogue of tile i overlaps the mainloop of tile i+1);
//   * the epilogue warps read TMEM with tcgen05.ld, apply the folded BatchNorm (fp32 scale/shift per
//     channel) + leaky-ReLU and store fp16 NHWC (optionally into a channel slice of a wider buffer,
//     which is how torch.cat at yolo2.py:129 disappears) or fp32 NCHW (the head, the tensor the
//     reference hands to the decoder).
//   * persistent: grid = min(#tiles, #SMs); warp 0 = TMA producer, warp 1 = MMA issuer + TMEM
//     owner, warps 2..5 = epilogue (TMEM lane quadrant = warp_id % 4).
#include "yb_common.h"
#include "yb_ptx.cuh"

namespace yb {

struct ConvParams {
  int m_total;      // B*H*W
  int height, width;
  int cin, cout;
  int ksize, pad;
  int kb_per_tap;   // Cin / BK
  int num_kb;       // ksize*ksize*kb_per_tap
  int m_tiles, n_tiles;
  int a_im2col;     // 1: im2col TMA, 0: plain 2D tiled TMA over [M, Cin] (1x1 only)
  const float* scale;
  const float* shift;
  float slope;
  void* y;
  long long y_ld;   // fp16 NHWC: elements per pixel row of the destination buffer
  int y_ch_off;     // fp16 NHWC: first destination channel
  int out_mode;     // 0: fp16 NHWC, 1: fp32 NCHW
  int hw;           // H*W
  int skip;         // profiling ablation (results are garbage): 1 = no A loads, 2 = no B loads, 4 = no MMA, 8 = no stores
  int* dbg;
};

constexpr int BM = 128;
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 192;
constexpr int kEpiThreads = 128;

// MT = number of 128-pixel M-subtiles per CTA tile (1 or 2).  MT = 2 makes the CTA tile 256 x BN: both
// subtiles reuse the same weight (B) tile from shared memory, halving the L2->SM weight traffic per MAC.
// kPair = true runs two CTAs of a (2,1,1) cluster as one cta_group::2 unit: every MMA is M = 256 (128 rows
// from each CTA), each CTA stages only HALF of the weight tile (N/2 rows) and the tensor cores read the other
// half from the peer SM, so the per-SM operand feed drops again by 25-33%.
template <int BN, int BK, int MT, bool kPair>
struct ConvCfg {
  static constexpr int kSwizzle = BK * 2;                       // bytes per smem row
  static constexpr int kASubBytes = BM * BK * 2;
  static constexpr int kABytes = MT * kASubBytes;
  static constexpr int kBRows = kPair ? BN / 2 : BN;            // weight rows staged by this CTA
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (196608 / kStageBytes) > 8 ? 8 : (196608 / kStageBytes);
  static constexpr int kAccCols = MT * BN;                      // TMEM columns of one accumulator stage
  static constexpr int kAccStages = (2 * kAccCols <= 512) ? 2 : 1;
  static constexpr int kTmemCols = kAccStages * kAccCols;       // power of two in [64, 512]
  static constexpr int kRowsPerCta = BM * MT;
  static constexpr int kRowsPerTile = kRowsPerCta * (kPair ? 2 : 1);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 2 * 2 * BN * 4 /*scale/shift x2*/ + 256 /*barriers*/;
  static_assert(kAccCols <= 512, "accumulator does not fit TMEM");
};

template <int BN, int BK, int MT, bool kPair>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const ConvParams p) {
  using Cfg = ConvCfg<BN, BK, MT, kPair>;
  constexpr int kAccStages = Cfg::kAccStages;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the swizzle atoms (identical offsets in both CTAs of a pair)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + kStages * Cfg::kABytes;
  float* ep_scale = reinterpret_cast<float*>(smem_gen + kStages * Cfg::kStageBytes);  // [2][BN]
  float* ep_shift = ep_scale + 2 * BN;                                                  // [2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ep_shift + 2 * BN);
  const uint32_t bar_full = smem_u32(bars);                 // [kStages]  (pair: only the leader's are used)
  const uint32_t bar_empty = bar_full + 8 * kStages;        // [kStages]
  const uint32_t bar_tfull = bar_empty + 8 * kStages;       // [2]
  const uint32_t bar_tempty = bar_tfull + 16;               // [2]        (pair: only the leader's are used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;     // 0 = leader
  const int unit_id = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_units = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_full + 8 * i, kPair ? 2 : 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, kPair ? 8 : 4);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    if (kPair) { tmem_alloc_pair(smem_u32(tmem_slot), Cfg::kTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit_id; tile < num_tiles; tile += num_units) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int m_cta = m_tile * Cfg::kRowsPerTile + static_cast<int>(rank) * Cfg::kRowsPerCta;
        int img[MT], h0[MT], w0[MT];
        int nsub = 0;                      // subtiles that start inside the tensor (the rest are skipped)
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          const int ms = m_cta + t * BM;
          if (ms < p.m_total) nsub = t + 1;
          img[t] = ms / p.hw;
          const int rem = ms - img[t] * p.hw;
          h0[t] = rem / p.width;
          w0[t] = rem - h0[t] * p.width;
        }
        if (p.skip & 1) nsub = 0;
        uint32_t tx_bytes = nsub * Cfg::kASubBytes + ((p.skip & 2) ? 0 : Cfg::kBBytes);
        if (kPair) {
          // the leader's barrier also counts the peer's bytes: recompute the peer's live subtiles
          const int m_peer = m_tile * Cfg::kRowsPerTile + Cfg::kRowsPerCta;
          int nsub_peer = 0;
#pragma unroll
          for (int t = 0; t < MT; ++t)
            if (m_peer + t * BM < p.m_total) nsub_peer = t + 1;
          if (p.skip & 1) nsub_peer = 0;
          tx_bytes += nsub_peer * Cfg::kASubBytes + ((p.skip & 2) ? 0 : Cfg::kBBytes);     // (only used by rank 0, whose own nsub is MT here or the tile is the last one)
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          const int tap = kb / p.kb_per_tap;
          const int c0 = (kb - tap * p.kb_per_tap) * BK;
          const int r = tap / p.ksize;
          const int s = tap - r * p.ksize;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1, p.dbg, 0x100 | stage);
          const uint32_t full = kPair ? leader_addr(bar_full + 8 * stage) : (bar_full + 8 * stage);
          if (!kPair || rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, tx_bytes);
          else mbar_arrive_remote(bar_full + 8 * stage, 0);
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            if (t < nsub) {
              const uint32_t dst = smem_a + stage * Cfg::kABytes + t * Cfg::kASubBytes;
              if (p.a_im2col) {
                if (kPair) tma_load_im2col_4d_pair(dst, &tmap_a, full, c0, w0[t] - p.pad, h0[t] - p.pad, img[t], static_cast<uint16_t>(s), static_cast<uint16_t>(r));
                else tma_load_im2col_4d(dst, &tmap_a, full, c0, w0[t] - p.pad, h0[t] - p.pad, img[t], static_cast<uint16_t>(s), static_cast<uint16_t>(r));
              } else {
                if (kPair) tma_load_2d_pair(dst, &tmap_a, full, c0, m_cta + t * BM);
                else tma_load_2d(dst, &tmap_a, full, c0, m_cta + t * BM);
              }
            }
          }
          const int brow = n_tile * BN + static_cast<int>(rank) * Cfg::kBRows;
          if (p.skip & 2) { /* ablation: weights not fetched */ }
          else if (kPair) tma_load_2d_pair(smem_b + stage * Cfg::kBBytes, &tmap_b, full, tap * p.cin + c0, brow);
          else tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, full, tap * p.cin + c0, brow);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair: leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(kPair ? 2 * BM : BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = unit_id; tile < num_tiles; tile += num_units) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1, p.dbg, 0x200 | acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase, p.dbg, 0x300 | stage);
          tc_fence_after();
          const uint64_t bdesc = make_kmajor_desc<Cfg::kSwizzle>(smem_b + stage * Cfg::kBBytes);
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            if (p.skip & 4) break;
            const uint64_t adesc = make_kmajor_desc<Cfg::kSwizzle>(smem_a + stage * Cfg::kABytes + t * Cfg::kASubBytes);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 fp16 = 32 bytes inside the swizzled row: +2 in the 16-byte address field
              if (kPair) umma_f16_pair(d_tmem + t * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
              else umma_f16(d_tmem + t * BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            }
          }
          // free the smem slot (in both CTAs of a pair) once these MMAs retire
          if (kPair) umma_commit_pair(bar_empty + 8 * stage); else umma_commit(bar_empty + 8 * stage);
          if (kb == p.num_kb - 1) {
            if (kPair) umma_commit_pair(bar_tfull + 8 * acc); else umma_commit(bar_tfull + 8 * acc);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                    // TMEM lane quadrant this warp may read
    const int et = threadIdx.x - 64;           // 0..127
    int acc = 0;
    uint32_t acc_phase = 0;
    int buf = 0;
    for (int tile = unit_id; tile < num_tiles; tile += num_units) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int n0 = n_tile * BN;
      const int m_cta = m_tile * Cfg::kRowsPerTile + static_cast<int>(rank) * Cfg::kRowsPerCta;
      // stage this tile's per-channel scale/shift (double-buffered: a warp can be one tile ahead)
      float* sc = ep_scale + buf * BN;
      float* sh = ep_shift + buf * BN;
      buf ^= 1;
      for (int i = et; i < BN; i += kEpiThreads) {
        const int c = n0 + i;
        sc[i] = (c < p.cout) ? __ldg(p.scale + c) : 0.f;
        sh[i] = (c < p.cout) ? __ldg(p.shift + c) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(bar_tfull + 8 * acc, acc_phase, p.dbg, 0x400 | acc);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < MT; ++t) {
        if (m_cta + t * BM >= p.m_total) break;   // warp-uniform: subtile entirely past the end
        const int row = m_cta + t * BM + q * 32 + lane;
        const bool row_ok = row < p.m_total;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::kAccCols + t * BN;
        int img = 0, pix = 0;
        if (p.out_mode == 1) { img = row / p.hw; pix = row - img * p.hw; }
#pragma unroll 1
        for (int cc = 0; cc < BN / 32; ++cc) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + cc * 32, v);
          tmem_ld_wait();
          const int cbase = n0 + cc * 32;
          if (cbase >= p.cout) continue;           // warp-uniform
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(v[j]) * sc[cc * 32 + j] + sh[cc * 32 + j];
            f[j] = x > 0.f ? x : x * p.slope;
          }
          if (p.skip & 8) continue;
          if (p.out_mode == 0) {
            if (row_ok) {
              __half* dst = reinterpret_cast<__half*>(p.y) + static_cast<long long>(row) * p.y_ld + p.y_ch_off + cbase;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (cbase + g * 8 < p.cout) {
                  uint4 pk;
                  __half2 h0 = __floats2half2_rn(f[g * 8 + 0], f[g * 8 + 1]);
                  __half2 h1 = __floats2half2_rn(f[g * 8 + 2], f[g * 8 + 3]);
                  __half2 h2 = __floats2half2_rn(f[g * 8 + 4], f[g * 8 + 5]);
                  __half2 h3 = __floats2half2_rn(f[g * 8 + 6], f[g * 8 + 7]);
                  pk.x = *reinterpret_cast<uint32_t*>(&h0);
                  pk.y = *reinterpret_cast<uint32_t*>(&h1);
                  pk.z = *reinterpret_cast<uint32_t*>(&h2);
                  pk.w = *reinterpret_cast<uint32_t*>(&h3);
                  *reinterpret_cast<uint4*>(dst + g * 8) = pk;
                }
              }
            }
          } else {
            if (row_ok) {
              float* dst = reinterpret_cast<float*>(p.y) + (static_cast<long long>(img) * p.cout + cbase) * p.hw + pix;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (cbase + j < p.cout) dst[static_cast<long long>(j) * p.hw] = f[j];
              }
            }
          }
        }
      }  // M-subtiles
      // release this accumulator stage back to the MMA warp (of the leader CTA)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!kPair || rank == 0) mbar_arrive(bar_tempty + 8 * acc);
        else mbar_arrive_remote(bar_tempty + 8 * acc, 0);
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();   // pair: nobody leaves while the peer may still touch its smem / barriers
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side: tensor-map encoding through the driver entry points (no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encoders(EncodeTiledFn* tiled, EncodeIm2colFn* im2col) {
  static EncodeTiledFn s_tiled = nullptr;
  static EncodeIm2colFn s_im2col = nullptr;
  if (s_tiled == nullptr || s_im2col == nullptr) {
    void* f0 = nullptr;
    void* f1 = nullptr;
    cudaDriverEntryPointQueryResult q0, q1;
    YB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f0, cudaEnableDefault, &q0));
    YB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f1, cudaEnableDefault, &q1));
    if (f0 == nullptr || f1 == nullptr || q0 != cudaDriverEntryPointSuccess || q1 != cudaDriverEntryPointSuccess)
      return fail(YB_ERR_DRIVER, "cuTensorMapEncode* driver entry points unavailable");
    s_tiled = reinterpret_cast<EncodeTiledFn>(f0);
    s_im2col = reinterpret_cast<EncodeIm2colFn>(f1);
  }
  *tiled = s_tiled;
  *im2col = s_im2col;
  return 0;
}

int get_tensor_map_encoders(EncodeTiledFn* tiled, EncodeIm2colFn* im2col) { return get_encoders(tiled, im2col); }

template <int BN, int BK, int MT, bool kPair>
static int launch_conv(const CUtensorMap& ta, const CUtensorMap& tb, const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<BN, BK, MT, kPair>;
  static bool attr_set = false;
  if (!attr_set) {
    YB_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN, BK, MT, kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  if (kPair) {
    const int pairs = sm_count() / 2;
    cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3(tiles < sm_count() ? tiles : sm_count());
  }
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  YB_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BN, BK, MT, kPair>, ta, tb, p));
  return check_launch("conv_igemm_kernel");
}

template <int BK, bool kPair>
static int dispatch_conv(int bn, int mt, const CUtensorMap& ta, const CUtensorMap& tb, const ConvParams& p, cudaStream_t stream) {
  if (mt == 1) {
    if (bn == 64) return launch_conv<64, BK, 1, kPair>(ta, tb, p, stream);
    if (bn == 128) return launch_conv<128, BK, 1, kPair>(ta, tb, p, stream);
    return launch_conv<256, BK, 1, kPair>(ta, tb, p, stream);
  }
  if (bn == 64) return launch_conv<64, BK, 2, kPair>(ta, tb, p, stream);
  if (bn == 128) return launch_conv<128, BK, 2, kPair>(ta, tb, p, stream);
  return launch_conv<256, BK, 2, kPair>(ta, tb, p, stream);
}

int conv_igemm_forward(const void* x, const void* w, const float* scale, const float* shift, float slope, void* y, int batch,
                       int height, int width, int cin, int cout, int ksize, int x_ld, long long y_ld, int y_ch_off, int out_mode,
                       int flags, cudaStream_t stream) {
  YB_REQUIRE(x && w && scale && shift && y, "conv: null pointer");
  YB_REQUIRE(ksize == 1 || ksize == 3, "conv: ksize %d unsupported (1 or 3)", ksize);
  YB_REQUIRE(batch > 0 && height > 0 && width > 0, "conv: bad shape");
  YB_REQUIRE(cin % 32 == 0, "conv: Cin=%d must be a multiple of 32 (layer 0 uses yb_conv0_*)", cin);
  YB_REQUIRE(x_ld >= cin && x_ld % 8 == 0, "conv: x_ld=%d", x_ld);
  YB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0, "conv: x/w must be 16B aligned");
  YB_REQUIRE(out_mode == 0 || out_mode == 1, "conv: out_mode");
  if (out_mode == 0) {
    YB_REQUIRE(cout % 8 == 0 && y_ld % 8 == 0 && y_ch_off % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
               "conv: fp16 NHWC output needs Cout, y_ld, y_ch_off multiples of 8 and a 16B aligned pointer");
  }
  const long long m_total_ll = static_cast<long long>(batch) * height * width;
  YB_REQUIRE(m_total_ll < (1ll << 31) - BM, "conv: too many pixels");
  const int bk = (cin % 64 == 0) ? 64 : 32;
  // tile shape: flags may force BLOCK_N (bits 8..17) and the number of M-subtiles (bits 20..21);
  // otherwise pick the (BLOCK_N, M-subtiles) pair with the lowest modelled time.  The model was
  // fitted to tools/conv_sweep.py on B200 (profiles/r01_conv_sweep.md): the kernel is bound by the
  // L2->SM operand feed (~95 B/ns per SM, ~11 TB/s chip-wide), a tile costs its operand bytes at
  // that rate (or its MMA time if larger), tiles run in ceil(tiles/SMs) rounds, and a CTA tile whose
  // accumulator fills all of TMEM (256x256) cannot overlap its epilogue with the next mainloop.
  int bn = 0, mt = 0, pair = 0;
  const int force_bn = (flags >> 8) & 0x3FF;
  const int force_mt = (flags >> 20) & 0x3;
  const int force_pair = (flags >> 22) & 0x3;      // 0 = auto, 1 = single-CTA, 2 = CTA pair (cta_group::2)
  {
    double best = 1e300;
    const int sms = sm_count();
    const int num_kb = ksize * ksize * (cin / bk);
    for (int cpair = 0; cpair <= 1; ++cpair) {
      if (force_pair && cpair != force_pair - 1) continue;
      if (!force_pair && cpair == 1 && (flags & 4) == 0) continue;   // pairs are opt-in (YB_CONV_ALLOW_PAIR) until tuned
      for (int cbn = 64; cbn <= 256; cbn *= 2) {
        if (force_bn && cbn != force_bn) continue;
        if (!force_bn && cbn > 64 && cbn / 2 >= cout) continue;       // do not pad Cout by more than 2x
        for (int cmt = 1; cmt <= 2; ++cmt) {
          if (force_mt && cmt != force_mt) continue;
          const int rows_tile = BM * cmt * (cpair ? 2 : 1);
          const double tiles = static_cast<double>((m_total_ll + rows_tile - 1) / rows_tile) * ((cout + cbn - 1) / cbn);
          const int units = cpair ? sms / 2 : sms;
          const double rounds = static_cast<double>((static_cast<long long>(tiles) + units - 1) / units);
          const double bytes_kb = (cmt * BM + (cpair ? cbn / 2 : cbn)) * bk * 2.0;        // per CTA
          const double mma_ns_kb = cmt * (bk / 16) * (cbn >= 128 ? cbn / 2.0 : 64.0) / 1.9;   // cycles(N) = max(N,128)/2 @ ~1.9 GHz
          const double kb_ns = bytes_kb / 95.0 > mma_ns_kb ? bytes_kb / 95.0 : mma_ns_kb;
          const bool single_acc = 2 * cmt * cbn > 512;
          const double tile_ns = num_kb * kb_ns + (single_acc ? 8000.0 : 500.0);
          const double agg_ns = tiles * (cpair ? 2 : 1) * num_kb * bytes_kb / 11000.0;
          double t = rounds * tile_ns;
          if (agg_ns > t) t = agg_ns;
          // ties (e.g. 256x128 vs 128x256, same operand bytes) go to the wider-N shape, which measured ~8% faster
          if (t < best * 0.9999 || (t <= best * 1.0001 && cbn > bn)) { best = t; bn = cbn; mt = cmt; pair = cpair; }
        }
      }
    }
    if (bn == 0) { bn = force_bn ? force_bn : 128; mt = force_mt ? force_mt : 1; pair = force_pair == 2; }
  }
  YB_REQUIRE(bn == 64 || bn == 128 || bn == 256, "conv: BN=%d", bn);
  YB_REQUIRE(mt == 1 || mt == 2, "conv: MT=%d", mt);
  const int a_im2col = (ksize == 3) ? 1 : ((flags & 1) ? 0 : 1);

  EncodeTiledFn enc_tiled;
  EncodeIm2colFn enc_im2col;
  int rc = get_encoders(&enc_tiled, &enc_im2col);
  if (rc) return rc;

  ConvParams p;
  p.m_total = static_cast<int>(m_total_ll);
  p.height = height; p.width = width; p.cin = cin; p.cout = cout; p.ksize = ksize; p.pad = (ksize - 1) / 2;
  p.kb_per_tap = cin / bk;
  p.num_kb = ksize * ksize * p.kb_per_tap;
  const int rows_tile = BM * mt * (pair ? 2 : 1);
  p.m_tiles = (p.m_total + rows_tile - 1) / rows_tile;
  p.n_tiles = (cout + bn - 1) / bn;
  p.a_im2col = a_im2col;
  p.scale = scale; p.shift = shift; p.slope = slope;
  p.y = y; p.y_ld = y_ld; p.y_ch_off = y_ch_off; p.out_mode = out_mode;
  p.hw = height * width;
  p.dbg = debug_word_device();
  p.skip = (flags >> 24) & 0xF;

  const CUtensorMapSwizzle swz = (bk == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  alignas(64) CUtensorMap ta, tb;
  CUresult cr;
  if (a_im2col) {
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(cin), static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(height),
                                static_cast<cuuint64_t>(batch)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(x_ld) * 2, static_cast<cuuint64_t>(x_ld) * 2 * width,
                                   static_cast<cuuint64_t>(x_ld) * 2 * width * height};
    const int lower[2] = {-p.pad, -p.pad};                    // {W, H}
    const int upper[2] = {p.pad - (ksize - 1), p.pad - (ksize - 1)};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    cr = enc_im2col(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, lower, upper,
                    static_cast<cuuint32_t>(bk), BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(YB_ERR_DRIVER, "cuTensorMapEncodeIm2col failed (%d)", static_cast<int>(cr));
    // Driver workaround (same one CUTLASS carries, cute/atom/copy_traits_sm90_im2col.hpp): for
    // tensors smaller than 128 KiB drivers <= 13.1 set a descriptor bit that breaks im2col loads.
    int drv = 0;
    cudaDriverGetVersion(&drv);
    const unsigned long long span_bytes = static_cast<unsigned long long>(x_ld) * 2ull * width * height * batch;
    if (drv <= 13010 && span_bytes < 131072ull) reinterpret_cast<uint64_t*>(&ta)[1] &= ~(1ull << 21);
  } else {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cin), static_cast<cuuint64_t>(p.m_total)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(x_ld) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(bk), BM};
    const cuuint32_t estr[2] = {1, 1};
    cr = enc_tiled(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(YB_ERR_DRIVER, "cuTensorMapEncodeTiled(A) failed (%d)", static_cast<int>(cr));
  }
  {
    const cuuint64_t k_total = static_cast<cuuint64_t>(ksize) * ksize * cin;
    const cuuint64_t dims[2] = {k_total, static_cast<cuuint64_t>(cout)};
    const cuuint64_t strides[1] = {k_total * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(bk), static_cast<cuuint32_t>(pair ? bn / 2 : bn)};   // a pair CTA stages half the tile
    const cuuint32_t estr[2] = {1, 1};
    cr = enc_tiled(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(YB_ERR_DRIVER, "cuTensorMapEncodeTiled(W) failed (%d)", static_cast<int>(cr));
  }

  if (pair) {
    if (bk == 64) return dispatch_conv<64, true>(bn, mt, ta, tb, p, stream);
    return dispatch_conv<32, true>(bn, mt, ta, tb, p, stream);
  }
  if (bk == 64) return dispatch_conv<64, false>(bn, mt, ta, tb, p, stream);
  return dispatch_conv<32, false>(bn, mt, ta, tb, p, stream);
}

// ---------------------------------------------------------------------------------------------
// CUDA-core reference of the same unit: test / bisect utility only (never on the product path).
// One thread per (pixel, cout); fp16 inputs, fp32 accumulate, identical epilogue and outputs.
// ---------------------------------------------------------------------------------------------
__global__ void conv_ref_kernel(const __half* __restrict__ x, const __half* __restrict__ w, ConvParams p, int x_ld) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(p.m_total) * p.cout;
  if (idx >= total) return;
  const int co = static_cast<int>(idx % p.cout);
  const int row = static_cast<int>(idx / p.cout);
  const int img = row / p.hw;
  const int pix = row - img * p.hw;
  const int h = pix / p.width;
  const int wq = pix - h * p.width;
  float acc = 0.f;
  for (int r = 0; r < p.ksize; ++r) {
    const int hi = h + r - p.pad;
    if (hi < 0 || hi >= p.height) continue;
    for (int s = 0; s < p.ksize; ++s) {
      const int wi = wq + s - p.pad;
      if (wi < 0 || wi >= p.width) continue;
      const __half* xp = x + (static_cast<long long>(img) * p.hw + static_cast<long long>(hi) * p.width + wi) * x_ld;
      const __half* wp = w + (static_cast<long long>(co) * p.ksize * p.ksize + r * p.ksize + s) * p.cin;
      for (int c = 0; c < p.cin; ++c) acc += __half2float(xp[c]) * __half2float(wp[c]);
    }
  }
  float v = acc * p.scale[co] + p.shift[co];
  v = v > 0.f ? v : v * p.slope;
  if (p.out_mode == 0) {
    reinterpret_cast<__half*>(p.y)[static_cast<long long>(row) * p.y_ld + p.y_ch_off + co] = __float2half_rn(v);
  } else {
    reinterpret_cast<float*>(p.y)[(static_cast<long long>(img) * p.cout + co) * p.hw + pix] = v;
  }
}

int conv_ref_forward(const void* x, const void* w, const float* scale, const float* shift, float slope, void* y, int batch,
                     int height, int width, int cin, int cout, int ksize, int x_ld, long long y_ld, int y_ch_off, int out_mode,
                     cudaStream_t stream) {
  YB_REQUIRE(x && w && scale && shift && y, "conv_ref: null pointer");
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.m_total = batch * height * width;
  p.height = height; p.width = width; p.cin = cin; p.cout = cout; p.ksize = ksize; p.pad = (ksize - 1) / 2;
  p.scale = scale; p.shift = shift; p.slope = slope; p.y = y; p.y_ld = y_ld; p.y_ch_off = y_ch_off; p.out_mode = out_mode;
  p.hw = height * width;
  const long long total = static_cast<long long>(p.m_total) * cout;
  conv_ref_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const __half*>(x),
                                                                                  reinterpret_cast<const __half*>(w), p, x_ld);
  return check_launch("conv_ref_kernel");
}

}  // namespace yb
