// Weight-gradient GEMM on the 5th-gen tensor cores (sm_100a).
//
//   dW[tap][n][c] += sum over output positions p of  dz[p, n] * x[p + offset(tap), c]
//
// (autograd's convolution_backward wgrad for every conv on the path: reference
// mask_cyclegan_vc/train.py:241,298 backward() through model.py's Conv2d/Conv1d layers).
// Both operands are "MN-major" for the MMA: dz is stored [positions][N] and x is stored
// [positions][C], i.e. the contraction index (positions) is the OUTER smem dimension.  TMA fetches
// 64-position x 64-channel boxes (128-byte rows, 128B swizzle) of each operand; the UMMA
// descriptors address them with leading-byte-offset = one 64-channel chunk, stride-byte-offset =
// one 8-position group.  One CTA owns one (tap, 128-row n tile, cTile-column c tile, K split) and
// reduces its K slice into TMEM; the epilogue adds the tile into the fp32 gradient with
// red.global.add.v4.f32, which also merges the K splits and the repeated uses of one module in a
// backward pass.
#include "gemm_types.cuh"
#include "epilogue.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

#include <cstdlib>
#include <cuda_fp16.h>

namespace mcgvc {


constexpr int kWgBlockPos = 64;                          // positions per k-block
constexpr int kChunkBytes = kWgBlockPos * kBlockK * 2;   // one 64-pos x 64-ch box = 8 KB

template <int CTILE, int NPASS>
struct WgradCfg {
  static constexpr int kZBytes = 2 * kChunkBytes;               // 128 rows of n
  static constexpr int kXBytes = (CTILE / 64) * kChunkBytes;
  static constexpr int kStageBytes = (kZBytes + kXBytes) * (NPASS == 3 ? 2 : 1);
  static constexpr int kOutStageBytes = 4 * kStageFloatsPerWarp * 4;   // coalescing buffers of the 4 epilogue warps
  static constexpr int kStagesRaw = (222 * 1024 - kOutStageBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = CTILE < 32 ? 32 : CTILE;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kOutStageBytes;
  static_assert(kStages >= 2, "need at least two pipeline stages");
};


template <int CTILE, int NPASS>
__global__ void __launch_bounds__(256, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmZh, const __grid_constant__ CUtensorMap tmZl,
                const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                const __grid_constant__ WgradGeom g) {
  using Cfg = WgradCfg<CTILE, NPASS>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work decode: blockIdx.x -> (tap, nTile, cTile, split)
  const int cTiles = g.C / CTILE;
  const int nTiles = g.N / 128;
  int w = blockIdx.x;
  // taps fastest, position split slowest: the ~74 pairs running at a time work on the same slice of
  // positions for every tap, so dz / x are streamed from HBM once and re-read from L2 (split-fastest
  // order re-streamed both operands per tap: 2-3x the algorithmic DRAM bytes), and the red.adds of
  // one dW tile are spread over the whole launch
  const int t = w % g.nTaps;
  w /= g.nTaps;
  const int nt = w % nTiles;
  w /= nTiles;
  const int ct = w % cTiles;
  const int split = w / cTiles;
  const Tap tap = g.taps[t];
  const Tap ztap = g.ztaps[t];
  const int n0 = nt * 128, c0 = ct * CTILE;

  const int posTiles = g.tilesX * g.tilesY * g.tilesB;
  const int per = (posTiles + g.splitK - 1) / g.splitK;
  const int kBegin = split * per;
  const int kEnd = (kBegin + per < posTiles) ? kBegin + per : posTiles;
  const int numK = kEnd - kBegin;  // may be <= 0 for a trailing split

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmZh);
    ptx::prefetch_tmap(&tmXh);
    if (NPASS == 3) {
      ptx::prefetch_tmap(&tmZl);
      ptx::prefetch_tmap(&tmXl);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(tfull, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (numK > 0) {
    if (warp == 0 && lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = kBegin; kt < kEnd; ++kt) {
        int m = kt;
        const int tx = m % g.tilesX;
        m /= g.tilesX;
        const int ty = m % g.tilesY;
        const int tb = m / g.tilesY;
        const int x0 = tx * g.BX, y0 = ty * g.BY, b0 = tb * g.BB;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kStageBytes;
        ptx::mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          ptx::tma_load_5d(st + j * kChunkBytes, &tmZh, &full[stage], n0 + j * 64, x0 + ztap.dx,
                           y0 + ztap.dy, 0, b0);
#pragma unroll
        for (int j = 0; j < CTILE / 64; ++j)
          ptx::tma_load_5d(st + Cfg::kZBytes + j * kChunkBytes, &tmXh, &full[stage], c0 + j * 64,
                           x0 + tap.dx, y0 + tap.dy, tap.plane, b0);
        if (NPASS == 3) {
          uint8_t* lo = st + Cfg::kZBytes + Cfg::kXBytes;
#pragma unroll
          for (int j = 0; j < 2; ++j)
            ptx::tma_load_5d(lo + j * kChunkBytes, &tmZl, &full[stage], n0 + j * 64, x0 + ztap.dx,
                             y0 + ztap.dy, 0, b0);
#pragma unroll
          for (int j = 0; j < CTILE / 64; ++j)
            ptx::tma_load_5d(lo + Cfg::kZBytes + j * kChunkBytes, &tmXl, &full[stage],
                             c0 + j * 64, x0 + tap.dx, y0 + tap.dy, tap.plane, b0);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1 && lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      const uint32_t idesc = ptx::umma_idesc_bf16(128, CTILE, 1, 1) & ~((NPASS == 1 && g.half16) ? ((1u << 7) | (1u << 10)) : 0u);
      constexpr uint32_t kLbo = kChunkBytes;  // next 64-channel chunk
      constexpr uint32_t kSbo = 1024;         // next 8-position group
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < numK; ++kb) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sZ = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t sX = sZ + Cfg::kZBytes;
        const uint32_t sZl = sX + Cfg::kXBytes;
        const uint32_t sXl = sZl + Cfg::kZBytes;
#pragma unroll
        for (int k = 0; k < kWgBlockPos / 16; ++k) {
          const uint64_t dZh = ptx::umma_smem_desc_sw128(sZ + k * 2048, kLbo, kSbo);
          const uint64_t dXh = ptx::umma_smem_desc_sw128(sX + k * 2048, kLbo, kSbo);
          ptx::umma_bf16(tmem_base, dZh, dXh, idesc, (kb | k) != 0);
          if (NPASS == 3) {
            const uint64_t dZl = ptx::umma_smem_desc_sw128(sZl + k * 2048, kLbo, kSbo);
            const uint64_t dXl = ptx::umma_smem_desc_sw128(sXl + k * 2048, kLbo, kSbo);
            ptx::umma_bf16(tmem_base, dZh, dXl, idesc, 1);
            ptx::umma_bf16(tmem_base, dZl, dXh, idesc, 1);
          }
        }
        ptx::umma_commit(&empty[stage]);
        if (kb == numK - 1) ptx::umma_commit(tfull);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (warp >= 4) {
      // ---------------------------------------------------------------- epilogue
      const int quad = warp & 3;
      const int n = n0 + quad * 32 + lane;
      float* drow = g.dw + ((long long)tap.w * g.N + n) * g.C + c0;
      float osc = 1.f;      // C8H: single fp16 pass over C8 operand planes, scaled by their records
      if (NPASS == 1 && g.half16) {
        osc = g.c8OutScale;
        if (g.c8RecZ && g.c8RecX) osc *= __ldg(g.c8RecZ) * __ldg(g.c8RecX);
      }
      // coalesced red.add through the warp's staging buffer (epilogue.cuh red_chunk_staged)
      float* rowp[8];
      {
        const unsigned long long mine = reinterpret_cast<unsigned long long>(drow);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          rowp[i] = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, mine, (i >> 2) * 16 + (i & 3) * 4 + (lane >> 3)));
      }
      float* sbuf = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes + 256) + (warp - 4) * kStageFloatsPerWarp;
      ptx::mbar_wait(tfull, 0);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
      for (int j = 0; j < CTILE / 32; ++j) {
        uint32_t v[32];
        ptx::tmem_ld32(taddr + j * 32, v);
        ptx::tmem_ld_wait();
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = osc * __uint_as_float(v[i]);
        float* rp[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rp[i] = rowp[i] + j * 32;
        red_chunk_staged(o, sbuf, rp, lane);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

static bool check_wgrad_geom(const WgradGeom& g) {
  if (g.BX * g.BY * g.BB != kWgBlockPos) { set_error("wgrad: box %dx%dx%d != 64", g.BX, g.BY, g.BB); return false; }
  if (g.N % 128 || g.dz.C != g.N) { set_error("wgrad: N=%d (dz.C=%d) must be a multiple of 128", g.N, g.dz.C); return false; }
  if (g.C % 64 || g.x.C != g.C) { set_error("wgrad: C=%d (x.C=%d) must be a multiple of 64", g.C, g.x.C); return false; }
  if (g.cTile != 64 && g.cTile != 128 && g.cTile != 256) { set_error("wgrad: cTile=%d", g.cTile); return false; }
  if (g.C % g.cTile) { set_error("wgrad: C=%d %% cTile=%d", g.C, g.cTile); return false; }
  if (g.nTaps < 1 || g.nTaps > kMaxTaps || g.splitK < 1) { set_error("wgrad: taps/splitK"); return false; }
  if (g.nPass != 1 && g.nPass != 3) { set_error("wgrad: nPass=%d", g.nPass); return false; }
  if (g.half16 && g.nPass != 1) { set_error("wgrad: half16 needs nPass = 1"); return false; }
  if (g.tapPair) {
    if (g.nPass != 1 || g.N % 256 || g.cTile < 128) { set_error("wgrad: tap pairing needs the single-pass pair kernel"); return false; }
    for (int i = 0; i < g.nTaps; ++i)
      if (g.ztaps[i].dx || g.ztaps[i].dy) { set_error("wgrad: tap pairing needs an unshifted dz operand"); return false; }
  }
  return true;
}

template <int CTILE, int NPASS>
static cudaError_t launch_wgrad_tc_t(const WgradGeom& g, cudaStream_t stream) {
  using Cfg = WgradCfg<CTILE, NPASS>;
  CUtensorMap tmZh, tmZl, tmXh, tmXl;
  if (!make_act_tmap(&tmZh, g.dz.hi, g.dz, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
  if (!make_act_tmap(&tmXh, g.x.hi, g.x, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
  if (NPASS == 3) {
    if (!make_act_tmap(&tmZl, g.dz.lo, g.dz, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
    if (!make_act_tmap(&tmXl, g.x.lo, g.x, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
  } else {
    tmZl = tmZh;
    tmXl = tmXh;
  }
  static bool attr_done[64] = {};   // cudaFuncSetAttribute is per device
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  bool& attr_set = attr_done[dev_id & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<CTILE, NPASS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("wgrad: smem attr: %s", cudaGetErrorString(e)); return e; }
    attr_set = true;
  }
  const long long grid = (long long)g.nTaps * (g.N / 128) * (g.C / CTILE) * g.splitK;
  profile_begin(1, g.algoFlops, stream);
  wgrad_tc_kernel<CTILE, NPASS><<<(unsigned)grid, 256, Cfg::kSmemBytes, stream>>>(tmZh, tmZl, tmXh,
                                                                                 tmXl, g);
  profile_end(stream);
  return launched();
}


// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2, M = 256 gradient rows per pair).  Each CTA stages its own 128
// rows of dz and HALF of the x channels of the tile, halving the operand bytes per SM per MMA.
template <int CTILE, int NPASS>
struct Wgrad2Cfg {
  static constexpr int kZBytes = 2 * kChunkBytes;                   // this CTA's 128 n rows
  static constexpr int kXBytes = (CTILE / 128) * kChunkBytes;       // this CTA's CTILE/2 channels
  static constexpr int kStageBytes = (kZBytes + kXBytes) * (NPASS == 3 ? 2 : 1);
  static constexpr int kOutStageBytes = 4 * kStageFloatsPerWarp * 4;   // coalescing buffers of the 4 epilogue warps
  static constexpr int kStagesRaw = (222 * 1024 - kOutStageBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = CTILE;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kOutStageBytes;
  static_assert(CTILE == 128 || CTILE == 256, "pair tile is 128 or 256 channels wide");
};

template <int CTILE, int NPASS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
wgrad_tc2_kernel(const __grid_constant__ CUtensorMap tmZh, const __grid_constant__ CUtensorMap tmZl,
                 const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                 const __grid_constant__ WgradGeom g) {
  using Cfg = Wgrad2Cfg<CTILE, NPASS>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;

  const int cTiles = g.C / CTILE;
  const int nTiles = g.N / 256;
  int w = blockIdx.x >> 1;
  // taps fastest, position split slowest: the ~74 pairs running at a time work on the same slice of
  // positions for every tap, so dz / x are streamed from HBM once and re-read from L2 (split-fastest
  // order re-streamed both operands per tap: 2-3x the algorithmic DRAM bytes), and the red.adds of
  // one dW tile are spread over the whole launch
  const int t = w % g.nTaps;
  w /= g.nTaps;
  const int nt = w % nTiles;
  w /= nTiles;
  const int ct = w % cTiles;
  const int split = w / cTiles;
  const Tap tap = g.taps[t];
  const Tap ztap = g.ztaps[t];
  const int n0 = nt * 256 + (int)rank * 128;            // this CTA's gradient rows
  const int c0 = ct * CTILE;                            // pair's channel tile
  const int cLoad = c0 + (int)rank * (CTILE / 2);       // this CTA's half of the x channels

  const int posTiles = g.tilesX * g.tilesY * g.tilesB;
  const int per = (posTiles + g.splitK - 1) / g.splitK;
  const int kBegin = split * per;
  const int kEnd = (kBegin + per < posTiles) ? kBegin + per : posTiles;
  const int numK = kEnd - kBegin;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmZh);
    ptx::prefetch_tmap(&tmXh);
    if (NPASS == 3) {
      ptx::prefetch_tmap(&tmZl);
      ptx::prefetch_tmap(&tmXl);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full[i], 2);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(tfull, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (numK > 0) {
    if (warp == 0 && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = kBegin; kt < kEnd; ++kt) {
        int m = kt;
        const int tx = m % g.tilesX;
        m /= g.tilesX;
        const int ty = m % g.tilesY;
        const int tb = m / g.tilesY;
        const int x0 = tx * g.BX, y0 = ty * g.BY, b0 = tb * g.BB;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kStageBytes;
        if (leader) ptx::mbar_arrive_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          ptx::tma_load_5d_2sm(st + j * kChunkBytes, &tmZh, &full[stage], n0 + j * 64, x0 + ztap.dx,
                               y0 + ztap.dy, 0, b0);
#pragma unroll
        for (int j = 0; j < CTILE / 128; ++j)
          ptx::tma_load_5d_2sm(st + Cfg::kZBytes + j * kChunkBytes, &tmXh, &full[stage],
                               cLoad + j * 64, x0 + tap.dx, y0 + tap.dy, tap.plane, b0);
        if (NPASS == 3) {
          uint8_t* lo = st + Cfg::kZBytes + Cfg::kXBytes;
#pragma unroll
          for (int j = 0; j < 2; ++j)
            ptx::tma_load_5d_2sm(lo + j * kChunkBytes, &tmZl, &full[stage], n0 + j * 64,
                                 x0 + ztap.dx, y0 + ztap.dy, 0, b0);
#pragma unroll
          for (int j = 0; j < CTILE / 128; ++j)
            ptx::tma_load_5d_2sm(lo + Cfg::kZBytes + j * kChunkBytes, &tmXl, &full[stage],
                                 cLoad + j * 64, x0 + tap.dx, y0 + tap.dy, tap.plane, b0);
        }
        if (!leader) ptx::mbar_arrive_remote(&full[stage], 0);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1 && lane == 0 && leader) {
      const uint32_t idesc = ptx::umma_idesc_bf16(256, CTILE, 1, 1) & ~((NPASS == 1 && g.half16) ? ((1u << 7) | (1u << 10)) : 0u);
      constexpr uint32_t kLbo = kChunkBytes;
      constexpr uint32_t kSbo = 1024;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < numK; ++kb) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sZ = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t sX = sZ + Cfg::kZBytes;
        const uint32_t sZl = sX + Cfg::kXBytes;
        const uint32_t sXl = sZl + Cfg::kZBytes;
#pragma unroll
        for (int k = 0; k < kWgBlockPos / 16; ++k) {
          const uint64_t dZh = ptx::umma_smem_desc_sw128(sZ + k * 2048, kLbo, kSbo);
          const uint64_t dXh = ptx::umma_smem_desc_sw128(sX + k * 2048, kLbo, kSbo);
          ptx::umma_bf16_2cta(tmem_base, dZh, dXh, idesc, (kb | k) != 0);
          if (NPASS == 3) {
            const uint64_t dZl = ptx::umma_smem_desc_sw128(sZl + k * 2048, kLbo, kSbo);
            const uint64_t dXl = ptx::umma_smem_desc_sw128(sXl + k * 2048, kLbo, kSbo);
            ptx::umma_bf16_2cta(tmem_base, dZh, dXl, idesc, 1);
            ptx::umma_bf16_2cta(tmem_base, dZl, dXh, idesc, 1);
          }
        }
        ptx::umma_commit_2cta(&empty[stage], 0x3);
        if (kb == numK - 1) ptx::umma_commit_2cta(tfull, 0x3);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (warp >= 4) {
      const int quad = warp & 3;
      const int n = n0 + quad * 32 + lane;
      float* drow = g.dw + ((long long)tap.w * g.N + n) * g.C + c0;
      float osc = 1.f;      // C8H: single fp16 pass over C8 operand planes, scaled by their records
      if (NPASS == 1 && g.half16) {
        osc = g.c8OutScale;
        if (g.c8RecZ && g.c8RecX) osc *= __ldg(g.c8RecZ) * __ldg(g.c8RecX);
      }
      // coalesced red.add through the warp's staging buffer (epilogue.cuh red_chunk_staged)
      float* rowp[8];
      {
        const unsigned long long mine = reinterpret_cast<unsigned long long>(drow);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          rowp[i] = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, mine, (i >> 2) * 16 + (i & 3) * 4 + (lane >> 3)));
      }
      float* sbuf = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes + 256) + (warp - 4) * kStageFloatsPerWarp;
      ptx::mbar_wait(tfull, 0);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
      for (int j = 0; j < CTILE / 32; ++j) {
        uint32_t v[32];
        ptx::tmem_ld32(taddr + j * 32, v);
        ptx::tmem_ld_wait();
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = osc * __uint_as_float(v[i]);
        float* rp[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rp[i] = rowp[i] + j * 32;
        red_chunk_staged(o, sbuf, rp, lane);
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 2) ptx::tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
}

template <int CTILE, int NPASS>
static cudaError_t launch_wgrad_tc2_t(const WgradGeom& g, cudaStream_t stream) {
  using Cfg = Wgrad2Cfg<CTILE, NPASS>;
  CUtensorMap tmZh, tmZl, tmXh, tmXl;
  if (!make_act_tmap(&tmZh, g.dz.hi, g.dz, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
  if (!make_act_tmap(&tmXh, g.x.hi, g.x, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
  if (NPASS == 3) {
    if (!make_act_tmap(&tmZl, g.dz.lo, g.dz, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
    if (!make_act_tmap(&tmXl, g.x.lo, g.x, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
  } else {
    tmZl = tmZh;
    tmXl = tmXh;
  }
  static bool attr_done[64] = {};   // cudaFuncSetAttribute is per device
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  bool& attr_set = attr_done[dev_id & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc2_kernel<CTILE, NPASS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("wgrad2: smem attr: %s", cudaGetErrorString(e)); return e; }
    attr_set = true;
  }
  const long long pairs = (long long)g.nTaps * (g.N / 256) * (g.C / CTILE) * g.splitK;
  profile_begin(1, g.algoFlops, stream);
  wgrad_tc2_kernel<CTILE, NPASS><<<(unsigned)(2 * pairs), 256, Cfg::kSmemBytes, stream>>>(tmZh, tmZl, tmXh, tmXl, g);
  profile_end(stream);
  return launched();
}

// ------------------------------------------------------------------------------------------------
// Tap-paired single-pass variant (WgradGeom::tapPair): same CTA-pair decomposition, but one work item
// owns TWO filter taps of its (256 gradient rows x CTILE channels) tile.  The dz k-block is staged once
// and multiplied with both taps' x boxes into two TMEM accumulators (2 x CTILE <= 512 columns).  A
// single bf16 / fp16 pass needs 32 KB of operands per 4 MMAs in the unpaired kernel (116 GB/s per SM at
// the MMA rate, more than an SM can ingest); pairing makes it 48 KB per 8 MMAs.
template <int CTILE>
struct WgradPCfg {
  static constexpr int kZBytes = 2 * kChunkBytes;                   // this CTA's 128 n rows
  static constexpr int kXBytes = (CTILE / 128) * kChunkBytes;       // this CTA's CTILE/2 channels, per tap
  static constexpr int kStageBytes = kZBytes + 2 * kXBytes;         // 48 KB (CTILE 256) / 32 KB (CTILE 128)
  static constexpr int kOutStageBytes = 4 * kStageFloatsPerWarp * 4;   // coalescing buffers of the 4 epilogue warps
  static constexpr int kStagesRaw = (222 * 1024 - kOutStageBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = 2 * CTILE;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kOutStageBytes;
  static_assert(CTILE == 128 || CTILE == 256, "pair tile is 128 or 256 channels wide");
};

template <int CTILE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
wgrad_tc2p_kernel(const __grid_constant__ CUtensorMap tmZh, const __grid_constant__ CUtensorMap tmXh,
                  const __grid_constant__ WgradGeom g) {
  using Cfg = WgradPCfg<CTILE>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;

  const int cTiles = g.C / CTILE;
  const int nTiles = g.N / 256;
  int w = blockIdx.x >> 1;
  const int tapPairs = (g.nTaps + 1) >> 1;               // tap pairs fastest, position split slowest (see wgrad_tc2_kernel)
  const int tp = w % tapPairs;
  w /= tapPairs;
  const int nt = w % nTiles;
  w /= nTiles;
  const int ct = w % cTiles;
  const int split = w / cTiles;
  const int t0 = 2 * tp, t1 = 2 * tp + 1;
  const bool two = t1 < g.nTaps;
  const Tap tapA = g.taps[t0];
  const Tap tapB = g.taps[two ? t1 : t0];
  const int n0 = nt * 256 + (int)rank * 128;
  const int c0 = ct * CTILE;
  const int cLoad = c0 + (int)rank * (CTILE / 2);

  const int posTiles = g.tilesX * g.tilesY * g.tilesB;
  const int per = (posTiles + g.splitK - 1) / g.splitK;
  const int kBegin = split * per;
  const int kEnd = (kBegin + per < posTiles) ? kBegin + per : posTiles;
  const int numK = kEnd - kBegin;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmZh);
    ptx::prefetch_tmap(&tmXh);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full[i], 2);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(tfull, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2cta(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (numK > 0) {
    if (warp == 0 && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // a lone last tap still stages (and multiplies) its x box twice into the second accumulator,
      // which the epilogue ignores: one code path, at most one wasted tap in 25
      for (int kt = kBegin; kt < kEnd; ++kt) {
        int m = kt;
        const int tx = m % g.tilesX;
        m /= g.tilesX;
        const int ty = m % g.tilesY;
        const int tb = m / g.tilesY;
        const int x0 = tx * g.BX, y0 = ty * g.BY, b0 = tb * g.BB;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kStageBytes;
        if (leader) ptx::mbar_arrive_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          ptx::tma_load_5d_2sm(st + j * kChunkBytes, &tmZh, &full[stage], n0 + j * 64, x0, y0, 0, b0);
#pragma unroll
        for (int j = 0; j < CTILE / 128; ++j) {
          ptx::tma_load_5d_2sm(st + Cfg::kZBytes + j * kChunkBytes, &tmXh, &full[stage], cLoad + j * 64,
                               x0 + tapA.dx, y0 + tapA.dy, tapA.plane, b0);
          ptx::tma_load_5d_2sm(st + Cfg::kZBytes + Cfg::kXBytes + j * kChunkBytes, &tmXh, &full[stage], cLoad + j * 64,
                               x0 + tapB.dx, y0 + tapB.dy, tapB.plane, b0);
        }
        if (!leader) ptx::mbar_arrive_remote(&full[stage], 0);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1 && lane == 0 && leader) {
      const uint32_t idesc = ptx::umma_idesc_bf16(256, CTILE, 1, 1) & ~(g.half16 ? ((1u << 7) | (1u << 10)) : 0u);
      constexpr uint32_t kLbo = kChunkBytes;
      constexpr uint32_t kSbo = 1024;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < numK; ++kb) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sZ = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t sXa = sZ + Cfg::kZBytes;
        const uint32_t sXb = sXa + Cfg::kXBytes;
#pragma unroll
        for (int k = 0; k < kWgBlockPos / 16; ++k) {
          const uint64_t dZ = ptx::umma_smem_desc_sw128(sZ + k * 2048, kLbo, kSbo);
          ptx::umma_bf16_2cta(tmem_base, dZ, ptx::umma_smem_desc_sw128(sXa + k * 2048, kLbo, kSbo), idesc, (kb | k) != 0);
          ptx::umma_bf16_2cta(tmem_base + CTILE, dZ, ptx::umma_smem_desc_sw128(sXb + k * 2048, kLbo, kSbo), idesc, (kb | k) != 0);
        }
        ptx::umma_commit_2cta(&empty[stage], 0x3);
        if (kb == numK - 1) ptx::umma_commit_2cta(tfull, 0x3);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (warp >= 4) {
      const int quad = warp & 3;
      const int n = n0 + quad * 32 + lane;
      float osc = 1.f;
      if (g.half16) {
        osc = g.c8OutScale;
        if (g.c8RecZ && g.c8RecX) osc *= __ldg(g.c8RecZ) * __ldg(g.c8RecX);
      }
      ptx::mbar_wait(tfull, 0);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
      for (int which = 0; which < (two ? 2 : 1); ++which) {
        const Tap tap = which ? tapB : tapA;
        float* drow = g.dw + ((long long)tap.w * g.N + n) * g.C + c0;
        float* rowp[8];
        {
          const unsigned long long mine = reinterpret_cast<unsigned long long>(drow);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            rowp[i] = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, mine, (i >> 2) * 16 + (i & 3) * 4 + (lane >> 3)));
        }
        float* sbuf = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes + 256) + (warp - 4) * kStageFloatsPerWarp;
#pragma unroll 1
        for (int j = 0; j < CTILE / 32; ++j) {
          uint32_t v[32];
          ptx::tmem_ld32(taddr + which * CTILE + j * 32, v);
          ptx::tmem_ld_wait();
          float o[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = osc * __uint_as_float(v[i]);
          float* rp[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) rp[i] = rowp[i] + j * 32;
          red_chunk_staged(o, sbuf, rp, lane);
        }
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 2) ptx::tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
}

template <int CTILE>
static cudaError_t launch_wgrad_tc2p_t(const WgradGeom& g, cudaStream_t stream) {
  using Cfg = WgradPCfg<CTILE>;
  CUtensorMap tmZh, tmXh;
  if (!make_act_tmap(&tmZh, g.dz.hi, g.dz, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
  if (!make_act_tmap(&tmXh, g.x.hi, g.x, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
  static bool attr_done[64] = {};
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  bool& attr_set = attr_done[dev_id & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc2p_kernel<CTILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("wgrad2p: smem attr: %s", cudaGetErrorString(e)); return e; }
    attr_set = true;
  }
  const long long pairs = (long long)((g.nTaps + 1) / 2) * (g.N / 256) * (g.C / CTILE) * g.splitK;
  profile_begin(1, g.algoFlops, stream);
  wgrad_tc2p_kernel<CTILE><<<(unsigned)(2 * pairs), 256, Cfg::kSmemBytes, stream>>>(tmZh, tmXh, g);
  profile_end(stream);
  return launched();
}

static int env_wg_cta2() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCGVC_WGRAD_CTA2");   // 0 = single-CTA kernel everywhere (A/B measurements)
    v = e ? atoi(e) : 1;
  }
  return v;
}
static int g_force_wg_cta2 = -1;
void set_force_wgrad_cta2(int v) { g_force_wg_cta2 = v; }

cudaError_t launch_wgrad_tc(const WgradGeom& g, cudaStream_t stream) {
  if (!check_wgrad_geom(g)) return cudaErrorInvalidValue;
  {
    const int want = g_force_wg_cta2 >= 0 ? g_force_wg_cta2 : env_wg_cta2();
    // pair tile: 256 gradient rows x (128 | 256) channels; cTile (64/128/256) keeps its meaning
    // of "channels per work item", so the pair kernel needs cTile >= 128
    if (want && g.N % 256 == 0 && g.cTile >= 128) {
      if (g.tapPair && g.nPass == 1) return g.cTile == 256 ? launch_wgrad_tc2p_t<256>(g, stream) : launch_wgrad_tc2p_t<128>(g, stream);
      if (g.nPass == 3) return g.cTile == 256 ? launch_wgrad_tc2_t<256, 3>(g, stream) : launch_wgrad_tc2_t<128, 3>(g, stream);
      return g.cTile == 256 ? launch_wgrad_tc2_t<256, 1>(g, stream) : launch_wgrad_tc2_t<128, 1>(g, stream);
    }
  }
  if (g.nPass == 3) {
    if (g.cTile == 256) return launch_wgrad_tc_t<256, 3>(g, stream);
    if (g.cTile == 128) return launch_wgrad_tc_t<128, 3>(g, stream);
    return launch_wgrad_tc_t<64, 3>(g, stream);
  }
  if (g.cTile == 256) return launch_wgrad_tc_t<256, 1>(g, stream);
  if (g.cTile == 128) return launch_wgrad_tc_t<128, 1>(g, stream);
  return launch_wgrad_tc_t<64, 1>(g, stream);
}

// ------------------------------------------------------------------------------------------------
// SIMT checking kernel (see conv_igemm.cu): one thread per (tap, n, c), serial over positions.
__device__ __forceinline__ float bf16_bits(uint16_t v) {
  return __uint_as_float(static_cast<uint32_t>(v) << 16);
}

__global__ void wgrad_simt_kernel(const __grid_constant__ WgradGeom g) {
  const long long total = (long long)g.nTaps * g.N * g.C;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % g.C);
  const int n = (int)((idx / g.C) % g.N);
  const int t = (int)(idx / ((long long)g.C * g.N));
  const Tap tap = g.taps[t];
  const Tap ztap = g.ztaps[t];
  const uint16_t* Zh = reinterpret_cast<const uint16_t*>(g.dz.hi);
  const uint16_t* Zl = reinterpret_cast<const uint16_t*>(g.dz.lo);
  const uint16_t* Xh = reinterpret_cast<const uint16_t*>(g.x.hi);
  const uint16_t* Xl = reinterpret_cast<const uint16_t*>(g.x.lo);
  float acc = 0.f;
  for (int b = 0; b < g.pB; ++b) {
    if (b >= g.dz.B || b >= g.x.B) continue;
    for (int y = 0; y < g.pY; ++y) {
      const int zy = y + ztap.dy, xy = y + tap.dy;
      if (zy < 0 || zy >= g.dz.Y || xy < 0 || xy >= g.x.Y) continue;
      for (int x = 0; x < g.pX; ++x) {
        const int zx = x + ztap.dx, xx = x + tap.dx;
        if (zx < 0 || zx >= g.dz.X || xx < 0 || xx >= g.x.X) continue;
        const long long zo = ((((long long)b * g.dz.P) * g.dz.Y + zy) * g.dz.X + zx) * g.dz.C + n;
        const long long xo =
            ((((long long)b * g.x.P + tap.plane) * g.x.Y + xy) * g.x.X + xx) * g.x.C + c;
        const float zh = g.half16 ? __half2float(__ushort_as_half(Zh[zo])) : bf16_bits(Zh[zo]);
        const float xh = g.half16 ? __half2float(__ushort_as_half(Xh[xo])) : bf16_bits(Xh[xo]);
        acc = fmaf(zh, xh, acc);
        if (g.nPass == 3) {
          acc = fmaf(zh, bf16_bits(Xl[xo]), acc);
          acc = fmaf(bf16_bits(Zl[zo]), xh, acc);
        }
      }
    }
  }
  if (g.half16) {
    float sc = g.c8OutScale;
    if (g.c8RecZ && g.c8RecX) sc *= g.c8RecZ[0] * g.c8RecX[0];
    acc *= sc;
  }
  atomicAdd(g.dw + ((long long)tap.w * g.N + n) * g.C + c, acc);
}

cudaError_t launch_wgrad_simt(const WgradGeom& g, cudaStream_t stream) {
  if (!check_wgrad_geom(g)) return cudaErrorInvalidValue;
  const long long total = (long long)g.nTaps * g.N * g.C;
  const int threads = 128;
  wgrad_simt_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(g);
  return launched();
}

}  // namespace mcgvc
