// Implicit-GEMM convolution, "16-bit main pass + 8-bit correction passes" precision scheme (C8).
//
// Same job, tiling and TMA/mbarrier pipeline as conv_tc2_kernel in conv_igemm.cu (CTA pair,
// cta_group::2, M = 256), but every operand value v is held as THREE planes
//     hi = fp16(v)  (or bf16)          h8 = e4m3(hi * 2^u)          l8 = e4m3((v - hi) * 2^(u+11))
// and the product is evaluated as
//     A*W ~= Ah*Wh  +  2^-(uA+uW+11) * (A8h*W8l + A8l*W8h)
// i.e. ONE 16-bit MMA (kind::f16) plus TWO e4m3 MMAs (kind::f8f6f4, twice the MAC rate): 2 tensor
// core time units per MAC instead of the 3 of the split-bf16 scheme, at the same operand bytes
// (4 B per value).  The residual v - hi is <= 2^-11 |v| for fp16, so e4m3's 4 significant bits on
// the correction terms leave an error of ~2^-17 per product -- the same class as split-bf16.
// The main products accumulate in one TMEM accumulator (D1), the scaled corrections in a second
// (D2); the epilogue merges them:  out = c1 * (D1 + c2 * D2) + bias + residual.
//
// Selected with mcgvc_set_precision(MCGVC_PRECISION_C8) for every layer but the stems, heads and the
// 1-D trunk (network.cu); also reachable in isolation through mcgvc_debug_conv_c8
// (tests/kernel_check.py, tools/layer_bench.py).  256-wide tiles use all 512 TMEM columns for D1 + D2,
// so their epilogue is not overlapped with the next tile's MMAs; a two-phase variant (all correction
// MMAs first, D2 parked as bf16 in shared memory, then the main pass) that restores the overlap was
// built and measured slower in isolation (up2: 787 us vs 759 us), so it is not kept.
#include "epilogue.cuh"
#include "gemm_types.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace mcgvc {

namespace {

// UMMA shared-memory descriptor, K-major, 64-byte swizzle: 8-row x 64-byte atoms, 512 B apart.
__device__ __forceinline__ uint64_t umma_smem_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((512u >> 4) & 0x3FFFu) << 32;   // stride byte offset
  d |= 1ull << 46;                                            // descriptor version
  d |= 4ull << 61;                                            // SWIZZLE_64B
  return d;
}
// kind::f16 instruction descriptor with a selectable 16-bit input format (0 = fp16, 1 = bf16)
__host__ __device__ constexpr uint32_t umma_idesc_16(uint32_t fmt, uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

template <int BLOCK_N>
struct C8Cfg {
  static constexpr int kA16 = kTileM * kBlockK * 2;           // 16 KB: this CTA's 128 positions, 16-bit
  static constexpr int kW16 = (BLOCK_N / 2) * kBlockK * 2;    // this CTA's half of the weight rows
  static constexpr int kA8 = kTileM * kBlockK;                // 8 KB per 8-bit plane
  static constexpr int kW8 = (BLOCK_N / 2) * kBlockK;
  static constexpr int kOffW16 = kA16;
  static constexpr int kOffA8h = kOffW16 + kW16;
  static constexpr int kOffA8l = kOffA8h + kA8;
  static constexpr int kOffW8h = kOffA8l + kA8;
  static constexpr int kOffW8l = kOffW8h + kW8;
  static constexpr int kStageBytes = kOffW8l + kW8;           // 64 KB (N = 256) / 48 KB (N = 128)
  static constexpr int kOutStageBytes = 8 * kStageFloatsPerWarp * 4;   // coalescing buffers of the 8 epilogue warps
  static constexpr int kStages = (222 * 1024 - kOutStageBytes) / kStageBytes > 8 ? 8 : (222 * 1024 - kOutStageBytes) / kStageBytes;
  static constexpr int kAccBufs = 512 / (2 * BLOCK_N);        // D1 + D2 per buffer: 1 (N=256) or 2 (N=128)
  static constexpr int kTmemCols = 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kOutStageBytes;
  static_assert(kStageBytes % 1024 == 0 && kOffA8h % 1024 == 0 && kOffW8h % 1024 == 0 && kOffW8l % 1024 == 0, "tile alignment");
  static_assert(kAccBufs >= 1 && kStages >= 2, "config");
};

// 384 threads: warps 0-3 = TMA producer / MMA issuer / TMEM allocator / spare, warps 4-11 = EIGHT epilogue
// warps, two per TMEM lane quadrant, each draining half of the tile's columns: with D1 + D2 filling all of
// TMEM the 256-wide tile cannot overlap its epilogue with the next tile's MMAs, so its duration sits on the
// critical path of every tile.
constexpr int kC8Threads = 384;
template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kC8Threads, 1)
conv_c8_kernel(const __grid_constant__ CUtensorMap tmA16, const __grid_constant__ CUtensorMap tmA8h,
               const __grid_constant__ CUtensorMap tmA8l, const __grid_constant__ CUtensorMap tmW16,
               const __grid_constant__ CUtensorMap tmW8h, const __grid_constant__ CUtensorMap tmW8l,
               const __grid_constant__ ConvGeom g) {
  using Cfg = C8Cfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kAccBufs = Cfg::kAccBufs;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full = bars;                      // leader only
  uint64_t* empty = bars + kStages;           // per CTA
  uint64_t* tfull = bars + 2 * kStages;       // per CTA
  uint64_t* tempty = bars + 2 * kStages + 2;  // leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA16); ptx::prefetch_tmap(&tmA8h); ptx::prefetch_tmap(&tmA8l);
    ptx::prefetch_tmap(&tmW16); ptx::prefetch_tmap(&tmW8h); ptx::prefetch_tmap(&tmW8l);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full[i], 2);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], 16);   // 8 epilogue warps x 2 CTAs
    }
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

  const int nTiles = g.w.N / BLOCK_N;
  const int mTiles = g.tilesX * g.tilesY * g.tilesB;
  const int pairM = (mTiles + 1) / 2;
  const int tilesPerGroup = nTiles * pairM;
  const int totalTiles = tilesPerGroup * g.nGroups;
  const int totalItems = conv_total_items(g, totalTiles);   // tiles, uniform K-slices or tail-split slices
  const int pairIdx = blockIdx.x >> 1;
  const int numPairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int item = pairIdx; item < totalItems; item += numPairs) {
      const ConvItem wi = conv_decode_item(g, item, totalTiles);
      const int tile = wi.tile, ks = wi.ks, kSplit = wi.nsplit;
      const int grp = tile / tilesPerGroup;
      const int tl = tile - grp * tilesPerGroup;
      const int nt = tl % nTiles;
      int mt = (tl / nTiles) * 2 + (int)rank;
      int x0, y0, b0;
      if (mt < mTiles) {
        const int tx = mt % g.tilesX;
        mt /= g.tilesX;
        x0 = tx * g.BX; y0 = (mt % g.tilesY) * g.BY; b0 = (mt / g.tilesY) * g.BB;
      } else {
        x0 = 0; y0 = 0; b0 = g.tilesB * g.BB;
      }
      const int n0 = nt * BLOCK_N + (int)rank * (BLOCK_N / 2);
      const int numK = g.grpTapCount[grp] * g.cBlocks;
      const int kBeg = ks * numK / kSplit, kEnd = (ks + 1) * numK / kSplit;
      int t = g.grpTapStart[grp] + kBeg / g.cBlocks;
      int cb = kBeg % g.cBlocks;
      for (int kb = kBeg; kb < kEnd; ++kb) {
        const Tap tap = g.taps[t];
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kStageBytes;
        if (leader) ptx::mbar_arrive_expect_tx(&full[stage], 2 * Cfg::kStageBytes);
        const int c0 = cb * kBlockK, xx = x0 + tap.dx, yy = y0 + tap.dy;
        ptx::tma_load_5d_2sm(st, &tmA16, &full[stage], c0, xx, yy, tap.plane, b0);
        ptx::tma_load_3d_2sm(st + Cfg::kOffW16, &tmW16, &full[stage], c0, n0, tap.w);
        ptx::tma_load_5d_2sm(st + Cfg::kOffA8h, &tmA8h, &full[stage], c0, xx, yy, tap.plane, b0);
        ptx::tma_load_5d_2sm(st + Cfg::kOffA8l, &tmA8l, &full[stage], c0, xx, yy, tap.plane, b0);
        ptx::tma_load_3d_2sm(st + Cfg::kOffW8h, &tmW8h, &full[stage], c0, n0, tap.w);
        ptx::tma_load_3d_2sm(st + Cfg::kOffW8l, &tmW8l, &full[stage], c0, n0, tap.w);
        if (!leader) ptx::mbar_arrive_remote(&full[stage], 0);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        if (++cb == g.cBlocks) { cb = 0; ++t; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    const uint32_t idesc16 = umma_idesc_16(g.mainBf16 ? 1u : 0u, 2 * kTileM, BLOCK_N);
    constexpr uint32_t idesc8 = ptx::umma_idesc_e4m3(2 * kTileM, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = pairIdx; item < totalItems; item += numPairs, ++it) {
      const ConvItem wi = conv_decode_item(g, item, totalTiles);
      const int tile = wi.tile, ks = wi.ks, kSplit = wi.nsplit;
      const int acc = it % kAccBufs;
      const uint32_t aphase = (it / kAccBufs) & 1;
      ptx::mbar_wait(&tempty[acc], aphase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d1 = tmem_base + acc * 2 * BLOCK_N;
      const uint32_t d2 = d1 + BLOCK_N;
      const int numKg = g.grpTapCount[tile / tilesPerGroup] * g.cBlocks;
      const int numK = (ks + 1) * numKg / kSplit - ks * numKg / kSplit;
      for (int kb = 0; kb < numK; ++kb) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sA = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)
          ptx::umma_bf16_2cta(d1, ptx::umma_smem_desc_sw128(sA + k * 32, 0, 1024),
                              ptx::umma_smem_desc_sw128(sA + Cfg::kOffW16 + k * 32, 0, 1024), idesc16, (kb | k) != 0);
#pragma unroll
        for (int k = 0; k < kBlockK / 32; ++k) {
          const uint64_t dA8h = umma_smem_desc_sw64(sA + Cfg::kOffA8h + k * 32);
          const uint64_t dA8l = umma_smem_desc_sw64(sA + Cfg::kOffA8l + k * 32);
          const uint64_t dW8h = umma_smem_desc_sw64(sA + Cfg::kOffW8h + k * 32);
          const uint64_t dW8l = umma_smem_desc_sw64(sA + Cfg::kOffW8l + k * 32);
          ptx::umma_f8_2cta(d2, dA8h, dW8l, idesc8, (kb | k) != 0);
          ptx::umma_f8_2cta(d2, dA8l, dW8h, idesc8, 1);
        }
        ptx::umma_commit_2cta(&empty[stage], 0x3);
        if (kb == numK - 1) ptx::umma_commit_2cta(&tfull[acc], 0x3);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int colHalf = (warp - 4) >> 2;              // which half of the tile's columns this warp drains
    constexpr int kChunksPerWarp = BLOCK_N / 64;
    // out = c1 * (D1 + c2 * D2): host multipliers times the operands' device-side scale records
    // {1/S, 1/E} (S scales the 16-bit plane, E the e4m3 planes; 2^-11 = residual pre-scale)
    float c1 = g.c8OutScale, c2 = g.c8CorrScale;
    if (g.c8RecA && g.c8RecW) {
      c1 *= __ldg(g.c8RecA) * __ldg(g.c8RecW);
      c2 *= __ldg(g.c8RecA + 1) * __ldg(g.c8RecW + 1);
    }
    int it = 0;
    for (int item = pairIdx; item < totalItems; item += numPairs, ++it) {
      const ConvItem wi = conv_decode_item(g, item, totalTiles);
      const int tile = wi.tile, ks = wi.ks, kSplit = wi.nsplit;
      const int acc = it % kAccBufs;
      const uint32_t aphase = (it / kAccBufs) & 1;
      const int grp = tile / tilesPerGroup;
      const int tl = tile - grp * tilesPerGroup;
      const int nt = tl % nTiles;
      int mt = (tl / nTiles) * 2 + (int)rank;
      const bool real = mt < mTiles;
      const int tx = mt % g.tilesX;
      mt /= g.tilesX;
      const int ty = mt % g.tilesY;
      const int tb = mt / g.tilesY;
      const int n0 = nt * BLOCK_N;
      const int bx = row % g.BX;
      const int by = (row / g.BX) % g.BY;
      const int bb = row / (g.BX * g.BY);
      const int x = tx * g.BX + bx, y = ty * g.BY + by;
      int b = tb * g.BB + bb;
      const bool valid = real && (x < g.oX) && (y < g.oY) && (b < g.oB);
      if (!real) b = g.oB;
      const long long off = g.grpOutOff[grp] + (long long)b * g.sB + (long long)y * g.sY +
                            (long long)x * g.sX + (long long)(n0 / g.nSplit) * g.sNhi + (n0 % g.nSplit);
      float* orow = g.out + off;
      const float* arow = g.addsrc ? g.addsrc + off : nullptr;

      // plain stores go through the warp's coalescing buffer (epilogue_chunk_staged)
      const bool staged = !arow && wi.slot < 0 && kSplit == 1;
      float* rowp[8];
      unsigned vmask = 0;
      if (staged) {
        vmask = __ballot_sync(0xffffffffu, valid);
        const unsigned long long mine = reinterpret_cast<unsigned long long>(orow);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          rowp[i] = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, mine, (i >> 2) * 16 + (i & 3) * 4 + (lane >> 3)));
      }
      float* sbuf = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes + 256) + (warp - 4) * kStageFloatsPerWarp;

      ptx::mbar_wait(&tfull[acc], aphase);
      ptx::tc_fence_after();
      const uint32_t t1 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 2 * BLOCK_N;
      const uint32_t t2 = t1 + BLOCK_N;
#pragma unroll 1
      for (int j = colHalf * kChunksPerWarp; j < (colHalf + 1) * kChunksPerWarp; ++j) {
        uint32_t v1[32], v2[32];
        ptx::tmem_ld32(t1 + j * 32, v1);
        ptx::tmem_ld32(t2 + j * 32, v2);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          v1[i] = __float_as_uint(c1 * fmaf(c2, __uint_as_float(v2[i]), __uint_as_float(v1[i])));
        if (wi.slot >= 0) {   // tail split: raw partial -> scratch, merged by conv_tail_fixup
          store_partial_chunk(g.tailScratch + (((size_t)wi.slot * 2 + rank) * kTileM + row) * BLOCK_N + j * 32, v1);
        } else if (staged) {
          float* rp[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) rp[i] = rowp[i] + j * 32;
          epilogue_chunk_staged(g, v1, valid, sbuf, rp, vmask, n0 + j * 32, lane, b);
        } else {
          epilogue_chunk(g, v1, valid, orow + j * 32, arow ? arow + j * 32 : nullptr, n0 + j * 32, lane, b, kSplit > 1, ks == 0);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) ptx::mbar_arrive(&tempty[acc]);
        else ptx::mbar_arrive_remote(&tempty[acc], 0);
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 2) ptx::tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
}


int num_sms_c8() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BLOCK_N>
cudaError_t launch_c8_t(const ConvGeom& g, cudaStream_t stream) {
  using Cfg = C8Cfg<BLOCK_N>;
  CUtensorMap tmA16, tmA8h, tmA8l, tmW16, tmW8h, tmW8l;
  if (!make_act_tmap(&tmA16, g.a.hi, g.a, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
  if (!make_wgt_tmap(&tmW16, g.w.hi, g.w, BLOCK_N / 2)) return cudaErrorInvalidValue;
  if (!make_plane8_tmap(&tmA8h, g.a.h8, g.a, kBlockK, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
  if (!make_plane8_tmap(&tmA8l, g.a.l8, g.a, kBlockK, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
  if (!make_wgt8_tmap(&tmW8h, g.w.h8, g.w, BLOCK_N / 2)) return cudaErrorInvalidValue;
  if (!make_wgt8_tmap(&tmW8l, g.w.l8, g.w, BLOCK_N / 2)) return cudaErrorInvalidValue;
  static bool attr_done[64] = {};
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  bool& attr_set = attr_done[dev_id & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_c8_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("conv_c8: smem attr: %s", cudaGetErrorString(e)); return e; }
    attr_set = true;
  }
  const int mTiles = g.tilesX * g.tilesY * g.tilesB;
  const int tilesAll = (g.w.N / BLOCK_N) * ((mTiles + 1) / 2) * g.nGroups;
  const int total = g.tailTiles > 0 ? tilesAll - g.tailTiles + g.tailTiles * g.tailSplit : tilesAll * (g.kSplit > 1 ? g.kSplit : 1);
  const int maxPairs = num_sms_c8() / 2;
  const int pairs = total < maxPairs ? total : maxPairs;
  profile_begin(BLOCK_N == 256 ? kProfConvC8w : kProfConvC8n, g.algoFlops, stream);
  conv_c8_kernel<BLOCK_N><<<2 * pairs, kC8Threads, Cfg::kSmemBytes, stream>>>(tmA16, tmA8h, tmA8l, tmW16, tmW8h, tmW8l, g);
  profile_end(stream);
  cudaError_t e = launched();
  if (e == cudaSuccess && g.tailTiles > 0) e = launch_conv_tail_fixup(g, BLOCK_N, stream);
  return e;
}

}  // namespace

cudaError_t launch_conv_c8(const ConvGeom& g, int blockN, cudaStream_t stream) {
  if (g.BX * g.BY * g.BB != kTileM) { set_error("conv_c8: box %dx%dx%d != 128", g.BX, g.BY, g.BB); return cudaErrorInvalidValue; }
  if (blockN != 128 && blockN != 256) { set_error("conv_c8: blockN must be 128 or 256"); return cudaErrorInvalidValue; }
  if (g.w.N % blockN || g.nSplit % blockN) { set_error("conv_c8: N=%d / nSplit=%d not multiples of %d", g.w.N, g.nSplit, blockN); return cudaErrorInvalidValue; }
  if (g.a.C % kBlockK || g.a.C != g.w.K || g.cBlocks != g.a.C / kBlockK) { set_error("conv_c8: C=%d K=%d", g.a.C, g.w.K); return cudaErrorInvalidValue; }
  if (!g.a.h8 || !g.a.l8 || !g.w.h8 || !g.w.l8) { set_error("conv_c8: 8-bit planes missing"); return cudaErrorInvalidValue; }
  if (g.nTaps < 1 || g.nTaps > kMaxTaps || g.nGroups < 1 || g.nGroups > 4) { set_error("conv_c8: taps/groups"); return cudaErrorInvalidValue; }
  if (g.tailTiles > 0 && (g.kSplit > 1 || g.tailSplit < 2 || !g.tailScratch)) { set_error("conv_c8: bad tail split"); return cudaErrorInvalidValue; }
  if (g.kSplit > 1) {
    if (g.statSum) { set_error("conv_c8: split-K cannot feed the fused statistics"); return cudaErrorInvalidValue; }
    for (int i = 0; i < g.nGroups; ++i)
      if (g.kSplit > g.grpTapCount[i] * g.cBlocks) { set_error("conv_c8: kSplit=%d exceeds the k-blocks of group %d", g.kSplit, i); return cudaErrorInvalidValue; }
  }
  return blockN == 256 ? launch_c8_t<256>(g, stream) : launch_c8_t<128>(g, stream);
}

// ------------------------------------------------------------------------------------------------
// SIMT checker of the same arithmetic on the same planes (shares no descriptor / swizzle logic).
namespace {
__device__ __forceinline__ float e4m3_to_float(uint8_t b) {
  const int e = (b >> 3) & 0xF, m = b & 7;
  const float mag = e == 0 ? (float)m * 0.001953125f /* 2^-9 */ : ldexpf(1.f + (float)m * 0.125f, e - 7);
  return (b & 0x80) ? -mag : mag;
}
__device__ __forceinline__ float f16_bits_to_float(uint16_t v, int isBf16) {
  if (isBf16) return __uint_as_float(static_cast<uint32_t>(v) << 16);
  return __half2float(__ushort_as_half(v));
}
__global__ void conv_c8_simt_kernel(const __grid_constant__ ConvGeom g) {
  const long long perGroup = (long long)g.tilesX * g.tilesY * g.tilesB * kTileM * g.w.N;
  const long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= perGroup * g.nGroups) return;
  const int grp = (int)(gidx / perGroup);
  const long long idx = gidx - grp * perGroup;
  const int n = (int)(idx % g.w.N);
  long long r = idx / g.w.N;
  const int row = (int)(r % kTileM);
  int mt = (int)(r / kTileM);
  const int tx = mt % g.tilesX;
  mt /= g.tilesX;
  const int ty = mt % g.tilesY;
  const int tb = mt / g.tilesY;
  const int bx = row % g.BX, by = (row / g.BX) % g.BY, bb = row / (g.BX * g.BY);
  const int x = tx * g.BX + bx, y = ty * g.BY + by, b = tb * g.BB + bb;
  if (x >= g.oX || y >= g.oY || b >= g.oB) return;
  const uint16_t* A16 = reinterpret_cast<const uint16_t*>(g.a.hi);
  const uint16_t* W16 = reinterpret_cast<const uint16_t*>(g.w.hi);
  const uint8_t* A8h = reinterpret_cast<const uint8_t*>(g.a.h8);
  const uint8_t* A8l = reinterpret_cast<const uint8_t*>(g.a.l8);
  const uint8_t* W8h = reinterpret_cast<const uint8_t*>(g.w.h8);
  const uint8_t* W8l = reinterpret_cast<const uint8_t*>(g.w.l8);
  float d1 = 0.f, d2 = 0.f;
  for (int t = g.grpTapStart[grp]; t < g.grpTapStart[grp] + g.grpTapCount[grp]; ++t) {
    const Tap tap = g.taps[t];
    const int xx = x + tap.dx, yy = y + tap.dy;
    if (xx < 0 || xx >= g.a.X || yy < 0 || yy >= g.a.Y) continue;
    const long long aoff = ((((long long)b * g.a.P + tap.plane) * g.a.Y + yy) * g.a.X + xx) * g.a.C;
    const long long woff = ((long long)tap.w * g.w.N + n) * g.w.K;
    for (int c = 0; c < g.a.C; ++c) {
      d1 = fmaf(f16_bits_to_float(A16[aoff + c], g.mainBf16), f16_bits_to_float(W16[woff + c], g.mainBf16), d1);
      d2 = fmaf(e4m3_to_float(A8h[aoff + c]), e4m3_to_float(W8l[woff + c]), d2);
      d2 = fmaf(e4m3_to_float(A8l[aoff + c]), e4m3_to_float(W8h[woff + c]), d2);
    }
  }
  const long long off = g.grpOutOff[grp] + (long long)b * g.sB + (long long)y * g.sY + (long long)x * g.sX +
                        (long long)(n / g.nSplit) * g.sNhi + (n % g.nSplit);
  float c1 = g.c8OutScale, c2 = g.c8CorrScale;
  if (g.c8RecA && g.c8RecW) {
    c1 *= g.c8RecA[0] * g.c8RecW[0];
    c2 *= g.c8RecA[1] * g.c8RecW[1];
  }
  float acc = c1 * fmaf(c2, d2, d1);
  if (g.bias) acc += g.bias[n];
  if (g.addsrc) acc += g.addsrc[off];
  g.out[off] = acc;
}
}  // namespace

cudaError_t launch_conv_c8_simt(const ConvGeom& g, cudaStream_t stream) {
  const long long total = (long long)g.tilesX * g.tilesY * g.tilesB * kTileM * g.w.N * g.nGroups;
  conv_c8_simt_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(g);
  return launched();
}

}  // namespace mcgvc
