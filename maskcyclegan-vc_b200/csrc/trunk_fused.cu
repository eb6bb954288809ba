// Fused forward of the Generator's six gated 1-D residual blocks (sm_100a).
//
// Reference: mask_cyclegan_vc/model.py:40-76 (ResidualLayer) applied six times at :258-263:
//     h = IN(conv3(x)) * sigmoid(IN(conv3_gates(x)));   y = x + IN(conv3_out(h))
// on a (B, 256, W2) tensor.  At batch 64 / 64 frames this is 1.6 % of the Generator's FLOPs but, run
// layer by layer (12 convolutions of 12-24 k-blocks on 32-128 CTAs, 12 statistics kernels, 12
// normalise kernels), 36 dependent launches of 5-25 us each.  Here the whole chain is ONE launch:
//
//   * InstanceNorm1d statistics span only the W2 positions of one sample, so a 128-row tile that
//     holds whole samples (BB samples x BX positions, BX = pow2 >= W2) is an INDEPENDENT chain
//     through all six blocks: no grid-wide dependency exists.
//   * One thread-block CLUSTER of 8 CTAs owns one tile.  CTA j computes 1/8 of every convolution's
//     output columns (conv a: 64 conv + 64 gate columns of the same channels, so the gate stays local;
//     conv b: 32 columns) with the usual TMA -> smem -> tcgen05.mma -> TMEM pipeline, normalises /
//     gates / adds the residual in its epilogue (statistics reduced inside the tile through shared
//     memory), writes its slice of the next operand (bf16 hi/lo planes) and meets the other seven
//     CTAs at a hardware cluster barrier; the next convolution's TMA loads read the slices back
//     through L2.  Two cluster barriers per block replace six kernel boundaries.
//   * Everything the backward pass needs (raw conv outputs z, mean / rstd, operands) is written in the
//     same layout the layer-by-layer path uses, so generator_backward is unchanged.
#include "trunk_fused.cuh"

#include "gemm_types.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

#include <cstdlib>

namespace mcgvc {

namespace {

constexpr int kCluster = 8;
constexpr int kRows = 128;                 // tile rows = UMMA M
constexpr int kNa = 128;                   // conv a columns per CTA: 64 conv + 64 gate
constexpr int kNb = 32;                    // conv b columns per CTA
constexpr int kPitchA = 132;               // fp32 epilogue tile pitch (floats): 16-byte aligned rows, bank spread
constexpr int kPitchB = 36;
constexpr int kMaxBB = 32;                 // samples per tile (BX >= 4)

template <int NPASS>
struct TrunkCfg {
  static constexpr int kABytes = kRows * kBlockK * 2;                 // 16 KB activation box
  static constexpr int kBBytes = kNa * kBlockK * 2;                   // 16 KB weight rows (conv a); conv b uses 4 KB of it
  static constexpr int kPlanes = NPASS == 3 ? 2 : 1;
  static constexpr int kStageBytes = (kABytes + kBBytes) * kPlanes;   // 64 KB / 32 KB
  static constexpr int kStages = NPASS == 3 ? 3 : 6;
  static constexpr int kRingBytes = kStages * kStageBytes;            // 192 KB
  // the epilogue's fp32 tile + per-sample statistics alias the (idle) ring
  static constexpr int kEpiBytes = kRows * kPitchA * 4 + kMaxBB * kNa * 2 * 4;
  static constexpr int kSmemBytes = kRingBytes + 1024 + 256;
  static_assert(kEpiBytes <= kRingBytes, "epilogue scratch must fit in the operand ring");
};

__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }

__device__ __forceinline__ void epi_bar() {   // the 128 epilogue threads (warps 4-7)
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

__device__ __forceinline__ void store_split8(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&v)[8]) {
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
    const __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(l0, l1);
    ph[i] = *reinterpret_cast<const uint32_t*>(&a);
    pl[i] = *reinterpret_cast<const uint32_t*>(&b);
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  *reinterpret_cast<uint4*>(lo) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

// per-(sample, column) mean / rstd over the W2 valid rows of each sample of the tile (two-pass, biased
// variance, eps = 1e-5 inside the sqrt: nn.InstanceNorm1d defaults).  nCols threads own one column each;
// with nCols < 128 the 128 / nCols thread groups share the samples.
__device__ __forceinline__ void tile_stats(const float* tile, int pitch, int nCols, float* stat, int et,
                                           int BX, int BB, int W2, int B, int b0, float* gMean, float* gRstd,
                                           int statN, int ncol) {
  const int col = et % nCols, grp = et / nCols, groups = 128 / nCols;
  const float inv = 1.f / (float)W2;
  for (int bb = grp; bb < BB; bb += groups) {
    if (b0 + bb >= B) break;
    const float* c = tile + (size_t)bb * BX * pitch + col;
    float s = 0.f;
    for (int x = 0; x < W2; ++x) s += c[(size_t)x * pitch];
    const float mean = s * inv;
    float q = 0.f;
    for (int x = 0; x < W2; ++x) {
      const float d = c[(size_t)x * pitch] - mean;
      q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(q * inv + 1e-5f);
    stat[(bb * nCols + col) * 2 + 0] = mean;
    stat[(bb * nCols + col) * 2 + 1] = rstd;
    gMean[(size_t)(b0 + bb) * statN + ncol] = mean;
    gRstd[(size_t)(b0 + bb) * statN + ncol] = rstd;
  }
}

template <int NPASS>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(256, 1)
trunk_fwd_kernel(const __grid_constant__ CUtensorMap tmRh, const __grid_constant__ CUtensorMap tmRl,
                 const __grid_constant__ CUtensorMap tmHh, const __grid_constant__ CUtensorMap tmHl,
                 const __grid_constant__ CUtensorMap tmWah, const __grid_constant__ CUtensorMap tmWal,
                 const __grid_constant__ CUtensorMap tmWbh, const __grid_constant__ CUtensorMap tmWbl,
                 const __grid_constant__ TrunkFwdArgs p) {
  using Cfg = TrunkCfg<NPASS>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kRingBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();
  const int tileIdx = blockIdx.x / kCluster;
  const int b0 = tileIdx * p.BB;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmRh); ptx::prefetch_tmap(&tmHh); ptx::prefetch_tmap(&tmWah); ptx::prefetch_tmap(&tmWbh);
    if (NPASS == 3) { ptx::prefetch_tmap(&tmRl); ptx::prefetch_tmap(&tmHl); ptx::prefetch_tmap(&tmWal); ptx::prefetch_tmap(&tmWbl); }
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
    ptx::tmem_alloc(tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // pipeline state of the producer / MMA threads (continues across the 12 convolutions)
  int stage = 0;
  uint32_t phase = 0;
  uint32_t tphase = 0;                       // tfull completes once per convolution
  const bool producer = warp == 0 && lane == 0;
  const bool issuer = warp == 1 && lane == 0;
  const bool epi = warp >= 4;
  const int et = threadIdx.x - 128;          // epilogue thread index
  const int quad = warp & 3;
  const int row = quad * 32 + lane;          // TMEM lane = tile row of this epilogue thread
  float* tile = reinterpret_cast<float*>(smem);
  float* stat = tile + kRows * kPitchA;

  for (int blk = 0; blk < kTrunkBlocks; ++blk) {
    // ================================================================ conv a: R[blk] -> z4, H
    if (producer) {
      for (int kb = 0; kb < 12; ++kb) {            // 3 taps x 4 channel blocks of 64
        const int tap = kb >> 2, cb = kb & 3;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kStageBytes;
        ptx::mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
        ptx::tma_load_4d(st, &tmRh, &full[stage], cb * 64, tap - 1, b0, blk);
        ptx::tma_load_4d(st + Cfg::kABytes, &tmWah, &full[stage], cb * 64, 64 * rank, tap, blk);
        ptx::tma_load_4d(st + Cfg::kABytes + 8192, &tmWah, &full[stage], cb * 64, 512 + 64 * rank, tap, blk);
        if (NPASS == 3) {
          uint8_t* lo = st + Cfg::kABytes + Cfg::kBBytes;
          ptx::tma_load_4d(lo, &tmRl, &full[stage], cb * 64, tap - 1, b0, blk);
          ptx::tma_load_4d(lo + Cfg::kABytes, &tmWal, &full[stage], cb * 64, 64 * rank, tap, blk);
          ptx::tma_load_4d(lo + Cfg::kABytes + 8192, &tmWal, &full[stage], cb * 64, 512 + 64 * rank, tap, blk);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (issuer) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(kRows, kNa, 0, 0);
      for (int kb = 0; kb < 12; ++kb) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sA = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t sB = sA + Cfg::kABytes;
        const uint32_t sAl = sB + Cfg::kBBytes;
        const uint32_t sBl = sAl + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t dAh = ptx::umma_smem_desc_sw128(sA + k * 32, 0, 1024);
          const uint64_t dBh = ptx::umma_smem_desc_sw128(sB + k * 32, 0, 1024);
          ptx::umma_bf16(tmem_base, dAh, dBh, idesc, (kb | k) != 0);
          if (NPASS == 3) {
            const uint64_t dAl = ptx::umma_smem_desc_sw128(sAl + k * 32, 0, 1024);
            const uint64_t dBl = ptx::umma_smem_desc_sw128(sBl + k * 32, 0, 1024);
            ptx::umma_bf16(tmem_base, dAh, dBl, idesc, 1);
            ptx::umma_bf16(tmem_base, dAl, dBh, idesc, 1);
          }
        }
        ptx::umma_commit(&empty[stage]);
        if (kb == 11) ptx::umma_commit(tfull);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (epi) {
      ptx::mbar_wait(tfull, tphase);
      ptx::tc_fence_after();
      // -- 1: TMEM -> fp32 tile (+ bias).  TMEM column c < 64: conv channel 64*rank + c; c >= 64: its gate
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      const float* bias = p.biasA[blk];
#pragma unroll 1
      for (int j = 0; j < kNa / 32; ++j) {
        uint32_t v[32];
        ptx::tmem_ld32(taddr + j * 32, v);
        ptx::tmem_ld_wait();
        const int n0 = (j < 2 ? 64 * rank + j * 32 : 512 + 64 * rank + (j - 2) * 32);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n0 + i));
          float4 t;
          t.x = __uint_as_float(v[i]) + bv.x; t.y = __uint_as_float(v[i + 1]) + bv.y;
          t.z = __uint_as_float(v[i + 2]) + bv.z; t.w = __uint_as_float(v[i + 3]) + bv.w;
          *reinterpret_cast<float4*>(tile + row * kPitchA + j * 32 + i) = t;
        }
      }
      ptx::tc_fence_before();
      epi_bar();
      // -- 2: InstanceNorm statistics per (sample, column)
      {
        const int col = et;   // 128 columns, one per thread
        const int ncol = col < 64 ? 64 * rank + col : 512 + 64 * rank + (col - 64);
        tile_stats(tile, kPitchA, kNa, stat, et, p.BX, p.BB, p.W2, p.B, b0, p.mean4[blk], p.rstd4[blk], 1024, ncol);
      }
      epi_bar();
      // -- 3: normalise, gate, write H (hi/lo) and the raw z4; thread = (row group, 8-channel group)
      {
        const int cg = et & 7, rsub = et >> 3;
        const int c0 = cg * 8;                         // column of the conv half; gate = 64 + c0
        const int chan = 64 * rank + c0;               // channel of H / conv column of z4
        float ga[8], ba[8], gg[8], bg[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          ga[k] = __ldg(p.gammaA[blk] + chan + k); ba[k] = __ldg(p.betaA[blk] + chan + k);
          gg[k] = __ldg(p.gammaA[blk] + 512 + chan + k); bg[k] = __ldg(p.betaA[blk] + 512 + chan + k);
        }
#pragma unroll 1
        for (int it = 0; it < kRows / 16; ++it) {
          const int r = it * 16 + rsub;
          const int bb = r / p.BX, bx = r - bb * p.BX;
          const int b = b0 + bb;
          if (b >= p.B || bx >= p.W2) continue;
          const float* tr = tile + r * kPitchA;
          const float* sa = stat + (bb * kNa + c0) * 2;
          const float* sg = stat + (bb * kNa + 64 + c0) * 2;
          float za[8], zg[8], h[8];
          *reinterpret_cast<float4*>(za) = *reinterpret_cast<const float4*>(tr + c0);
          *reinterpret_cast<float4*>(za + 4) = *reinterpret_cast<const float4*>(tr + c0 + 4);
          *reinterpret_cast<float4*>(zg) = *reinterpret_cast<const float4*>(tr + 64 + c0);
          *reinterpret_cast<float4*>(zg + 4) = *reinterpret_cast<const float4*>(tr + 64 + c0 + 4);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float a = fmaf((za[k] - sa[2 * k]) * sa[2 * k + 1], ga[k], ba[k]);
            const float g = fmaf((zg[k] - sg[2 * k]) * sg[2 * k + 1], gg[k], bg[k]);
            h[k] = a * sigmoid_fast(g);
          }
          const size_t grow = (size_t)b * p.W2 + bx;
          float* zo = p.z4[blk] + grow * 1024 + chan;
          *reinterpret_cast<float4*>(zo) = *reinterpret_cast<const float4*>(za);
          *reinterpret_cast<float4*>(zo + 4) = *reinterpret_cast<const float4*>(za + 4);
          *reinterpret_cast<float4*>(zo + 512) = *reinterpret_cast<const float4*>(zg);
          *reinterpret_cast<float4*>(zo + 516) = *reinterpret_cast<const float4*>(zg + 4);
          store_split8(p.Hhi[blk] + grow * 512 + chan, p.Hlo[blk] + grow * 512 + chan, h);
        }
      }
      ptx::fence_proxy_async_all();   // H written with st.global is read by the other CTAs' TMA; ring smem reused by TMA
    }
    tphase ^= 1;
    ptx::cluster_sync_relacq();
    if (producer) ptx::fence_proxy_async_all();

    // ================================================================ conv b: H[blk] -> z5, R[blk + 1]
    if (producer) {
      constexpr uint32_t kTx = (Cfg::kABytes + kNb * kBlockK * 2) * Cfg::kPlanes;
      for (int kb = 0; kb < 24; ++kb) {            // 3 taps x 8 channel blocks of 64
        const int tap = kb >> 3, cb = kb & 7;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kStageBytes;
        ptx::mbar_arrive_expect_tx(&full[stage], kTx);
        ptx::tma_load_4d(st, &tmHh, &full[stage], cb * 64, tap - 1, b0, blk);
        ptx::tma_load_4d(st + Cfg::kABytes, &tmWbh, &full[stage], cb * 64, kNb * rank, tap, blk);
        if (NPASS == 3) {
          uint8_t* lo = st + Cfg::kABytes + Cfg::kBBytes;
          ptx::tma_load_4d(lo, &tmHl, &full[stage], cb * 64, tap - 1, b0, blk);
          ptx::tma_load_4d(lo + Cfg::kABytes, &tmWbl, &full[stage], cb * 64, kNb * rank, tap, blk);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (issuer) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(kRows, kNb, 0, 0);
      const uint32_t d_tmem = tmem_base + kNa;
      for (int kb = 0; kb < 24; ++kb) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sA = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t sB = sA + Cfg::kABytes;
        const uint32_t sAl = sB + Cfg::kBBytes;
        const uint32_t sBl = sAl + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t dAh = ptx::umma_smem_desc_sw128(sA + k * 32, 0, 1024);
          const uint64_t dBh = ptx::umma_smem_desc_sw128(sB + k * 32, 0, 1024);
          ptx::umma_bf16(d_tmem, dAh, dBh, idesc, (kb | k) != 0);
          if (NPASS == 3) {
            const uint64_t dAl = ptx::umma_smem_desc_sw128(sAl + k * 32, 0, 1024);
            const uint64_t dBl = ptx::umma_smem_desc_sw128(sBl + k * 32, 0, 1024);
            ptx::umma_bf16(d_tmem, dAh, dBl, idesc, 1);
            ptx::umma_bf16(d_tmem, dAl, dBh, idesc, 1);
          }
        }
        ptx::umma_commit(&empty[stage]);
        if (kb == 23) ptx::umma_commit(tfull);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (epi) {
      ptx::mbar_wait(tfull, tphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + kNa;
      const int nb0 = kNb * rank;                      // first of this CTA's 32 output channels
      {
        uint32_t v[32];
        ptx::tmem_ld32(taddr, v);
        ptx::tmem_ld_wait();
        const float* bias = p.biasB[blk] + nb0;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + i));
          float4 t;
          t.x = __uint_as_float(v[i]) + bv.x; t.y = __uint_as_float(v[i + 1]) + bv.y;
          t.z = __uint_as_float(v[i + 2]) + bv.z; t.w = __uint_as_float(v[i + 3]) + bv.w;
          *reinterpret_cast<float4*>(tile + row * kPitchB + i) = t;
        }
      }
      ptx::tc_fence_before();
      epi_bar();
      tile_stats(tile, kPitchB, kNb, stat, et, p.BX, p.BB, p.W2, p.B, b0, p.mean5[blk], p.rstd5[blk], 256, nb0 + (et & 31));
      epi_bar();
      {
        const int cg = et & 3, rsub = et >> 2;         // 4 channel groups of 8, 32 rows per pass
        const int c0 = cg * 8;
        const int chan = nb0 + c0;
        float gm[8], bt[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { gm[k] = __ldg(p.gammaB[blk] + chan + k); bt[k] = __ldg(p.betaB[blk] + chan + k); }
#pragma unroll 1
        for (int it = 0; it < kRows / 32; ++it) {
          const int r = it * 32 + rsub;
          const int bb = r / p.BX, bx = r - bb * p.BX;
          const int b = b0 + bb;
          if (b >= p.B || bx >= p.W2) continue;
          const float* tr = tile + r * kPitchB + c0;
          const float* ss = stat + (bb * kNb + c0) * 2;
          const size_t grow = (size_t)b * p.W2 + bx;
          float z[8], res[8], o[8];
          *reinterpret_cast<float4*>(z) = *reinterpret_cast<const float4*>(tr);
          *reinterpret_cast<float4*>(z + 4) = *reinterpret_cast<const float4*>(tr + 4);
          const float* rin = p.Rf[blk] + grow * 256 + chan;
          *reinterpret_cast<float4*>(res) = *reinterpret_cast<const float4*>(rin);
          *reinterpret_cast<float4*>(res + 4) = *reinterpret_cast<const float4*>(rin + 4);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = fmaf((z[k] - ss[2 * k]) * ss[2 * k + 1], gm[k], bt[k]) + res[k];
          float* zo = p.z5[blk] + grow * 256 + chan;
          *reinterpret_cast<float4*>(zo) = *reinterpret_cast<const float4*>(z);
          *reinterpret_cast<float4*>(zo + 4) = *reinterpret_cast<const float4*>(z + 4);
          float* ro = p.Rf[blk + 1] + grow * 256 + chan;
          *reinterpret_cast<float4*>(ro) = *reinterpret_cast<const float4*>(o);
          *reinterpret_cast<float4*>(ro + 4) = *reinterpret_cast<const float4*>(o + 4);
          store_split8(p.Rhi[blk + 1] + grow * 256 + chan, p.Rlo[blk + 1] + grow * 256 + chan, o);
        }
      }
      ptx::fence_proxy_async_all();
    }
    tphase ^= 1;
    ptx::cluster_sync_relacq();
    if (producer) ptx::fence_proxy_async_all();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 256);
}

// ================================================================================================
// Backward chain.  Per block i = 5..0 (dR[i+1] lives in each CTA's shared memory as its 32-column slice):
//   S1  InstanceNorm backward of conv b's output on the CTA's 32 columns -> dz5 slice (global hi/lo)
//       -- cluster barrier: dz5 complete --
//   S2  data gradient of conv b, N-split: dH[:, 64j..64j+64) = sum_taps dz5 * Wb^T      (12 k-blocks)
//   S3  GLU + InstanceNorm backward on the SAME channels (conv and gate columns 64j.. of z4) -> dz4 slice
//   S4  data gradient of conv a, K-SPLIT: the CTA's own 128 dz4 channels x 3 taps against all 256 output
//       columns (12 half-k-block items of 64 KB) -> partial dR[i] (128 x 256 fp32) in shared memory
//       -- cluster barrier: partials ready --
//   S5  each CTA sums the 8 partials of its 32-column slice over distributed shared memory and adds the
//       skip connection: dR[i] slice                                -- cluster barrier: partials consumed --
// The K-split keeps conv a's data gradient (K = 3072) from streaming the whole dz4 tile into every CTA.
constexpr int kPitchH = 68;                // dH tile pitch (floats)
constexpr int kPitchP = 260;               // partial dR tile pitch (floats)
constexpr int kPitchR = 36;                // dR slice pitch (floats)

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

template <int NPASS>
struct TrunkBwdCfg {
  static constexpr int kSlot = NPASS == 3 ? 65536 : 32768;            // ring slot: (A 16 KB + B 16 KB) x planes
  static constexpr int kPlanes = NPASS == 3 ? 2 : 1;
  static constexpr int kStages = NPASS == 3 ? 3 : 6;
  static constexpr int kRingBytes = kStages * kSlot;                  // 192 KB
  static constexpr int kDRBytes = kRows * kPitchR * 4;                // persistent dR slice
  static constexpr int kSmemBytes = kRingBytes + kDRBytes + 1024 + 256;
  // scratch of the three epilogues (aliases the idle ring)
  static constexpr int kS3Bytes = (kRows * kPitchH + kRows * kPitchA + kMaxBB * 64 * 4) * 4;
  static constexpr int kS5Bytes = kRows * kPitchP * 4;
  static_assert(kS3Bytes <= kRingBytes && kS5Bytes <= kRingBytes, "epilogue scratch must fit in the ring");
};

template <int NPASS>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(256, 1)
trunk_bwd_kernel(const __grid_constant__ CUtensorMap tmZ5h, const __grid_constant__ CUtensorMap tmZ5l,
                 const __grid_constant__ CUtensorMap tmZ4h, const __grid_constant__ CUtensorMap tmZ4l,
                 const __grid_constant__ CUtensorMap tmWbh, const __grid_constant__ CUtensorMap tmWbl,
                 const __grid_constant__ CUtensorMap tmWah, const __grid_constant__ CUtensorMap tmWal,
                 const __grid_constant__ TrunkBwdArgs p) {
  using Cfg = TrunkBwdCfg<NPASS>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* dRloc = reinterpret_cast<float*>(smem + Cfg::kRingBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kRingBytes + Cfg::kDRBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();
  const int tileIdx = blockIdx.x / kCluster;
  const int b0 = tileIdx * p.BB;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmZ5h); ptx::prefetch_tmap(&tmZ4h); ptx::prefetch_tmap(&tmWbh); ptx::prefetch_tmap(&tmWah);
    if (NPASS == 3) { ptx::prefetch_tmap(&tmZ5l); ptx::prefetch_tmap(&tmZ4l); ptx::prefetch_tmap(&tmWbl); ptx::prefetch_tmap(&tmWal); }
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
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();          // barriers initialised and shared memory live in every CTA before any DSMEM access
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int stage = 0;
  uint32_t phase = 0, tphase = 0;
  const bool producer = warp == 0 && lane == 0;
  const bool issuer = warp == 1 && lane == 0;
  const bool epi = warp >= 4;
  const int et = threadIdx.x - 128;
  const int quad = warp & 3;
  const int row = quad * 32 + lane;
  const float invW = 1.f / (float)p.W2;
  float* ring = reinterpret_cast<float*>(smem);
  const int nb0 = 32 * rank;                 // this CTA's 32 columns of the 256-wide tensors
  const int ch0 = 64 * rank;                 // this CTA's 64 gated channels

  // dR[6] slice -> shared memory (rows outside the batch / beyond W2 stay zero)
  if (epi) {
    const int cg = et & 3, rsub = et >> 2;
    for (int it = 0; it < kRows / 32; ++it) {
      const int r = it * 32 + rsub;
      const int bb = r / p.BX, bx = r - bb * p.BX, b = b0 + bb;
      float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
      if (b < p.B && bx < p.W2) {
        const float* src = p.dR6 + ((size_t)b * p.W2 + bx) * 256 + nb0 + cg * 8;
        v0 = *reinterpret_cast<const float4*>(src);
        v1 = *reinterpret_cast<const float4*>(src + 4);
      }
      *reinterpret_cast<float4*>(dRloc + r * kPitchR + cg * 8) = v0;
      *reinterpret_cast<float4*>(dRloc + r * kPitchR + cg * 8 + 4) = v1;
    }
    epi_bar();
  }

  for (int blk = kTrunkBlocks - 1; blk >= 0; --blk) {
    // ================================================================ S1: IN backward of conv b (32 columns)
    if (epi) {
      float* zt = ring;                              // [128][36] z5 slice
      float* kst = ring + kRows * kPitchR;           // [BB][32][2] k1 = t1/W2, k2 = t2/W2
      {
        const int cg = et & 3, rsub = et >> 2;
        for (int it = 0; it < kRows / 32; ++it) {
          const int r = it * 32 + rsub;
          const int bb = r / p.BX, bx = r - bb * p.BX, b = b0 + bb;
          float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
          if (b < p.B && bx < p.W2) {
            const float* src = p.z5[blk] + ((size_t)b * p.W2 + bx) * 256 + nb0 + cg * 8;
            v0 = *reinterpret_cast<const float4*>(src);
            v1 = *reinterpret_cast<const float4*>(src + 4);
          }
          *reinterpret_cast<float4*>(zt + r * kPitchR + cg * 8) = v0;
          *reinterpret_cast<float4*>(zt + r * kPitchR + cg * 8 + 4) = v1;
        }
      }
      epi_bar();
      {
        const int col = et & 31, grp = et >> 5;
        float gsum1 = 0.f, gsum2 = 0.f;
        for (int bb = grp; bb < p.BB; bb += 4) {
          if (b0 + bb >= p.B) break;
          const float mean = p.mean5[blk][(size_t)(b0 + bb) * 256 + nb0 + col];
          const float rstd = p.rstd5[blk][(size_t)(b0 + bb) * 256 + nb0 + col];
          float t1 = 0.f, t2 = 0.f;
          for (int x = 0; x < p.W2; ++x) {
            const int r = bb * p.BX + x;
            const float dy = dRloc[r * kPitchR + col];
            const float xh = (zt[r * kPitchR + col] - mean) * rstd;
            t1 += dy;
            t2 = fmaf(dy, xh, t2);
          }
          kst[(bb * 32 + col) * 2 + 0] = t1 * invW;
          kst[(bb * 32 + col) * 2 + 1] = t2 * invW;
          gsum1 += t1;
          gsum2 += t2;
        }
        atomicAdd(p.dbetaB[blk] + nb0 + col, gsum1);
        atomicAdd(p.dgammaB[blk] + nb0 + col, gsum2);
      }
      epi_bar();
      {
        const int cg = et & 3, rsub = et >> 2;
        const int c0 = cg * 8;
        float gm[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) gm[k] = __ldg(p.gammaB[blk] + nb0 + c0 + k);
        for (int it = 0; it < kRows / 32; ++it) {
          const int r = it * 32 + rsub;
          const int bb = r / p.BX, bx = r - bb * p.BX, b = b0 + bb;
          if (b >= p.B || bx >= p.W2) continue;
          const float* mp = p.mean5[blk] + (size_t)b * 256 + nb0 + c0;
          const float* rp = p.rstd5[blk] + (size_t)b * 256 + nb0 + c0;
          float dz[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float rs = __ldg(rp + k);
            const float xh = (zt[r * kPitchR + c0 + k] - __ldg(mp + k)) * rs;
            const float dy = dRloc[r * kPitchR + c0 + k];
            dz[k] = rs * gm[k] * (dy - kst[(bb * 32 + c0 + k) * 2] - xh * kst[(bb * 32 + c0 + k) * 2 + 1]);
          }
          const size_t grow = (size_t)b * p.W2 + bx;
          store_split8(p.dz5hi[blk] + grow * 256 + nb0 + c0, p.dz5lo[blk] + grow * 256 + nb0 + c0, dz);
        }
      }
      ptx::fence_proxy_async_all();
    }
    ptx::cluster_sync_relacq();                      // dz5[blk] complete in all 8 slices
    if (producer) ptx::fence_proxy_async_all();

    // ================================================================ S2: dH slice = dgrad of conv b (N = 64)
    if (producer) {
      constexpr uint32_t kTx = (16384 + 8192) * Cfg::kPlanes;
      for (int kb = 0; kb < 12; ++kb) {              // 3 taps x 4 channel blocks of dz5
        const int tap = kb >> 2, cb = kb & 3;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kSlot;
        ptx::mbar_arrive_expect_tx(&full[stage], kTx);
        // data gradient: dH[x] += dz5[x - (tap - 1)] * W[tap]  (the forward tap read x + tap - 1)
        ptx::tma_load_4d(st, &tmZ5h, &full[stage], cb * 64, 1 - tap, b0, blk);
        ptx::tma_load_4d(st + 16384, &tmWbh, &full[stage], cb * 64, ch0, tap, blk);
        if (NPASS == 3) {
          uint8_t* lo = st + 32768;
          ptx::tma_load_4d(lo, &tmZ5l, &full[stage], cb * 64, 1 - tap, b0, blk);
          ptx::tma_load_4d(lo + 16384, &tmWbl, &full[stage], cb * 64, ch0, tap, blk);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (issuer) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(kRows, 64, 0, 0);
      for (int kb = 0; kb < 12; ++kb) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sA = ptx::smem_u32(smem + stage * Cfg::kSlot);
        const uint32_t sB = sA + 16384, sAl = sA + 32768, sBl = sAl + 16384;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t dAh = ptx::umma_smem_desc_sw128(sA + k * 32, 0, 1024);
          const uint64_t dBh = ptx::umma_smem_desc_sw128(sB + k * 32, 0, 1024);
          ptx::umma_bf16(tmem_base, dAh, dBh, idesc, (kb | k) != 0);
          if (NPASS == 3) {
            const uint64_t dAl = ptx::umma_smem_desc_sw128(sAl + k * 32, 0, 1024);
            const uint64_t dBl = ptx::umma_smem_desc_sw128(sBl + k * 32, 0, 1024);
            ptx::umma_bf16(tmem_base, dAh, dBl, idesc, 1);
            ptx::umma_bf16(tmem_base, dAl, dBh, idesc, 1);
          }
        }
        ptx::umma_commit(&empty[stage]);
        if (kb == 11) ptx::umma_commit(tfull);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (epi) {
      // ============================================================== S3: GLU + IN backward -> dz4 slice
      float* dHt = ring;                               // [128][68]
      float* zt = ring + kRows * kPitchH;              // [128][132]: conv cols 0..63, gate cols 64..127
      float* kst = zt + kRows * kPitchA;               // [BB][64][4]: k1a, k2a, k1g, k2g
      // z4 slice -> shared memory while the MMAs run (the ring slots used by the pipeline are ahead of
      // this scratch only in time, not in space: wait for the accumulator first)
      ptx::mbar_wait(tfull, tphase);
      ptx::tc_fence_after();
      {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          uint32_t v[32];
          ptx::tmem_ld32(taddr + j * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(dHt + row * kPitchH + j * 32 + i) =
                make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        }
        ptx::tc_fence_before();
      }
      {
        const int cg = et & 7, rsub = et >> 3;
        for (int it = 0; it < kRows / 16; ++it) {
          const int r = it * 16 + rsub;
          const int bb = r / p.BX, bx = r - bb * p.BX, b = b0 + bb;
          float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, g0 = a0, g1 = a0;
          if (b < p.B && bx < p.W2) {
            const float* src = p.z4[blk] + ((size_t)b * p.W2 + bx) * 1024 + ch0 + cg * 8;
            a0 = *reinterpret_cast<const float4*>(src); a1 = *reinterpret_cast<const float4*>(src + 4);
            g0 = *reinterpret_cast<const float4*>(src + 512); g1 = *reinterpret_cast<const float4*>(src + 516);
          }
          float* d = zt + r * kPitchA + cg * 8;
          *reinterpret_cast<float4*>(d) = a0; *reinterpret_cast<float4*>(d + 4) = a1;
          *reinterpret_cast<float4*>(d + 64) = g0; *reinterpret_cast<float4*>(d + 68) = g1;
        }
      }
      epi_bar();
      {
        const int col = et & 63, grp = et >> 6;        // 64 channels x 2 sample groups
        const float gaa = __ldg(p.gammaA[blk] + ch0 + col), baa = __ldg(p.betaA[blk] + ch0 + col);
        const float gag = __ldg(p.gammaA[blk] + 512 + ch0 + col), bag = __ldg(p.betaA[blk] + 512 + ch0 + col);
        float s1a = 0.f, s2a = 0.f, s1g = 0.f, s2g = 0.f;
        for (int bb = grp; bb < p.BB; bb += 2) {
          if (b0 + bb >= p.B) break;
          const size_t so = (size_t)(b0 + bb) * 1024 + ch0 + col;
          const float ma = p.mean4[blk][so], ra = p.rstd4[blk][so], mg = p.mean4[blk][so + 512], rg = p.rstd4[blk][so + 512];
          float t1a = 0.f, t2a = 0.f, t1g = 0.f, t2g = 0.f;
          for (int x = 0; x < p.W2; ++x) {
            const int r = bb * p.BX + x;
            const float d = dHt[r * kPitchH + col];
            const float xa = (zt[r * kPitchA + col] - ma) * ra, xg = (zt[r * kPitchA + 64 + col] - mg) * rg;
            const float ya = fmaf(xa, gaa, baa), yg = fmaf(xg, gag, bag);
            const float sg = sigmoid_fast(yg);
            const float dya = d * sg, dyg = d * ya * sg * (1.f - sg);
            t1a += dya; t2a = fmaf(dya, xa, t2a);
            t1g += dyg; t2g = fmaf(dyg, xg, t2g);
          }
          float* k = kst + (bb * 64 + col) * 4;
          k[0] = t1a * invW; k[1] = t2a * invW; k[2] = t1g * invW; k[3] = t2g * invW;
          s1a += t1a; s2a += t2a; s1g += t1g; s2g += t2g;
        }
        atomicAdd(p.dbetaA[blk] + ch0 + col, s1a);
        atomicAdd(p.dgammaA[blk] + ch0 + col, s2a);
        atomicAdd(p.dbetaA[blk] + 512 + ch0 + col, s1g);
        atomicAdd(p.dgammaA[blk] + 512 + ch0 + col, s2g);
      }
      epi_bar();
      {
        const int cg = et & 7, rsub = et >> 3;
        const int c0 = cg * 8;
        float gaa[8], baa[8], gag[8], bag[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          gaa[k] = __ldg(p.gammaA[blk] + ch0 + c0 + k); baa[k] = __ldg(p.betaA[blk] + ch0 + c0 + k);
          gag[k] = __ldg(p.gammaA[blk] + 512 + ch0 + c0 + k); bag[k] = __ldg(p.betaA[blk] + 512 + ch0 + c0 + k);
        }
        for (int it = 0; it < kRows / 16; ++it) {
          const int r = it * 16 + rsub;
          const int bb = r / p.BX, bx = r - bb * p.BX, b = b0 + bb;
          if (b >= p.B || bx >= p.W2) continue;
          const size_t so = (size_t)b * 1024 + ch0 + c0;
          float dza[8], dzg[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float ma = __ldg(p.mean4[blk] + so + k), ra = __ldg(p.rstd4[blk] + so + k);
            const float mg = __ldg(p.mean4[blk] + so + 512 + k), rg = __ldg(p.rstd4[blk] + so + 512 + k);
            const float d = dHt[r * kPitchH + c0 + k];
            const float xa = (zt[r * kPitchA + c0 + k] - ma) * ra, xg = (zt[r * kPitchA + 64 + c0 + k] - mg) * rg;
            const float ya = fmaf(xa, gaa[k], baa[k]), yg = fmaf(xg, gag[k], bag[k]);
            const float sg = sigmoid_fast(yg);
            const float dya = d * sg, dyg = d * ya * sg * (1.f - sg);
            const float* kk = kst + (bb * 64 + c0 + k) * 4;
            dza[k] = ra * gaa[k] * (dya - kk[0] - xa * kk[1]);
            dzg[k] = rg * gag[k] * (dyg - kk[2] - xg * kk[3]);
          }
          const size_t grow = (size_t)b * p.W2 + bx;
          store_split8(p.dz4hi[blk] + grow * 1024 + ch0 + c0, p.dz4lo[blk] + grow * 1024 + ch0 + c0, dza);
          store_split8(p.dz4hi[blk] + grow * 1024 + 512 + ch0 + c0, p.dz4lo[blk] + grow * 1024 + 512 + ch0 + c0, dzg);
        }
      }
      ptx::fence_proxy_async_all();
    }
    tphase ^= 1;
    __syncthreads();                                 // this CTA's dz4 slice is written: its own TMA may read it
    if (producer) ptx::fence_proxy_async_all();

    // ================================================================ S4: partial dR = dgrad of conv a over the CTA's K slice
    if (producer) {
      constexpr uint32_t kTx = (16384 + 16384) * Cfg::kPlanes;
      for (int item = 0; item < 12; ++item) {        // (N half) x (tap) x (conv / gate channel block)
        const int nh = item / 6, rem = item - nh * 6, tap = rem >> 1, kc = (rem & 1) ? 512 + ch0 : ch0;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kSlot;
        ptx::mbar_arrive_expect_tx(&full[stage], kTx);
        ptx::tma_load_4d(st, &tmZ4h, &full[stage], kc, 1 - tap, b0, blk);
        ptx::tma_load_4d(st + 16384, &tmWah, &full[stage], kc, nh * 128, tap, blk);
        if (NPASS == 3) {
          uint8_t* lo = st + 32768;
          ptx::tma_load_4d(lo, &tmZ4l, &full[stage], kc, 1 - tap, b0, blk);
          ptx::tma_load_4d(lo + 16384, &tmWal, &full[stage], kc, nh * 128, tap, blk);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (issuer) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(kRows, 128, 0, 0);
      for (int item = 0; item < 12; ++item) {
        const int nh = item / 6, rem = item - nh * 6;
        const uint32_t d_tmem = tmem_base + 128 + nh * 128;
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sA = ptx::smem_u32(smem + stage * Cfg::kSlot);
        const uint32_t sB = sA + 16384, sAl = sA + 32768, sBl = sAl + 16384;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t dAh = ptx::umma_smem_desc_sw128(sA + k * 32, 0, 1024);
          const uint64_t dBh = ptx::umma_smem_desc_sw128(sB + k * 32, 0, 1024);
          ptx::umma_bf16(d_tmem, dAh, dBh, idesc, (rem | k) != 0);
          if (NPASS == 3) {
            const uint64_t dAl = ptx::umma_smem_desc_sw128(sAl + k * 32, 0, 1024);
            const uint64_t dBl = ptx::umma_smem_desc_sw128(sBl + k * 32, 0, 1024);
            ptx::umma_bf16(d_tmem, dAh, dBl, idesc, 1);
            ptx::umma_bf16(d_tmem, dAl, dBh, idesc, 1);
          }
        }
        ptx::umma_commit(&empty[stage]);
        if (item == 11) ptx::umma_commit(tfull);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (epi) {
      float* part = ring;                              // [128][260] this CTA's partial dR[blk]
      ptx::mbar_wait(tfull, tphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + 128;
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        uint32_t v[32];
        ptx::tmem_ld32(taddr + j * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(part + row * kPitchP + j * 32 + i) =
              make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
      }
      ptx::tc_fence_before();
    }
    tphase ^= 1;
    ptx::cluster_sync_relacq();                      // all 8 partials are in shared memory

    // ================================================================ S5: reduce the CTA's 32-column slice over the cluster
    if (epi) {
      const int cg = et & 7, rsub = et >> 3;           // 8 float4 per 32-column row, 16 rows per pass
      const uint32_t partBase = ptx::smem_u32(ring);
      for (int it = 0; it < kRows / 16; ++it) {
        const int r = it * 16 + rsub;
        const uint32_t off = partBase + (uint32_t)((r * kPitchP + nb0 + cg * 4) * 4);
        float4 acc = *reinterpret_cast<const float4*>(dRloc + r * kPitchR + cg * 4);   // skip connection dR[blk + 1]
#pragma unroll
        for (int q = 0; q < kCluster; ++q) {
          const float4 v = ld_dsmem_f4(mapa_u32(off, (uint32_t)((rank + q) & (kCluster - 1))));
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(dRloc + r * kPitchR + cg * 4) = acc;
        if (blk == 0) {
          const int bb = r / p.BX, bx = r - bb * p.BX, b = b0 + bb;
          if (b < p.B && bx < p.W2)
            *reinterpret_cast<float4*>(p.dR0 + ((size_t)b * p.W2 + bx) * 256 + nb0 + cg * 4) = acc;
        }
      }
      ptx::fence_proxy_async_all();                    // the ring is TMA territory again after the next barrier
    }
    ptx::cluster_sync_relacq();                      // every CTA has consumed every partial
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

int env_fused_trunk() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MCGVC_FUSED_TRUNK"); v = e ? atoi(e) : 1; }
  return v;
}

bool make_plane_map(CUtensorMap* m, const void* base, int C, int W2, int B, int nBlk, long long blkStrideBytes,
                    int BX, int BB) {
  const unsigned long long dims[4] = {(unsigned long long)C, (unsigned long long)W2, (unsigned long long)B, (unsigned long long)nBlk};
  const unsigned long long str[3] = {(unsigned long long)C * 2, (unsigned long long)C * 2 * W2, (unsigned long long)blkStrideBytes};
  const unsigned int box[4] = {(unsigned)kBlockK, (unsigned)BX, (unsigned)BB, 1u};
  return make_tmap16(m, base, 4, dims, str, box);
}
bool make_weight_map(CUtensorMap* m, const void* base, int K, int N, int nBlk, long long blkStrideBytes, int boxN) {
  const unsigned long long dims[4] = {(unsigned long long)K, (unsigned long long)N, 3ull, (unsigned long long)nBlk};
  const unsigned long long str[3] = {(unsigned long long)K * 2, (unsigned long long)K * 2 * N, (unsigned long long)blkStrideBytes};
  const unsigned int box[4] = {(unsigned)kBlockK, (unsigned)boxN, 1u, 1u};
  return make_tmap16(m, base, 4, dims, str, box);
}

template <int NPASS>
cudaError_t launch_t(const TrunkFwdArgs& a, const TrunkFwdMaps& m, cudaStream_t stream) {
  using Cfg = TrunkCfg<NPASS>;
  CUtensorMap tRh, tRl, tHh, tHl, tWah, tWal, tWbh, tWbl;
  if (!make_plane_map(&tRh, m.Rhi, 256, a.W2, a.B, kTrunkBlocks + 1, m.RStrideBytes, a.BX, a.BB)) return cudaErrorInvalidValue;
  if (!make_plane_map(&tHh, m.Hhi, 512, a.W2, a.B, kTrunkBlocks, m.HStrideBytes, a.BX, a.BB)) return cudaErrorInvalidValue;
  if (!make_weight_map(&tWah, m.Wah, 256, 1024, kTrunkBlocks, m.WaStrideBytes, 64)) return cudaErrorInvalidValue;
  if (!make_weight_map(&tWbh, m.Wbh, 512, 256, kTrunkBlocks, m.WbStrideBytes, kNb)) return cudaErrorInvalidValue;
  if (NPASS == 3) {
    if (!make_plane_map(&tRl, m.Rlo, 256, a.W2, a.B, kTrunkBlocks + 1, m.RStrideBytes, a.BX, a.BB)) return cudaErrorInvalidValue;
    if (!make_plane_map(&tHl, m.Hlo, 512, a.W2, a.B, kTrunkBlocks, m.HStrideBytes, a.BX, a.BB)) return cudaErrorInvalidValue;
    if (!make_weight_map(&tWal, m.Wal, 256, 1024, kTrunkBlocks, m.WaStrideBytes, 64)) return cudaErrorInvalidValue;
    if (!make_weight_map(&tWbl, m.Wbl, 512, 256, kTrunkBlocks, m.WbStrideBytes, kNb)) return cudaErrorInvalidValue;
  } else {
    tRl = tRh; tHl = tHh; tWal = tWah; tWbl = tWbh;
  }
  static bool attr_done[64] = {};
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  bool& attr_set = attr_done[dev_id & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(trunk_fwd_kernel<NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("trunk_fwd: smem attr: %s", cudaGetErrorString(e)); return e; }
    attr_set = true;
  }
  const int tiles = (a.B + a.BB - 1) / a.BB;
  // algorithmic FLOPs: per row 2 * (1024 * 768 + 256 * 1536) per block
  profile_begin(kProfTrunk, 2.0 * (double)a.B * a.W2 * (1024.0 * 768.0 + 256.0 * 1536.0) * kTrunkBlocks, stream);
  trunk_fwd_kernel<NPASS><<<tiles * kCluster, 256, Cfg::kSmemBytes, stream>>>(tRh, tRl, tHh, tHl, tWah, tWal, tWbh, tWbl, a);
  profile_end(stream);
  return launched();
}

template <int NPASS>
cudaError_t launch_bwd_t(const TrunkBwdArgs& a, const TrunkBwdMaps& m, cudaStream_t stream) {
  using Cfg = TrunkBwdCfg<NPASS>;
  CUtensorMap t5h, t5l, t4h, t4l, tWbh, tWbl, tWah, tWal;
  if (!make_plane_map(&t5h, m.Z5hi, 256, a.W2, a.B, kTrunkBlocks, m.Z5StrideBytes, a.BX, a.BB)) return cudaErrorInvalidValue;
  if (!make_plane_map(&t4h, m.Z4hi, 1024, a.W2, a.B, kTrunkBlocks, m.Z4StrideBytes, a.BX, a.BB)) return cudaErrorInvalidValue;
  if (!make_weight_map(&tWbh, m.Wbh, 256, 512, kTrunkBlocks, m.WbStrideBytes, 64)) return cudaErrorInvalidValue;
  if (!make_weight_map(&tWah, m.Wah, 1024, 256, kTrunkBlocks, m.WaStrideBytes, 128)) return cudaErrorInvalidValue;
  if (NPASS == 3) {
    if (!make_plane_map(&t5l, m.Z5lo, 256, a.W2, a.B, kTrunkBlocks, m.Z5StrideBytes, a.BX, a.BB)) return cudaErrorInvalidValue;
    if (!make_plane_map(&t4l, m.Z4lo, 1024, a.W2, a.B, kTrunkBlocks, m.Z4StrideBytes, a.BX, a.BB)) return cudaErrorInvalidValue;
    if (!make_weight_map(&tWbl, m.Wbl, 256, 512, kTrunkBlocks, m.WbStrideBytes, 64)) return cudaErrorInvalidValue;
    if (!make_weight_map(&tWal, m.Wal, 1024, 256, kTrunkBlocks, m.WaStrideBytes, 128)) return cudaErrorInvalidValue;
  } else {
    t5l = t5h; t4l = t4h; tWbl = tWbh; tWal = tWah;
  }
  static bool attr_done[64] = {};
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  bool& attr_set = attr_done[dev_id & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(trunk_bwd_kernel<NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("trunk_bwd: smem attr: %s", cudaGetErrorString(e)); return e; }
    attr_set = true;
  }
  const int tiles = (a.B + a.BB - 1) / a.BB;
  profile_begin(kProfTrunk, 2.0 * (double)a.B * a.W2 * (1024.0 * 768.0 + 256.0 * 1536.0) * kTrunkBlocks, stream);
  trunk_bwd_kernel<NPASS><<<tiles * kCluster, 256, Cfg::kSmemBytes, stream>>>(t5h, t5l, t4h, t4l, tWbh, tWbl, tWah, tWal, a);
  profile_end(stream);
  return launched();
}

}  // namespace

bool trunk_bwd_supported(int B, int W2) { return trunk_fwd_supported(B, W2); }

cudaError_t launch_trunk_bwd(const TrunkBwdArgs& a, const TrunkBwdMaps& m, cudaStream_t stream) {
  if (a.BX * a.BB != kRows || a.BX < a.W2 || a.BB > kMaxBB) { set_error("trunk_bwd: tile %d x %d", a.BB, a.BX); return cudaErrorInvalidValue; }
  if ((m.Z5StrideBytes | m.Z4StrideBytes | m.WaStrideBytes | m.WbStrideBytes) & 15) { set_error("trunk_bwd: block strides must be 16-byte multiples"); return cudaErrorInvalidValue; }
  return a.nPass == 3 ? launch_bwd_t<3>(a, m, stream) : launch_bwd_t<1>(a, m, stream);
}

bool trunk_fwd_supported(int B, int W2) {
  return env_fused_trunk() != 0 && B >= 1 && W2 >= 4 && W2 <= kRows;
}

cudaError_t launch_trunk_fwd(const TrunkFwdArgs& a, const TrunkFwdMaps& m, cudaStream_t stream) {
  if (a.BX * a.BB != kRows || a.BX < a.W2 || a.BB > kMaxBB) { set_error("trunk_fwd: tile %d x %d", a.BB, a.BX); return cudaErrorInvalidValue; }
  if ((m.RStrideBytes | m.HStrideBytes | m.WaStrideBytes | m.WbStrideBytes) & 15) { set_error("trunk_fwd: block strides must be 16-byte multiples"); return cudaErrorInvalidValue; }
  return a.nPass == 3 ? launch_t<3>(a, m, stream) : launch_t<1>(a, m, stream);
}

}  // namespace mcgvc
