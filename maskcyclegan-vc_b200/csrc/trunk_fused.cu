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
  profile_begin(0, 2.0 * (double)a.B * a.W2 * (1024.0 * 768.0 + 256.0 * 1536.0) * kTrunkBlocks, stream);
  trunk_fwd_kernel<NPASS><<<tiles * kCluster, 256, Cfg::kSmemBytes, stream>>>(tRh, tRl, tHh, tHl, tWah, tWal, tWbh, tWbl, a);
  profile_end(stream);
  return launched();
}

}  // namespace

bool trunk_fwd_supported(int B, int W2) {
  return env_fused_trunk() != 0 && B >= 1 && W2 >= 4 && W2 <= kRows;
}

cudaError_t launch_trunk_fwd(const TrunkFwdArgs& a, const TrunkFwdMaps& m, cudaStream_t stream) {
  if (a.BX * a.BB != kRows || a.BX < a.W2 || a.BB > kMaxBB) { set_error("trunk_fwd: tile %d x %d", a.BB, a.BX); return cudaErrorInvalidValue; }
  if ((m.RStrideBytes | m.HStrideBytes | m.WaStrideBytes | m.WbStrideBytes) & 15) { set_error("trunk_fwd: block strides must be 16-byte multiples"); return cudaErrorInvalidValue; }
  return a.nPass == 3 ? launch_t<3>(a, m, stream) : launch_t<1>(a, m, stream);
}

}  // namespace mcgvc
