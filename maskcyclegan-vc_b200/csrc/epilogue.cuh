// Epilogue helpers shared by the implicit-GEMM convolution kernels (conv_igemm.cu: single-CTA and
// CTA-pair split-bf16 / bf16 kernels; conv_c8.cu: 16-bit + e4m3 kernels).
#pragma once
#include "gemm_types.cuh"

namespace mcgvc {

// Fused InstanceNorm statistics (optional): every epilogue warp reduces its 32 rows x 32 columns
// chunk to per-column sums of z and z^2 with a butterfly "transpose" reduction (31 shuffles per
// quantity instead of 160) and adds them into [image][column] accumulators with atomics.  Rows of a
// warp belong to one image when a tile's positions-per-image (BX*BY) is >= 32; for smaller planes
// (the 1-D trunk) the reduction is segmented: SEG lanes per image, each lane ends up owning 32/SEG
// consecutive columns of its segment's sums.
template <int SEG>
__device__ __forceinline__ void warp_colsum_add(float (&a)[32], float (&b)[32], int lane,
                                                float* dstA, float* dstB, bool img_ok) {
  int n = 32;
#pragma unroll
  for (int s = SEG >> 1; s >= 1; s >>= 1) {
    n >>= 1;
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i < n) {
        const float sendA = up ? a[i] : a[i + n];
        const float keepA = up ? a[i + n] : a[i];
        const float sendB = up ? b[i] : b[i + n];
        const float keepB = up ? b[i + n] : b[i];
        a[i] = keepA + __shfl_xor_sync(0xffffffffu, sendA, s);
        b[i] = keepB + __shfl_xor_sync(0xffffffffu, sendB, s);
      }
    }
  }
  constexpr int kPer = 32 / SEG;   // columns owned by this lane
  if (img_ok) {
    const int start = (lane % SEG) * kPer;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      atomicAdd(dstA + start + i, a[i]);
      atomicAdd(dstB + start + i, b[i]);
    }
  }
}

// Work item of the persistent pair kernels: a whole tile, one of kSplit uniform K-slices of a tile, or --
// tail split -- one of tailSplit K-slices of one of the last tailTiles tiles.
struct ConvItem {
  int tile, ks, nsplit;
  int slot;       // tail split: index of the slice's scratch slab, else -1
};
__device__ __forceinline__ int conv_total_items(const ConvGeom& g, int totalTiles) {
  if (g.tailTiles > 0) return totalTiles - g.tailTiles + g.tailTiles * g.tailSplit;
  return totalTiles * (g.kSplit > 1 ? g.kSplit : 1);
}
__device__ __forceinline__ ConvItem conv_decode_item(const ConvGeom& g, int item, int totalTiles) {
  ConvItem it;
  if (g.tailTiles > 0) {
    const int mainTiles = totalTiles - g.tailTiles;
    if (item < mainTiles) { it.tile = item; it.ks = 0; it.nsplit = 1; it.slot = -1; return it; }
    const int t = item - mainTiles;
    it.tile = mainTiles + t / g.tailSplit;
    it.ks = t - (t / g.tailSplit) * g.tailSplit;
    it.nsplit = g.tailSplit;
    it.slot = t;
    return it;
  }
  const int kSplit = g.kSplit > 1 ? g.kSplit : 1;
  it.tile = item / kSplit;
  it.ks = item - it.tile * kSplit;
  it.nsplit = kSplit;
  it.slot = -1;
  return it;
}
// raw partial accumulator chunk -> scratch slab (tail split)
__device__ __forceinline__ void store_partial_chunk(float* dst, const uint32_t (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 4)
    *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                      __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
}

// one 32-column chunk of this thread's output row: +bias, +residual, store, optional statistics
__device__ __forceinline__ void red_add_f32x4(float* addr, float4 t) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(t.x), "f"(t.y), "f"(t.z),
               "f"(t.w)
               : "memory");
}

// `partial`: this work item holds one K-slice of the tile (split-K): its accumulator is ADDED to the
// zero-initialised output with red.global.add; bias and residual ride on the first slice only.
__device__ __forceinline__ void epilogue_chunk(const ConvGeom& g, const uint32_t (&v)[32], bool valid,
                                               float* orow, const float* arow, int ncol0, int lane,
                                               int b, bool partial, bool first) {
  float o[32];
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float4 t;
    t.x = __uint_as_float(v[i + 0]);
    t.y = __uint_as_float(v[i + 1]);
    t.z = __uint_as_float(v[i + 2]);
    t.w = __uint_as_float(v[i + 3]);
    if (g.bias && first) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(g.bias + ncol0 + i));
      t.x += bv.x; t.y += bv.y; t.z += bv.z; t.w += bv.w;
    }
    if (valid) {
      if (arow && first) {
        const float4 av = *reinterpret_cast<const float4*>(arow + i);
        t.x += av.x; t.y += av.y; t.z += av.z; t.w += av.w;
      }
      if (partial) red_add_f32x4(orow + i, t);
      else *reinterpret_cast<float4*>(orow + i) = t;
    }
    o[i] = t.x; o[i + 1] = t.y; o[i + 2] = t.z; o[i + 3] = t.w;
  }
  if (g.statSum) {
    float q[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      o[i] = valid ? o[i] : 0.f;
      q[i] = o[i] * o[i];
    }
    const bool img_ok = b < g.oB;
    const long long so = (long long)(img_ok ? b : 0) * g.w.N + ncol0;
    float* dA = g.statSum + so;
    float* dB = g.statSq + so;
    switch (g.statSeg) {
      case 32: warp_colsum_add<32>(o, q, lane, dA, dB, img_ok); break;
      case 16: warp_colsum_add<16>(o, q, lane, dA, dB, img_ok); break;
      case 8: warp_colsum_add<8>(o, q, lane, dA, dB, img_ok); break;
      case 4: warp_colsum_add<4>(o, q, lane, dA, dB, img_ok); break;
      case 2: warp_colsum_add<2>(o, q, lane, dA, dB, img_ok); break;
      default: warp_colsum_add<1>(o, q, lane, dA, dB, img_ok); break;
    }
  }
}


// Coalesced variant of the plain-store path.  A TMEM lane is an output row, so the direct epilogue
// writes 32 different 128-byte lines with every STG.128 (16 bytes each): short-K layers (Generator
// stem, heads: 2-3 k-blocks per tile) then run at the rate the L2 takes those partial-sector writes,
// ~1.7 TB/s.  Here a warp stages its chunk through a 16-row x 36-float shared buffer (two rounds of
// 16 rows; conflict-free 16-byte accesses both ways) so that every STG.128 covers four full lines.
// `rowp[h*4+k]` = output pointer (chunk column 0) of row h*16 + k*4 + lane/8, `vmask` = ballot of the
// rows' validity.  Statistics are taken from the registers exactly as in epilogue_chunk.
constexpr int kStagePitch = 36;                          // floats per staged row (32 + 4: no bank conflicts)
constexpr int kStageFloatsPerWarp = 16 * kStagePitch;
__device__ __forceinline__ void epilogue_chunk_staged(const ConvGeom& g, const uint32_t (&v)[32], bool valid,
                                                      float* sbuf, float* const (&rowp)[8], unsigned vmask,
                                                      int ncol0, int lane, int b) {
  float o[32];
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    float4 t;
    t.x = __uint_as_float(v[i + 0]);
    t.y = __uint_as_float(v[i + 1]);
    t.z = __uint_as_float(v[i + 2]);
    t.w = __uint_as_float(v[i + 3]);
    if (g.bias) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(g.bias + ncol0 + i));
      t.x += bv.x; t.y += bv.y; t.z += bv.z; t.w += bv.w;
    }
    o[i] = t.x; o[i + 1] = t.y; o[i + 2] = t.z; o[i + 3] = t.w;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if ((lane >> 4) == h) {
      float* d = sbuf + (lane & 15) * kStagePitch;
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(d + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = k * 4 + (lane >> 3), jj = (lane & 7) * 4;
      const float4 t = *reinterpret_cast<const float4*>(sbuf + r * kStagePitch + jj);
      if ((vmask >> (h * 16 + r)) & 1u) *reinterpret_cast<float4*>(rowp[h * 4 + k] + jj) = t;
    }
    __syncwarp();
  }
  if (g.statSum) {
    float q[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      o[i] = valid ? o[i] : 0.f;
      q[i] = o[i] * o[i];
    }
    const bool img_ok = b < g.oB;
    const long long so = (long long)(img_ok ? b : 0) * g.w.N + ncol0;
    float* dA = g.statSum + so;
    float* dB = g.statSq + so;
    switch (g.statSeg) {
      case 32: warp_colsum_add<32>(o, q, lane, dA, dB, img_ok); break;
      case 16: warp_colsum_add<16>(o, q, lane, dA, dB, img_ok); break;
      case 8: warp_colsum_add<8>(o, q, lane, dA, dB, img_ok); break;
      case 4: warp_colsum_add<4>(o, q, lane, dA, dB, img_ok); break;
      case 2: warp_colsum_add<2>(o, q, lane, dA, dB, img_ok); break;
      default: warp_colsum_add<1>(o, q, lane, dA, dB, img_ok); break;
    }
  }
}


// Coalesced red.global.add of one 32-row x 32-column chunk (weight-gradient kernels): same staging as
// epilogue_chunk_staged; `rowp[h*4+k]` = destination (chunk column 0) of row h*16 + k*4 + lane/8.
__device__ __forceinline__ void red_chunk_staged(const float (&o)[32], float* sbuf, float* const (&rowp)[8],
                                                 int lane) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if ((lane >> 4) == h) {
      float* d = sbuf + (lane & 15) * kStagePitch;
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(d + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = k * 4 + (lane >> 3), jj = (lane & 7) * 4;
      red_add_f32x4(rowp[h * 4 + k] + jj, *reinterpret_cast<const float4*>(sbuf + r * kStagePitch + jj));
    }
    __syncwarp();
  }
}

}  // namespace mcgvc
