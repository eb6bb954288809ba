// Implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a).
//
// One kernel serves every forward convolution and every data-gradient convolution on the
// MaskCycleGAN-VC path (reference: mask_cyclegan_vc/model.py:116-211 Generator layers,
// :290-334 Discriminator layers; their autograd dgrads).  The conv is never lowered to an explicit
// im2col matrix: for every filter tap the A operand is a shifted 128-position x 64-channel box of
// the NHWC activation, fetched by TMA (out-of-bounds = zero padding), the B operand is the tap's
// [N][64] weight slice, and tcgen05.mma accumulates all taps x channel blocks into TMEM.
//
// Warp roles (256 threads, persistent over output tiles):
//   warp 0  lane 0 : TMA producer  (full/empty mbarrier ring)
//   warp 1  lane 0 : MMA issuer    (tcgen05.mma, tcgen05.commit)
//   warp 2         : TMEM allocator
//   warps 4..7     : epilogue (tcgen05.ld -> bias/residual -> global), double-buffered in TMEM
//
// Precision: nPass = 3 runs the split-bf16 scheme  A*W ~= Ah*Wh + Ah*Wl + Al*Wh  (fp32 accumulate),
// nPass = 1 runs plain bf16.
#include "epilogue.cuh"
#include "gemm_types.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <mutex>
#include <vector>

namespace mcgvc {

// ------------------------------------------------------------------------------------------------
// error string (C ABI surfaces it through mcgvc_last_error)
static thread_local char g_err[512] = "";
const char* last_error() { return g_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// launch accounting and optional per-launch event timing
static std::atomic<long long> g_launches{0};
cudaError_t launched() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}
long long launch_count() { return g_launches.load(); }
void add_launches(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

namespace {
struct ProfRec { cudaEvent_t a, b; int kind; double flops; };
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_event_pool;
bool g_prof_on = false;
std::mutex g_prof_mu;
cudaEvent_t get_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace
void profile_enable(bool on) { g_prof_on = on; }
bool profile_enabled() { return g_prof_on; }
void profile_begin(int kind, double flops, cudaStream_t s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{get_event(), get_event(), kind, flops};
  cudaEventRecord(r.a, s);
  g_prof.push_back(r);
}
void profile_end(cudaStream_t s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().b, s);
}
const char* profile_kind_name(int kind) {
  static const char* names[kProfKinds] = {"conv_tc_kernel", "wgrad_tc_kernel / wgrad_tc2_kernel", "conv_c8_kernel<256>",
                                          "conv_c8_kernel<128>", "conv_tc2_kernel<256>", "conv_tc2_kernel<128>",
                                          "trunk_fwd_kernel / trunk_bwd_kernel", "wgrad_c8_kernel"};
  return kind >= 0 && kind < kProfKinds ? names[kind] : "?";
}
void profile_collect_kinds(KernelProfile* out) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < kProfKinds; ++i) out[i] = KernelProfile{0, 0, 0};
  for (ProfRec& r : g_prof) {
    cudaEventSynchronize(r.b);
    float ms = 0.f;
    if (r.kind >= 0 && r.kind < kProfKinds && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      out[r.kind].ms += ms;
      out[r.kind].flops += r.flops;
      out[r.kind].launches += 1;
    }
    g_event_pool.push_back(r.a);
    g_event_pool.push_back(r.b);
  }
  g_prof.clear();
}
void profile_collect(KernelProfile* conv, KernelProfile* wgrad) {
  KernelProfile k[kProfKinds];
  profile_collect_kinds(k);
  KernelProfile c{0, 0, 0}, w{0, 0, 0};
  for (int i = 0; i < kProfKinds; ++i) {
    KernelProfile& d = (i == kProfWgrad || i == kProfWgradC8) ? w : c;
    d.ms += k[i].ms; d.flops += k[i].flops; d.launches += k[i].launches;
  }
  if (conv) *conv = c;
  if (wgrad) *wgrad = w;
}

// ------------------------------------------------------------------------------------------------
// driver entry point for tensor-map encoding (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// bf16 tensor [B][P][Y][X][C] -> rank-5 map, box (64, BX, BY, 1, BB), 128B swizzle, zero OOB fill.
bool make_act_tmap(CUtensorMap* m, const void* base, const ActOperand& a, int BX, int BY, int BB) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[5] = {(cuuint64_t)a.C, (cuuint64_t)a.X, (cuuint64_t)a.Y, (cuuint64_t)a.P,
                        (cuuint64_t)a.B};
  cuuint64_t strides[4];
  strides[0] = (cuuint64_t)a.C * 2;
  strides[1] = strides[0] * a.X;
  strides[2] = strides[1] * a.Y;
  strides[3] = strides[2] * a.P;
  cuuint32_t box[5] = {(cuuint32_t)kBlockK, (cuuint32_t)BX, (cuuint32_t)BY, 1u, (cuuint32_t)BB};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(act C=%d X=%d Y=%d P=%d B=%d box %d,%d,%d) failed: %d", a.C,
              a.X, a.Y, a.P, a.B, BX, BY, BB, (int)r);
    return false;
  }
  return true;
}
// bf16 weights [T][N][K] -> rank-3 map, box (64, boxN, 1).
bool make_wgt_tmap(CUtensorMap* m, const void* base, const WgtOperand& w, int boxN) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)w.K, (cuuint64_t)w.N, (cuuint64_t)w.T};
  cuuint64_t strides[2] = {(cuuint64_t)w.K * 2, (cuuint64_t)w.K * 2 * w.N};
  cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)boxN, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(wgt K=%d N=%d T=%d boxN=%d) failed: %d", w.K, w.N, w.T, boxN,
              (int)r);
    return false;
  }
  return true;
}

bool make_tmap16(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims,
                 const unsigned long long* stridesBytes, const unsigned int* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = stridesBytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(rank %d) failed: %d", rank, (int)r); return false; }
  return true;
}

bool make_plane8_tmap(CUtensorMap* m, const void* base, const ActOperand& a, int boxC, int BX, int BY, int BB) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[5] = {(cuuint64_t)a.C, (cuuint64_t)a.X, (cuuint64_t)a.Y, (cuuint64_t)a.P, (cuuint64_t)a.B};
  cuuint64_t strides[4];
  strides[0] = (cuuint64_t)a.C;
  strides[1] = strides[0] * a.X;
  strides[2] = strides[1] * a.Y;
  strides[3] = strides[2] * a.P;
  cuuint32_t box[5] = {(cuuint32_t)boxC, (cuuint32_t)BX, (cuuint32_t)BY, 1u, (cuuint32_t)BB};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, boxC == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(plane8 C=%d X=%d Y=%d box %d) failed: %d", a.C, a.X, a.Y, boxC, (int)r);
    return false;
  }
  return true;
}
bool make_wgt8_tmap(CUtensorMap* m, const void* base, const WgtOperand& w, int boxN) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)w.K, (cuuint64_t)w.N, (cuuint64_t)w.T};
  cuuint64_t strides[2] = {(cuuint64_t)w.K, (cuuint64_t)w.K * w.N};
  cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)boxN, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(wgt8 K=%d N=%d T=%d) failed: %d", w.K, w.N, w.T, (int)r);
    return false;
  }
  return true;
}

bool choose_box(int B, int Y, int X, int boxPositions, int* oBX, int* oBY, int* oBB) {
  long long best = -1;
  int bx = 0, by = 0, bb = 0;
  for (int BX = 1; BX <= boxPositions && BX <= 256; BX *= 2) {
    for (int BY = 1; BX * BY <= boxPositions; BY *= 2) {
      int BB = boxPositions / (BX * BY);
      if (BX * BY * BB != boxPositions || BB > 256) continue;
      long long tiles = (long long)((X + BX - 1) / BX) * ((Y + BY - 1) / BY) * ((B + BB - 1) / BB);
      // fewer tiles first; tie -> widest X box (longest contiguous runs), then tallest Y box
      long long score = tiles * 1000000LL - BX * 1000LL - BY;
      if (best < 0 || score < best) {
        best = score;
        bx = BX; by = BY; bb = BB;
      }
    }
  }
  if (best < 0) return false;
  *oBX = bx; *oBY = by; *oBB = bb;
  return true;
}


// C8H backward: scale of a single fp16 pass over C8 operand planes (1 for every other call)
template <int NPASS>
__device__ __forceinline__ float half16_scale(const ConvGeom& g) {
  if (NPASS != 1 || !g.half16) return 1.f;
  float s = g.c8OutScale;
  if (g.c8RecA && g.c8RecW) s *= __ldg(g.c8RecA) * __ldg(g.c8RecW);
  return s;
}

// ------------------------------------------------------------------------------------------------
template <int BLOCK_N, int NPASS>
struct ConvCfg {
  static constexpr int kABytes = kTileM * kBlockK * 2;    // 16 KB
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = (kABytes + kBBytes) * (NPASS == 3 ? 2 : 1);
  static constexpr int kOutStageBytes = 4 * kStageFloatsPerWarp * 4;   // coalescing buffers of the 4 epilogue warps
  static constexpr int kBudget = 222 * 1024 - kOutStageBytes;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = 2 * BLOCK_N;           // two accumulator buffers
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kOutStageBytes;
  static_assert(kStages >= 2, "need at least two pipeline stages");
  static_assert(kTmemCols >= 32 && kTmemCols <= 512, "TMEM columns");
};

template <int BLOCK_N, int NPASS>
__global__ void __launch_bounds__(256, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
               const __grid_constant__ ConvGeom g) {
  using Cfg = ConvCfg<BLOCK_N, NPASS>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint64_t* tempty = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmAh);
    ptx::prefetch_tmap(&tmWh);
    if (NPASS == 3) {
      ptx::prefetch_tmap(&tmAl);
      ptx::prefetch_tmap(&tmWl);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], 4);  // one arrival per epilogue warp
    }
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

  const int nTiles = g.w.N / BLOCK_N;
  const int mTiles = g.tilesX * g.tilesY * g.tilesB;
  const int tilesPerGroup = nTiles * mTiles;
  const int totalTiles = tilesPerGroup * g.nGroups;   // groups: same operands, own tap list + output base
  // split-K: a work item is (tile, K-slice); the slices of one tile are neighbouring items, so they
  // run at the same time on neighbouring SMs (shared activation rows and output lines stay in L2)
  const int kSplit = g.kSplit > 1 ? g.kSplit : 1;
  const int totalItems = totalTiles * kSplit;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < totalItems; item += gridDim.x) {
      const int tile = item / kSplit;
      const int ks = item - tile * kSplit;
      const int grp = tile / tilesPerGroup;
      const int tl = tile - grp * tilesPerGroup;
      const int nt = tl % nTiles;
      int mt = tl / nTiles;
      const int tx = mt % g.tilesX;
      mt /= g.tilesX;
      const int ty = mt % g.tilesY;
      const int tb = mt / g.tilesY;
      const int x0 = tx * g.BX, y0 = ty * g.BY, b0 = tb * g.BB, n0 = nt * BLOCK_N;
      const int numK = g.grpTapCount[grp] * g.cBlocks;
      const int kBeg = ks * numK / kSplit, kEnd = (ks + 1) * numK / kSplit;
      int t = g.grpTapStart[grp] + kBeg / g.cBlocks;
      int cb = kBeg % g.cBlocks;
      for (int kb = kBeg; kb < kEnd; ++kb) {
        const Tap tap = g.taps[t];
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kStageBytes;
        ptx::mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
        ptx::tma_load_5d(st, &tmAh, &full[stage], cb * kBlockK, x0 + tap.dx, y0 + tap.dy,
                         tap.plane, b0);
        ptx::tma_load_3d(st + Cfg::kABytes, &tmWh, &full[stage], cb * kBlockK, n0, tap.w);
        if (NPASS == 3) {
          uint8_t* lo = st + Cfg::kABytes + Cfg::kBBytes;
          ptx::tma_load_5d(lo, &tmAl, &full[stage], cb * kBlockK, x0 + tap.dx, y0 + tap.dy,
                           tap.plane, b0);
          ptx::tma_load_3d(lo + Cfg::kABytes, &tmWl, &full[stage], cb * kBlockK, n0, tap.w);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        if (++cb == g.cBlocks) { cb = 0; ++t; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    // kind::f16 input format bits [7,10) / [10,13): 1 = bf16, 0 = fp16 (C8H backward on fp16 planes)
    const uint32_t idesc = ptx::umma_idesc_bf16(kTileM, BLOCK_N, 0, 0) & ~((NPASS == 1 && g.half16) ? ((1u << 7) | (1u << 10)) : 0u);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < totalItems; item += gridDim.x, ++it) {
      const int tile = item / kSplit;
      const int ks = item - tile * kSplit;
      const int acc = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ptx::mbar_wait(&tempty[acc], aphase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      const int numKg = g.grpTapCount[tile / tilesPerGroup] * g.cBlocks;
      const int numK = (ks + 1) * numKg / kSplit - ks * numKg / kSplit;
      for (int kb = 0; kb < numK; ++kb) {
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
        if (kb == numK - 1) ptx::umma_commit(&tfull[acc]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;
    const float osc = half16_scale<NPASS>(g);
    int it = 0;
    for (int item = blockIdx.x; item < totalItems; item += gridDim.x, ++it) {
      const int tile = item / kSplit;
      const int ks = item - tile * kSplit;
      const int acc = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int grp = tile / tilesPerGroup;
      const int tl = tile - grp * tilesPerGroup;
      const int nt = tl % nTiles;
      int mt = tl / nTiles;
      const int tx = mt % g.tilesX;
      mt /= g.tilesX;
      const int ty = mt % g.tilesY;
      const int tb = mt / g.tilesY;
      const int n0 = nt * BLOCK_N;
      const int bx = row % g.BX;
      const int by = (row / g.BX) % g.BY;
      const int bb = row / (g.BX * g.BY);
      const int x = tx * g.BX + bx, y = ty * g.BY + by, b = tb * g.BB + bb;
      const bool valid = (x < g.oX) && (y < g.oY) && (b < g.oB);
      const long long off = g.grpOutOff[grp] + (long long)b * g.sB + (long long)y * g.sY +
                            (long long)x * g.sX + (long long)(n0 / g.nSplit) * g.sNhi + (n0 % g.nSplit);
      float* orow = g.out + off;
      const float* arow = g.addsrc ? g.addsrc + off : nullptr;

      // plain stores go through the warp's coalescing buffer (epilogue_chunk_staged)
      const bool staged = !arow && kSplit == 1;
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
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int j = 0; j < BLOCK_N / 32; ++j) {
        uint32_t v[32];
        ptx::tmem_ld32(taddr + j * 32, v);
        ptx::tmem_ld_wait();
        if (NPASS == 1 && g.half16) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(osc * __uint_as_float(v[i]));
        }
        if (staged) {
          float* rp[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) rp[i] = rowp[i] + j * 32;
          epilogue_chunk_staged(g, v, valid, sbuf, rp, vmask, n0 + j * 32, lane, b);
        } else {
          epilogue_chunk(g, v, valid, orow + j * 32, arow ? arow + j * 32 : nullptr, n0 + j * 32, lane, b,
                         kSplit > 1, ks == 0);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

static int g_force_block_n = 0;
void set_force_block_n(int n) { g_force_block_n = n; }

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static bool check_conv_geom(const ConvGeom& g, int blockN) {
  if (g.BX * g.BY * g.BB != kTileM) { set_error("conv: box %dx%dx%d != 128", g.BX, g.BY, g.BB); return false; }
  if (g.w.N % blockN) { set_error("conv: N=%d not a multiple of %d", g.w.N, blockN); return false; }
  if (g.a.C % kBlockK || g.a.C != g.w.K) { set_error("conv: C=%d K=%d must match and be multiples of 64", g.a.C, g.w.K); return false; }
  if (g.nSplit % blockN) { set_error("conv: nSplit=%d not a multiple of %d", g.nSplit, blockN); return false; }
  if (g.nTaps < 1 || g.nTaps > kMaxTaps) { set_error("conv: nTaps=%d", g.nTaps); return false; }
  if (g.cBlocks != g.a.C / kBlockK) { set_error("conv: cBlocks"); return false; }
  if (g.nPass != 1 && g.nPass != 3) { set_error("conv: nPass=%d", g.nPass); return false; }
  if (g.half16 && g.nPass != 1) { set_error("conv: half16 needs nPass = 1"); return false; }
  if (g.nGroups < 1 || g.nGroups > 4) { set_error("conv: nGroups=%d", g.nGroups); return false; }
  for (int i = 0; i < g.nGroups; ++i)
    if (g.grpTapCount[i] < 1 || g.grpTapStart[i] + g.grpTapCount[i] > g.nTaps) { set_error("conv: group %d taps", i); return false; }
  if (g.tailTiles > 0 && (g.kSplit > 1 || g.tailSplit < 2 || !g.tailScratch)) { set_error("conv: bad tail split"); return false; }
  if (g.kSplit > 1) {
    if (g.statSum) { set_error("conv: split-K cannot feed the fused statistics"); return false; }
    for (int i = 0; i < g.nGroups; ++i)
      if (g.kSplit > g.grpTapCount[i] * g.cBlocks) { set_error("conv: kSplit=%d exceeds the %d k-blocks of group %d", g.kSplit, g.grpTapCount[i] * g.cBlocks, i); return false; }
  }
  return true;
}

template <int BLOCK_N, int NPASS>
static cudaError_t launch_conv_tc_t(const ConvGeom& g, cudaStream_t stream) {
  using Cfg = ConvCfg<BLOCK_N, NPASS>;
  if (!check_conv_geom(g, BLOCK_N)) return cudaErrorInvalidValue;
  CUtensorMap tmAh, tmAl, tmWh, tmWl;
  if (!make_act_tmap(&tmAh, g.a.hi, g.a, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
  if (!make_wgt_tmap(&tmWh, g.w.hi, g.w, BLOCK_N)) return cudaErrorInvalidValue;
  if (NPASS == 3) {
    if (!make_act_tmap(&tmAl, g.a.lo, g.a, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
    if (!make_wgt_tmap(&tmWl, g.w.lo, g.w, BLOCK_N)) return cudaErrorInvalidValue;
  } else {
    tmAl = tmAh;
    tmWl = tmWh;
  }
  static bool attr_done[64] = {};   // cudaFuncSetAttribute is per device
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  bool& attr_set = attr_done[dev_id & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, NPASS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("conv: smem attr: %s", cudaGetErrorString(e)); return e; }
    attr_set = true;
  }
  const int total = (g.w.N / BLOCK_N) * g.tilesX * g.tilesY * g.tilesB * g.nGroups * (g.kSplit > 1 ? g.kSplit : 1);
  const int grid = total < num_sms() ? total : num_sms();
  profile_begin(0, g.algoFlops, stream);
  conv_tc_kernel<BLOCK_N, NPASS><<<grid, 256, Cfg::kSmemBytes, stream>>>(tmAh, tmAl, tmWh, tmWl, g);
  profile_end(stream);
  return launched();
}


// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): one 256-position x BLOCK_N tile per pair of CTAs on neighbouring
// SMs.  Each CTA stages its own 128-position activation box and HALF of the weight rows, so the
// operand bytes every SM pulls from L2 (and reads from shared memory) per MMA are halved versus
// the single-CTA kernel -- the limiter measured in profiles/r01_ncu_full_summary.md.  The leader
// CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 with M = 256; accumulator rows 0-127 live
// in the leader's TMEM, rows 128-255 in the peer's; each CTA's epilogue drains its own half.
template <int BLOCK_N, int NPASS>
struct Conv2Cfg {
  static constexpr int kABytes = kTileM * kBlockK * 2;          // 16 KB (this CTA's 128 positions)
  static constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;   // this CTA's half of the weight rows
  static constexpr int kStageBytes = (kABytes + kBBytes) * (NPASS == 3 ? 2 : 1);
  static constexpr int kOutStageBytes = 8 * kStageFloatsPerWarp * 4;   // coalescing buffers of the 8 epilogue warps
  static constexpr int kStagesRaw = (222 * 1024 - kOutStageBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kOutStageBytes;
  static_assert(kStages >= 2, "need at least two pipeline stages");
  static_assert(kTmemCols >= 32 && kTmemCols <= 512, "TMEM columns");
};

// 384 threads: eight epilogue warps (two per TMEM lane quadrant, half of the tile's columns each) so that
// short-K layers (the Generator stem: 5 k-blocks per tile) are not bound by draining the accumulator.
constexpr int kConv2Threads = 384;
template <int BLOCK_N, int NPASS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConv2Threads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                const __grid_constant__ ConvGeom g) {
  using Cfg = Conv2Cfg<BLOCK_N, NPASS>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full = bars;                      // used in the leader only (2 arrivals + 2x tx bytes)
  uint64_t* empty = bars + kStages;           // per CTA, released by the leader's multicast commit
  uint64_t* tfull = bars + 2 * kStages;       // per CTA
  uint64_t* tempty = bars + 2 * kStages + 2;  // leader only: 4 epilogue warps x 2 CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmAh);
    ptx::prefetch_tmap(&tmWh);
    if (NPASS == 3) {
      ptx::prefetch_tmap(&tmAl);
      ptx::prefetch_tmap(&tmWl);
    }
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
  ptx::cluster_sync();   // barriers of BOTH CTAs initialised before any remote arrive / multicast
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
      } else {          // odd tile count: the peer's half of the last pair is all padding (zero fill)
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
        ptx::tma_load_5d_2sm(st, &tmAh, &full[stage], cb * kBlockK, x0 + tap.dx, y0 + tap.dy,
                             tap.plane, b0);
        ptx::tma_load_3d_2sm(st + Cfg::kABytes, &tmWh, &full[stage], cb * kBlockK, n0, tap.w);
        if (NPASS == 3) {
          uint8_t* lo = st + Cfg::kABytes + Cfg::kBBytes;
          ptx::tma_load_5d_2sm(lo, &tmAl, &full[stage], cb * kBlockK, x0 + tap.dx, y0 + tap.dy,
                               tap.plane, b0);
          ptx::tma_load_3d_2sm(lo + Cfg::kABytes, &tmWl, &full[stage], cb * kBlockK, n0, tap.w);
        }
        if (!leader) ptx::mbar_arrive_remote(&full[stage], 0);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        if (++cb == g.cBlocks) { cb = 0; ++t; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    const uint32_t idesc = ptx::umma_idesc_bf16(2 * kTileM, BLOCK_N, 0, 0) & ~((NPASS == 1 && g.half16) ? ((1u << 7) | (1u << 10)) : 0u);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = pairIdx; item < totalItems; item += numPairs, ++it) {
      const ConvItem wi = conv_decode_item(g, item, totalTiles);
      const int tile = wi.tile, ks = wi.ks, kSplit = wi.nsplit;
      const int acc = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ptx::mbar_wait(&tempty[acc], aphase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      const int numKg = g.grpTapCount[tile / tilesPerGroup] * g.cBlocks;
      const int numK = (ks + 1) * numKg / kSplit - ks * numKg / kSplit;
      for (int kb = 0; kb < numK; ++kb) {
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
          ptx::umma_bf16_2cta(d_tmem, dAh, dBh, idesc, (kb | k) != 0);
          if (NPASS == 3) {
            const uint64_t dAl = ptx::umma_smem_desc_sw128(sAl + k * 32, 0, 1024);
            const uint64_t dBl = ptx::umma_smem_desc_sw128(sBl + k * 32, 0, 1024);
            ptx::umma_bf16_2cta(d_tmem, dAh, dBl, idesc, 1);
            ptx::umma_bf16_2cta(d_tmem, dAl, dBh, idesc, 1);
          }
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
    const float osc = half16_scale<NPASS>(g);
    int it = 0;
    for (int item = pairIdx; item < totalItems; item += numPairs, ++it) {
      const ConvItem wi = conv_decode_item(g, item, totalTiles);
      const int tile = wi.tile, ks = wi.ks, kSplit = wi.nsplit;
      const int acc = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
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
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int j = colHalf * kChunksPerWarp; j < (colHalf + 1) * kChunksPerWarp; ++j) {
        uint32_t v[32];
        ptx::tmem_ld32(taddr + j * 32, v);
        ptx::tmem_ld_wait();
        if (NPASS == 1 && g.half16) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(osc * __uint_as_float(v[i]));
        }
        if (wi.slot >= 0) {   // tail split: raw partial -> scratch, merged by conv_tail_fixup
          store_partial_chunk(g.tailScratch + (((size_t)wi.slot * 2 + rank) * kTileM + row) * BLOCK_N + j * 32, v);
        } else if (staged) {
          float* rp[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) rp[i] = rowp[i] + j * 32;
          epilogue_chunk_staged(g, v, valid, sbuf, rp, vmask, n0 + j * 32, lane, b);
        } else {
          epilogue_chunk(g, v, valid, orow + j * 32, arow ? arow + j * 32 : nullptr, n0 + j * 32, lane, b,
                         kSplit > 1, ks == 0);
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
  ptx::cluster_sync();   // neither CTA may retire (smem / TMEM) while its peer still uses it
  if (warp == 2) ptx::tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
}

template <int BLOCK_N, int NPASS>
static cudaError_t launch_conv_tc2_t(const ConvGeom& g, cudaStream_t stream) {
  using Cfg = Conv2Cfg<BLOCK_N, NPASS>;
  if (!check_conv_geom(g, BLOCK_N)) return cudaErrorInvalidValue;
  CUtensorMap tmAh, tmAl, tmWh, tmWl;
  if (!make_act_tmap(&tmAh, g.a.hi, g.a, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
  if (!make_wgt_tmap(&tmWh, g.w.hi, g.w, BLOCK_N / 2)) return cudaErrorInvalidValue;
  if (NPASS == 3) {
    if (!make_act_tmap(&tmAl, g.a.lo, g.a, g.BX, g.BY, g.BB)) return cudaErrorInvalidValue;
    if (!make_wgt_tmap(&tmWl, g.w.lo, g.w, BLOCK_N / 2)) return cudaErrorInvalidValue;
  } else {
    tmAl = tmAh;
    tmWl = tmWh;
  }
  static bool attr_done[64] = {};   // cudaFuncSetAttribute is per device
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  bool& attr_set = attr_done[dev_id & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BLOCK_N, NPASS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("conv2: smem attr: %s", cudaGetErrorString(e)); return e; }
    attr_set = true;
  }
  const int mTiles = g.tilesX * g.tilesY * g.tilesB;
  const int tilesAll = (g.w.N / BLOCK_N) * ((mTiles + 1) / 2) * g.nGroups;
  const int total = g.tailTiles > 0 ? tilesAll - g.tailTiles + g.tailTiles * g.tailSplit : tilesAll * (g.kSplit > 1 ? g.kSplit : 1);
  const int maxPairs = num_sms() / 2;
  const int pairs = total < maxPairs ? total : maxPairs;
  profile_begin(BLOCK_N == 256 ? kProfConv2w : kProfConv2n, g.algoFlops, stream);
  conv_tc2_kernel<BLOCK_N, NPASS><<<2 * pairs, kConv2Threads, Cfg::kSmemBytes, stream>>>(tmAh, tmAl, tmWh, tmWl, g);
  profile_end(stream);
  cudaError_t e = launched();
  if (e == cudaSuccess && g.tailTiles > 0) e = launch_conv_tail_fixup(g, BLOCK_N, stream);
  return e;
}

// CTA-pair kernel selection: on by default for layers with enough tiles to fill the chip
// (MCGVC_CTA2=0 falls back to the single-CTA kernel everywhere, for A/B measurements).
static int env_cta2() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCGVC_CTA2");
    v = e ? atoi(e) : 1;
  }
  return v;
}
static int g_force_cta2 = -1;
void set_force_cta2(int v) { g_force_cta2 = v; }

// BLOCK_N selection: widest tile that divides N and nSplit (256 halves B-operand smem traffic per
// MMA; 64 exists for the narrow data-gradient outputs of the two stem layers).

static int env_block_n() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCGVC_BLOCK_N");  // tuning override: 64 / 128 / 256
    v = e ? atoi(e) : 0;
  }
  return v;
}

// Which kernel a geometry runs on: CTA pair (M = 256) or single CTA (M = 128), and its BLOCK_N.
struct ConvVariant {
  bool pair;
  int bn;
  long long tiles;   // work items before split-K
  int slots;         // concurrent work items on the chip
};
static ConvVariant conv_variant(const ConvGeom& g) {
  ConvVariant v{};
  const int mTiles = g.tilesX * g.tilesY * g.tilesB;
  const int want = g_force_cta2 >= 0 ? g_force_cta2 : env_cta2();
  if (want && g.w.N % 128 == 0 && g.nSplit % 128 == 0) {
    int bn = (g.w.N % 256 == 0 && g.nSplit % 256 == 0) ? 256 : 128;
    if (g_force_block_n == 128) bn = 128;
    const long long pairTiles = (long long)(g.w.N / bn) * ((mTiles + 1) / 2) * g.nGroups;
    if (g_force_cta2 >= 0 || pairTiles >= num_sms() / 2) {   // small layers: 1-CTA kernel
      v.pair = true; v.bn = bn; v.tiles = pairTiles; v.slots = num_sms() / 2;
      return v;
    }
  }
  int bn = 64;
  if (g.w.N % 128 == 0 && g.nSplit % 128 == 0) bn = 128;
  // small position grids (the 1-D trunk): narrower tiles so that more SMs get a tile
  if (bn == 128 && (long long)mTiles * (g.w.N / 128) * g.nGroups < num_sms()) bn = 64;
  if (g_force_block_n == 256 && g.w.N % 256 == 0 && g.nSplit % 256 == 0) bn = 256;
  if (g_force_block_n == 64) bn = 64;
  v.pair = false; v.bn = bn; v.tiles = (long long)mTiles * (g.w.N / bn) * g.nGroups; v.slots = num_sms();
  return v;
}

static int env_ksplit() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCGVC_KSPLIT");   // 0 disables split-K (A/B measurements)
    v = e ? atoi(e) : 1;
  }
  return v;
}

// wave arithmetic shared by the planners: `tiles` work items before splitting, `slots` concurrent ones
int plan_ksplit_waves(long long tiles, int slots, const ConvGeom& g, double minGain) {
  if (!env_ksplit() || g.statSum) return 1;
  if (tiles < slots) return 1;                 // sub-wave layers are latency-bound, not wave-bound
  int minK = 1 << 30;
  for (int i = 0; i < g.nGroups; ++i) {
    const int k = g.grpTapCount[i] * g.cBlocks;
    if (k < minK) minK = k;
  }
  int maxS = minK / 4;                          // at least 4 k-blocks per slice
  if (maxS > 8) maxS = 8;
  auto cost = [&](int s) {
    const double waves = (double)(tiles * s) / slots;
    const double rounds = (double)((tiles * s + slots - 1) / slots);
    return rounds / waves + 0.01 * (s - 1);
  };
  int best = 1;
  double bestCost = cost(1);
  for (int s = 2; s <= maxS; ++s) {
    const double c = cost(s);
    if (c < bestCost - 1e-9) { bestCost = c; best = s; }
  }
  if (best > 1 && cost(1) - bestCost < minGain * cost(1)) best = 1;
  return best;
}

int conv_plan_ksplit(const ConvGeom& g, double minGain) {
  if (g_force_block_n == 0 && env_block_n() != 0) g_force_block_n = env_block_n();
  const ConvVariant v = conv_variant(g);
  return plan_ksplit_waves(v.tiles, v.slots, g, minGain);
}

// ------------------------------------------------------------------------------------------------
// Tail split (see ConvGeom::tailTiles).  With T tiles on S = 74 CTA pairs the persistent kernels run
// ceil(T / S) rounds; when the last round is mostly empty (up2 / up1 at batch 64: 320 tiles = 4.32 rounds)
// its tiles are cut into K-slices so that all pairs share it: 4 + 1/3 rounds instead of 5.  Sub-wave
// launches (small batches: 10 tiles of 100 k-blocks at batch 1) are split the same way, which turns a
// 55 us serial K loop into ~8 us.
int conv_pair_block_n(const ConvGeom& g) {
  if (g_force_block_n == 0 && env_block_n() != 0) g_force_block_n = env_block_n();
  const ConvVariant v = conv_variant(g);
  return v.pair ? v.bn : 0;
}
static int env_tail_split() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MCGVC_TAIL_SPLIT"); v = e ? atoi(e) : 1; }
  return v;
}
bool conv_plan_tail(ConvGeom& g, int blockN, float* scratch) {
  g.tailTiles = 0; g.tailSplit = 0; g.tailScratch = nullptr;
  if (!env_tail_split() || !scratch || g.kSplit > 1) return false;
  const int bn = blockN ? blockN : conv_pair_block_n(g);
  if (bn == 0 || g.w.N % bn || g.nSplit % bn) return false;
  const int slots = num_sms() / 2;
  if (slots > 74) return false;                      // kTailScratchFloats is sized for 74 slice slabs (148 SMs)
  const int mTiles = g.tilesX * g.tilesY * g.tilesB;
  const long long T = (long long)(g.w.N / bn) * ((mTiles + 1) / 2) * g.nGroups;
  int minK = 1 << 30;
  for (int i = 0; i < g.nGroups; ++i) {
    const int k = g.grpTapCount[i] * g.cBlocks;
    if (k < minK) minK = k;
  }
  if (minK < 16 || g.BX * g.BY < 16) return false;   // short K loops: the fix-up pass would cost more than it saves
  if (g.nGroups > 1) return false;                   // grouped stride-2 data gradients: the last group's short tap list
                                                     // leaves slices of a few k-blocks (measured slower: 216 -> 262 us)
  const long long rem = T % slots;
  if (rem == 0) return false;
  int s = (int)(slots / rem);
  if (s > 8) s = 8;
  if (s > minK / 8) s = minK / 8;                    // at least 8 k-blocks per slice
  if (s < 2) return false;
  g.tailTiles = (int)rem;
  g.tailSplit = s;
  g.tailScratch = scratch;
  return true;
}

constexpr int kFixRows = 16;   // rows of a half tile per fix-up block (an image's BX*BY rows are a multiple of it)
template <int BLOCK_N>
__global__ void __launch_bounds__(256) conv_tail_fixup_kernel(const __grid_constant__ ConvGeom g) {
  // grid (tail tile, CTA half, 16-row chunk); thread = (4-column group, row group)
  constexpr int kColGroups = BLOCK_N / 4;
  constexpr int kRowsPerPass = 256 / kColGroups;
  __shared__ float4 red[2][256];
  const int nTiles = g.w.N / BLOCK_N;
  const int mTiles = g.tilesX * g.tilesY * g.tilesB;
  const int pairM = (mTiles + 1) / 2;
  const int tilesPerGroup = nTiles * pairM;
  const int totalTiles = tilesPerGroup * g.nGroups;
  const int tl_ = blockIdx.x, rank = blockIdx.y;
  const int tile = totalTiles - g.tailTiles + tl_;
  const int grp = tile / tilesPerGroup;
  const int tl = tile - grp * tilesPerGroup;
  const int nt = tl % nTiles;
  int mt = (tl / nTiles) * 2 + rank;
  if (mt >= mTiles) return;                          // odd tile count: the peer half of the last pair is padding
  const int tx = mt % g.tilesX;
  mt /= g.tilesX;
  const int ty = mt % g.tilesY;
  const int tb = mt / g.tilesY;
  const int n0 = nt * BLOCK_N;
  const int cg = threadIdx.x % kColGroups, rg = threadIdx.x / kColGroups;
  const int c = cg * 4;
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g.bias) bv = __ldg(reinterpret_cast<const float4*>(g.bias + n0 + c));
  const int rowsPerImg = g.BX * g.BY;                // a half tile holds BB images of BX*BY (>= 16) positions each
  const int row0 = blockIdx.z * kFixRows;
  const int b = tb * g.BB + row0 / rowsPerImg;       // all 16 rows of this block belong to one image
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  for (int row = row0 + rg; row < row0 + kFixRows; row += kRowsPerPass) {
    const int ri = row % rowsPerImg;
    const int bx = ri % g.BX, by = ri / g.BX;
    const int x = tx * g.BX + bx, y = ty * g.BY + by;
    if (x >= g.oX || y >= g.oY || b >= g.oB) continue;
    float4 acc = bv;
    for (int s = 0; s < g.tailSplit; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(
          g.tailScratch + ((((size_t)tl_ * g.tailSplit + s) * 2 + rank) * kTileM + row) * BLOCK_N + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const long long off = g.grpOutOff[grp] + (long long)b * g.sB + (long long)y * g.sY + (long long)x * g.sX +
                          (long long)((n0 + c) / g.nSplit) * g.sNhi + ((n0 + c) % g.nSplit);
    if (g.addsrc) {
      const float4 av = *reinterpret_cast<const float4*>(g.addsrc + off);
      acc.x += av.x; acc.y += av.y; acc.z += av.z; acc.w += av.w;
    }
    *reinterpret_cast<float4*>(g.out + off) = acc;
    s1.x += acc.x; s1.y += acc.y; s1.z += acc.z; s1.w += acc.w;
    s2.x = fmaf(acc.x, acc.x, s2.x); s2.y = fmaf(acc.y, acc.y, s2.y); s2.z = fmaf(acc.z, acc.z, s2.z); s2.w = fmaf(acc.w, acc.w, s2.w);
  }
  if (g.statSum) {                                    // uniform branch: every thread of the block takes it
    red[0][threadIdx.x] = s1;
    red[1][threadIdx.x] = s2;
    __syncthreads();
    if (rg == 0 && b < g.oB) {
      for (int k = 1; k < kRowsPerPass; ++k) {
        const float4 u = red[0][k * kColGroups + cg], q = red[1][k * kColGroups + cg];
        s1.x += u.x; s1.y += u.y; s1.z += u.z; s1.w += u.w;
        s2.x += q.x; s2.y += q.y; s2.z += q.z; s2.w += q.w;
      }
      float* dA = g.statSum + (long long)b * g.w.N + n0 + c;
      float* dB = g.statSq + (long long)b * g.w.N + n0 + c;
      atomicAdd(dA + 0, s1.x); atomicAdd(dA + 1, s1.y); atomicAdd(dA + 2, s1.z); atomicAdd(dA + 3, s1.w);
      atomicAdd(dB + 0, s2.x); atomicAdd(dB + 1, s2.y); atomicAdd(dB + 2, s2.z); atomicAdd(dB + 3, s2.w);
    }
  }
}
cudaError_t launch_conv_tail_fixup(const ConvGeom& g, int blockN, cudaStream_t stream) {
  dim3 grid(g.tailTiles, 2, kTileM / kFixRows);
  if (blockN == 256) conv_tail_fixup_kernel<256><<<grid, 256, 0, stream>>>(g);
  else if (blockN == 128) conv_tail_fixup_kernel<128><<<grid, 256, 0, stream>>>(g);
  else { set_error("conv tail fix-up: blockN=%d", blockN); return cudaErrorInvalidValue; }
  return launched();
}

cudaError_t launch_conv_tc(const ConvGeom& g, cudaStream_t stream) {
  if (g_force_block_n == 0 && env_block_n() != 0) g_force_block_n = env_block_n();
  const ConvVariant v = conv_variant(g);
  if (v.pair) {
    if (g.nPass == 3) return v.bn == 256 ? launch_conv_tc2_t<256, 3>(g, stream) : launch_conv_tc2_t<128, 3>(g, stream);
    return v.bn == 256 ? launch_conv_tc2_t<256, 1>(g, stream) : launch_conv_tc2_t<128, 1>(g, stream);
  }
  if (g.nPass == 3) {
    if (v.bn == 256) return launch_conv_tc_t<256, 3>(g, stream);
    if (v.bn == 128) return launch_conv_tc_t<128, 3>(g, stream);
    return launch_conv_tc_t<64, 3>(g, stream);
  }
  if (v.bn == 256) return launch_conv_tc_t<256, 1>(g, stream);
  if (v.bn == 128) return launch_conv_tc_t<128, 1>(g, stream);
  return launch_conv_tc_t<64, 1>(g, stream);
}

// ------------------------------------------------------------------------------------------------
// SIMT checking kernel: same operands, same tap table, one thread per output element.  It exists so
// the tcgen05 kernel can be validated against something that shares no descriptor/swizzle logic
// with it; it is not used by the network path unless MCGVC_BACKEND=simt is set for debugging.
__device__ __forceinline__ float bf16_bits_to_f(uint16_t v) {
  return __uint_as_float(static_cast<uint32_t>(v) << 16);
}

__global__ void conv_simt_kernel(const __grid_constant__ ConvGeom g) {
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
  const uint16_t* Ah = reinterpret_cast<const uint16_t*>(g.a.hi);
  const uint16_t* Al = reinterpret_cast<const uint16_t*>(g.a.lo);
  const uint16_t* Wh = reinterpret_cast<const uint16_t*>(g.w.hi);
  const uint16_t* Wl = reinterpret_cast<const uint16_t*>(g.w.lo);
  float acc = 0.f;
  for (int t = g.grpTapStart[grp]; t < g.grpTapStart[grp] + g.grpTapCount[grp]; ++t) {
    const Tap tap = g.taps[t];
    const int xx = x + tap.dx, yy = y + tap.dy;
    if (xx < 0 || xx >= g.a.X || yy < 0 || yy >= g.a.Y) continue;
    const long long aoff = ((((long long)b * g.a.P + tap.plane) * g.a.Y + yy) * g.a.X + xx) * g.a.C;
    const long long woff = ((long long)tap.w * g.w.N + n) * g.w.K;
    for (int c = 0; c < g.a.C; ++c) {
      const float ah = g.half16 ? __half2float(__ushort_as_half(Ah[aoff + c])) : bf16_bits_to_f(Ah[aoff + c]);
      const float wh = g.half16 ? __half2float(__ushort_as_half(Wh[woff + c])) : bf16_bits_to_f(Wh[woff + c]);
      acc = fmaf(ah, wh, acc);
      if (g.nPass == 3) {
        const float al = bf16_bits_to_f(Al[aoff + c]);
        const float wl = bf16_bits_to_f(Wl[woff + c]);
        acc = fmaf(ah, wl, acc);
        acc = fmaf(al, wh, acc);
      }
    }
  }
  const long long off = g.grpOutOff[grp] + (long long)b * g.sB + (long long)y * g.sY + (long long)x * g.sX +
                        (long long)(n / g.nSplit) * g.sNhi + (n % g.nSplit);
  if (g.half16) {
    float sc = g.c8OutScale;
    if (g.c8RecA && g.c8RecW) sc *= g.c8RecA[0] * g.c8RecW[0];
    acc *= sc;
  }
  if (g.bias) acc += g.bias[n];
  if (g.addsrc) acc += g.addsrc[off];
  g.out[off] = acc;
}

cudaError_t launch_conv_simt(const ConvGeom& g, cudaStream_t stream) {
  if (!check_conv_geom(g, 64)) return cudaErrorInvalidValue;
  const long long total = (long long)g.tilesX * g.tilesY * g.tilesB * kTileM * g.w.N * g.nGroups;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  conv_simt_kernel<<<(unsigned)blocks, threads, 0, stream>>>(g);
  return launched();
}

}  // namespace mcgvc
