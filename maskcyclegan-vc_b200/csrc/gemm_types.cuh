// Shared descriptions of the two tensor-core kernels (implicit-GEMM conv and weight-gradient GEMM).
// The same structs drive the tcgen05/TMA kernels and the plain SIMT checking kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mcgvc {

constexpr int kMaxTaps = 80;   // 5x15 head/stem = 75 taps is the largest filter on the path
constexpr int kTileM = 128;    // output positions per tile (UMMA M)
constexpr int kBlockK = 64;    // channels per k-block = one 128-byte swizzle row of bf16

// One filter tap: where to read the activation (offset in the operand's X/Y grid and which
// parity plane for stride-2 layers) and which weight slice it multiplies.
struct Tap {
  int8_t dx, dy;
  uint8_t plane, w;
};

// Activation operand: bf16 hi/lo planes laid out [B][P][Y][X][C] (C contiguous).
struct ActOperand {
  const void* hi;
  const void* lo;
  int C, X, Y, P, B;
  // C8 scheme only (conv_c8.cu): e4m3 planes of hi * 2^u and (v - hi) * 2^(u+11), same indexing, 1 B/elem
  const void* h8;
  const void* l8;
  const float* rec;  // device-side {1/S, 1/E}: S scales the 16-bit plane, E the e4m3 planes
};
// Weight operand: bf16 hi/lo, [T][N][K] (K contiguous, K = channels per tap).
struct WgtOperand {
  const void* hi;
  const void* lo;
  int K, N, T;
  const void* h8;   // C8 scheme only, see ActOperand
  const void* l8;
  const float* rec;
};

// Implicit-GEMM convolution:  out[row(b,y,x), n] = bias[n] + addsrc[..] +
//     sum_t sum_c A[b, plane_t, y+dy_t, x+dx_t, c] * W[w_t][n][c]
// Rows are enumerated over output positions (oB, oY, oX) in boxes of BB x BY x BX = 128.
struct ConvGeom {
  ActOperand a;
  WgtOperand w;
  int oX, oY, oB;
  int BX, BY, BB;
  int tilesX, tilesY, tilesB;
  int nTaps, cBlocks;
  Tap taps[kMaxTaps];
  // tile groups: up to 4 convolutions over the same operands and position grid that differ only in
  // their tap sub-list and output base (the four parity planes of a stride-2 data gradient) run as
  // ONE launch so that their tiles fill the SMs together; nGroups = 1 for ordinary convolutions
  int nGroups;
  int grpTapStart[4], grpTapCount[4];
  long long grpOutOff[4];
  // output addressing, in floats
  long long sB, sY, sX;
  int nSplit;          // column n lives at (n / nSplit) * sNhi + (n % nSplit)
  long long sNhi;
  float* out;
  const float* bias;    // [N] or null
  const float* addsrc;  // same addressing as out, or null
  int nPass;            // 1 = bf16 (hi only), 3 = split-bf16 (hi*hi + hi*lo + lo*hi)
  double algoFlops;     // algorithmic FLOPs of this launch (real conv MACs x 2), for profiling
  // optional fused InstanceNorm statistics: per-(image, column) sums of z and z^2, [oB][N] each
  float* statSum;
  float* statSq;
  int statSeg;          // lanes of an epilogue warp that share an image: min(32, BX*BY)
  // split-K: every tile's k-blocks (taps x channel blocks) are cut into kSplit slices that run as
  // separate work items and are ADDED into `out` (red.global.add) -- the caller zero-fills `out`
  // first; incompatible with the fused statistics.  0 / 1 = off.  See conv_plan_ksplit().
  int kSplit;
  // C8 scheme (conv_c8.cu): out = c8OutScale * (D1 + c8CorrScale * D2); 16-bit planes are fp16 unless mainBf16
  int mainBf16;
  float c8OutScale, c8CorrScale;
  const float* c8RecA;  // device-side scale records {1/S, 1/E} of the two operands (null: host multipliers only)
  const float* c8RecW;
  // Tail split (pair kernels only, kSplit <= 1): the LAST tailTiles tiles -- the partial wave that would
  // otherwise occupy a few CTA pairs for a whole extra round -- are each cut into tailSplit K-slices that
  // run as separate work items; a slice stores its raw partial accumulator to tailScratch
  // ([slot][cta][128 rows][BLOCK_N] fp32) and conv_tail_fixup adds the slices, the bias / residual and the
  // fused statistics.  0 = off.  See conv_plan_tail().
  int tailTiles, tailSplit;
  float* tailScratch;
  // "C8H" backward (nPass = 1 on C8 operand planes): ONE fp16 MMA per MAC on the 16-bit planes only,
  // out = c8OutScale * recA[0] * recW[0] * D; the e4m3 planes are not read.
  int half16;
};

// Weight-gradient GEMM:  dW[w_t][n][c] += sum over positions (b,y,x) in this CTA's K-slice of
//     dz[b, y+dyz_t, x+dxz_t, n] * x[b, plane_t, y+dy_t, x+dx_t, c]
// Positions are enumerated over (pB, pY, pX) in boxes of BB x BY x BX = 64 (one k-block).
struct WgradGeom {
  ActOperand dz;   // channels = N
  ActOperand x;    // channels = C
  int pX, pY, pB;
  int BX, BY, BB;
  int tilesX, tilesY, tilesB;
  int nTaps;
  Tap taps[kMaxTaps];    // offsets applied to the x operand
  Tap ztaps[kMaxTaps];   // offsets applied to the dz operand (dx, dy only)
  int N, C;              // dW slice is [N][C] per tap
  int cTile;             // columns of C per CTA (64, 128 or 256)
  int splitK;
  float* dw;             // [T][N][C] fp32, accumulated with atomics
  int nPass;
  double algoFlops;
  // C8 scheme (wgrad_c8.cu): dW += c8OutScale * (D1 + c8CorrScale * D2)
  int mainBf16;
  float c8OutScale, c8CorrScale;
  const float* c8RecZ;   // device-side scale records {1/S, 1/E} of dz and x (null: host multipliers only)
  const float* c8RecX;
  int half16;            // nPass = 1 on the fp16 planes of C8 operands, see ConvGeom::half16
  // Tap pairing (single-pass pair kernel only, all ztaps zero): one work item stages a dz k-block ONCE and
  // multiplies it with the x boxes of TWO filter taps (two TMEM accumulators): 48 KB instead of 64 KB of
  // operands per 8 MMAs, which takes the single-pass kernel from ingest-bound to MMA-bound.
  int tapPair;
};

cudaError_t launch_conv_tc(const ConvGeom& g, cudaStream_t stream);
// Split-K factor that fills whole waves of SMs best for this geometry (1 = leave it alone).  Layers
// whose tile count is a little over one wave (80 CTA-pair tiles on 74 pairs) otherwise pay two full
// rounds for 1.08 rounds of work.  `minGain` = required relative time saving (the zero-fill and the
// atomic merge are not free).
int conv_plan_ksplit(const ConvGeom& g, double minGain);
int plan_ksplit_waves(long long tiles, int slots, const ConvGeom& g, double minGain);
// Tail split planner + fix-up launch for the pair kernels.  blockN = the C8 kernel's tile width, or 0 for the
// split-bf16 / fp16 pair kernel (width chosen by its own variant logic).  Returns true and fills
// g.tailTiles / g.tailSplit when the geometry profits; scratch must hold kTailScratchFloats floats.
constexpr long long kTailScratchFloats = 74LL * 2 * 128 * 256;
bool conv_plan_tail(ConvGeom& g, int blockN, float* scratch);
cudaError_t launch_conv_tail_fixup(const ConvGeom& g, int blockN, cudaStream_t stream);
int conv_pair_block_n(const ConvGeom& g);   // 0 when the geometry runs on the single-CTA kernel
cudaError_t launch_conv_simt(const ConvGeom& g, cudaStream_t stream);
// 16-bit main pass + two e4m3 correction passes (CTA-pair kernel, blockN 128 or 256) and its SIMT checker
cudaError_t launch_conv_c8(const ConvGeom& g, int blockN, cudaStream_t stream);
cudaError_t launch_conv_c8_simt(const ConvGeom& g, cudaStream_t stream);
cudaError_t launch_wgrad_c8(const WgradGeom& g, cudaStream_t stream);
cudaError_t launch_wgrad_c8_simt(const WgradGeom& g, cudaStream_t stream);
cudaError_t launch_wgrad_tc(const WgradGeom& g, cudaStream_t stream);
cudaError_t launch_wgrad_simt(const WgradGeom& g, cudaStream_t stream);

// Fill tile counts / box shape for a position grid (B, Y, X): picks BX*BY*BB == boxPositions
// minimising padded work. Returns false if impossible.
bool choose_box(int B, int Y, int X, int boxPositions, int* BX, int* BY, int* BB);

// Every kernel launch in the library goes through launched(): counts it and returns the launch status.
cudaError_t launched();
long long launch_count();

// Optional per-launch CUDA-event timing of the two tensor-core kernels (bench.py roofline).
struct KernelProfile {
  double ms;          // summed launch durations
  double flops;       // summed algorithmic FLOPs
  long long launches;
};
void profile_enable(bool on);
bool profile_enabled();
// kinds: one per tensor-core kernel family, so that bench.py can name the DOMINANT kernel of a step
enum ProfileKind {
  kProfConv1 = 0,      // conv_tc_kernel (single-CTA tiles: small layers)
  kProfWgrad = 1,      // wgrad_tc_kernel / wgrad_tc2_kernel (split-bf16 / fp16)
  kProfConvC8w = 2,    // conv_c8_kernel<256>
  kProfConvC8n = 3,    // conv_c8_kernel<128>
  kProfConv2w = 4,     // conv_tc2_kernel<256, *>
  kProfConv2n = 5,     // conv_tc2_kernel<128, *>
  kProfTrunk = 6,      // trunk_fwd_kernel / trunk_bwd_kernel
  kProfWgradC8 = 7,    // wgrad_c8_kernel
  kProfKinds = 8
};
const char* profile_kind_name(int kind);
void profile_begin(int kind, double flops, cudaStream_t s);
void profile_end(cudaStream_t s);
void profile_collect(KernelProfile* conv, KernelProfile* wgrad);  // all conv kinds / all wgrad kinds; synchronises, then resets
void profile_collect_kinds(KernelProfile* out /* [kProfKinds] */);

const char* last_error();
void set_error(const char* fmt, ...);

}  // namespace mcgvc
