// Kernel-level C entry points used by tests/ to check the tensor-core kernels in isolation
// (tcgen05 path against the SIMT checking kernels and against torch on the same operands).
#include "../../include/mcgvc.h"
#include "gemm_types.cuh"

using namespace mcgvc;

namespace mcgvc { void set_force_block_n(int n); void set_force_cta2(int v); void set_force_wgrad_cta2(int v); }

static int g_debug_ksplit = 1;

extern "C" {

/* split-K factor the next mcgvc_debug_conv calls use on the tensor-core backends (the caller
 * zero-fills `out`; bias / addsrc ride on the first slice); 1 = off. */
int mcgvc_debug_set_conv_ksplit(int k) {
  g_debug_ksplit = k < 1 ? 1 : k;
  return 0;
}

/* Split-K factor the planner picks for a convolution geometry on the split-bf16 kernels (host logic only:
 * works without a GPU, assuming 148 SMs). */
int mcgvc_debug_plan_ksplit(int oB, int oY, int oX, int C, int N, int nSplit, int nTaps, double minGain) {
  ConvGeom g{};
  g.a.C = C; g.w.K = C; g.w.N = N;
  g.oX = oX; g.oY = oY; g.oB = oB;
  if (!choose_box(oB, oY, oX, kTileM, &g.BX, &g.BY, &g.BB)) return -1;
  g.tilesX = (oX + g.BX - 1) / g.BX;
  g.tilesY = (oY + g.BY - 1) / g.BY;
  g.tilesB = (oB + g.BB - 1) / g.BB;
  g.nTaps = nTaps; g.cBlocks = C / kBlockK;
  g.nGroups = 1; g.grpTapStart[0] = 0; g.grpTapCount[0] = nTaps;
  g.nSplit = nSplit;
  return conv_plan_ksplit(g, minGain);
}

/* Tail-split plan for the same kind of geometry on a pair kernel of tile width blockN (host logic only):
 * returns the number of K-slices per tail tile (0 = no tail split) and stores the number of tail tiles. */
int mcgvc_debug_plan_tail(int oB, int oY, int oX, int C, int N, int nSplit, int nTaps, int blockN, int* tail_tiles) {
  ConvGeom g{};
  g.a.C = C; g.w.K = C; g.w.N = N;
  g.oX = oX; g.oY = oY; g.oB = oB;
  if (!choose_box(oB, oY, oX, kTileM, &g.BX, &g.BY, &g.BB)) return -1;
  g.tilesX = (oX + g.BX - 1) / g.BX;
  g.tilesY = (oY + g.BY - 1) / g.BY;
  g.tilesB = (oB + g.BB - 1) / g.BB;
  g.nTaps = nTaps; g.cBlocks = C / kBlockK;
  g.nGroups = 1; g.grpTapStart[0] = 0; g.grpTapCount[0] = nTaps;
  g.nSplit = nSplit;
  static float dummy;                       // the planner only needs a non-null scratch pointer
  const bool on = conv_plan_tail(g, blockN, &dummy);
  if (tail_tiles) *tail_tiles = on ? g.tailTiles : 0;
  return on ? g.tailSplit : 0;
}

int mcgvc_debug_conv(const void* a_hi, const void* a_lo, int aC, int aX, int aY, int aP, int aB,
                     const void* w_hi, const void* w_lo, int wK, int wN, int wT, int oX, int oY,
                     int oB, int nTaps, const int8_t* taps4, float* out, long long sB,
                     long long sY, long long sX, int nSplit, long long sNhi, const float* bias,
                     const float* addsrc, int nPass, int backend, int blockN, void* stream) {
  ConvGeom g{};
  g.a = ActOperand{a_hi, a_lo, aC, aX, aY, aP, aB};
  g.w = WgtOperand{w_hi, w_lo, wK, wN, wT};
  g.oX = oX; g.oY = oY; g.oB = oB;
  if (!choose_box(oB, oY, oX, kTileM, &g.BX, &g.BY, &g.BB)) { set_error("choose_box failed"); return 1; }
  g.tilesX = (oX + g.BX - 1) / g.BX;
  g.tilesY = (oY + g.BY - 1) / g.BY;
  g.tilesB = (oB + g.BB - 1) / g.BB;
  if (nTaps > kMaxTaps) { set_error("too many taps"); return 1; }
  g.nTaps = nTaps;
  g.cBlocks = aC / kBlockK;
  for (int t = 0; t < nTaps; ++t) {
    g.taps[t].dx = taps4[4 * t + 0];
    g.taps[t].dy = taps4[4 * t + 1];
    g.taps[t].plane = (uint8_t)taps4[4 * t + 2];
    g.taps[t].w = (uint8_t)taps4[4 * t + 3];
  }
  g.nGroups = 1; g.grpTapStart[0] = 0; g.grpTapCount[0] = nTaps; g.grpOutOff[0] = 0;
  g.sB = sB; g.sY = sY; g.sX = sX; g.nSplit = nSplit; g.sNhi = sNhi;
  g.out = out; g.bias = bias; g.addsrc = addsrc; g.nPass = nPass;
  g.kSplit = backend == 1 ? 1 : g_debug_ksplit;
  // backend: 0 = tcgen05 single-CTA, 1 = SIMT checker, 2 = tcgen05 CTA-pair (cta_group::2)
  set_force_block_n(blockN);
  set_force_cta2(backend == 2 ? 1 : 0);
  cudaError_t e = backend == 1 ? launch_conv_simt(g, (cudaStream_t)stream)
                               : launch_conv_tc(g, (cudaStream_t)stream);
  set_force_block_n(0);
  set_force_cta2(-1);
  if (e != cudaSuccess) {
    if (!last_error()[0]) set_error("conv launch: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

/* C8 scheme (conv_c8.cu): a16/w16 are fp16 (main_bf16 = 0) or bf16 planes, a8h/a8l/w8h/w8l e4m3 planes;
 * out = out_scale * (sum a16*w16 + corr_scale * sum (a8h*w8l + a8l*w8h)) + bias + addsrc.
 * backend: 1 = SIMT checker, 2 = tcgen05 CTA-pair kernel. */
int mcgvc_debug_conv_c8(const void* a16, const void* a8h, const void* a8l, int aC, int aX, int aY, int aP,
                        int aB, const void* w16, const void* w8h, const void* w8l, int wK, int wN, int wT,
                        int oX, int oY, int oB, int nTaps, const int8_t* taps4, float* out,
                        const float* bias, const float* addsrc, int main_bf16, float out_scale,
                        float corr_scale, int backend, int blockN, void* stream) {
  ConvGeom g{};
  g.a = ActOperand{a16, nullptr, aC, aX, aY, aP, aB, a8h, a8l};
  g.w = WgtOperand{w16, nullptr, wK, wN, wT, w8h, w8l};
  g.oX = oX; g.oY = oY; g.oB = oB;
  if (!choose_box(oB, oY, oX, kTileM, &g.BX, &g.BY, &g.BB)) { set_error("choose_box failed"); return 1; }
  g.tilesX = (oX + g.BX - 1) / g.BX;
  g.tilesY = (oY + g.BY - 1) / g.BY;
  g.tilesB = (oB + g.BB - 1) / g.BB;
  if (nTaps > kMaxTaps) { set_error("too many taps"); return 1; }
  g.nTaps = nTaps;
  g.cBlocks = aC / kBlockK;
  for (int t = 0; t < nTaps; ++t) {
    g.taps[t].dx = taps4[4 * t + 0];
    g.taps[t].dy = taps4[4 * t + 1];
    g.taps[t].plane = (uint8_t)taps4[4 * t + 2];
    g.taps[t].w = (uint8_t)taps4[4 * t + 3];
  }
  g.nGroups = 1; g.grpTapStart[0] = 0; g.grpTapCount[0] = nTaps; g.grpOutOff[0] = 0;
  g.sB = (long long)oY * oX * wN; g.sY = (long long)oX * wN; g.sX = wN; g.nSplit = wN; g.sNhi = 0;
  g.out = out; g.bias = bias; g.addsrc = addsrc; g.nPass = 2; g.kSplit = 1;
  g.mainBf16 = main_bf16; g.c8OutScale = out_scale; g.c8CorrScale = corr_scale;
  g.algoFlops = 2.0 * oB * oY * oX * (double)wN * nTaps * aC;
  cudaError_t e = backend == 1 ? launch_conv_c8_simt(g, (cudaStream_t)stream)
                               : launch_conv_c8(g, blockN, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    if (!last_error()[0]) set_error("conv_c8 launch: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int mcgvc_debug_wgrad(const void* z_hi, const void* z_lo, int zC, int zX, int zY, int zB,
                      const void* x_hi, const void* x_lo, int xC, int xX, int xY, int xP, int xB,
                      int pX, int pY, int pB, int nTaps, const int8_t* taps4,
                      const int8_t* ztaps4, float* dw, int cTile, int splitK, int nPass,
                      int backend, void* stream) {
  WgradGeom g{};
  g.dz = ActOperand{z_hi, z_lo, zC, zX, zY, 1, zB};
  g.x = ActOperand{x_hi, x_lo, xC, xX, xY, xP, xB};
  g.pX = pX; g.pY = pY; g.pB = pB;
  if (!choose_box(pB, pY, pX, 64, &g.BX, &g.BY, &g.BB)) { set_error("choose_box failed"); return 1; }
  g.tilesX = (pX + g.BX - 1) / g.BX;
  g.tilesY = (pY + g.BY - 1) / g.BY;
  g.tilesB = (pB + g.BB - 1) / g.BB;
  if (nTaps > kMaxTaps) { set_error("too many taps"); return 1; }
  g.nTaps = nTaps;
  for (int t = 0; t < nTaps; ++t) {
    g.taps[t].dx = taps4[4 * t + 0];
    g.taps[t].dy = taps4[4 * t + 1];
    g.taps[t].plane = (uint8_t)taps4[4 * t + 2];
    g.taps[t].w = (uint8_t)taps4[4 * t + 3];
    g.ztaps[t].dx = ztaps4[4 * t + 0];
    g.ztaps[t].dy = ztaps4[4 * t + 1];
    g.ztaps[t].plane = 0;
    g.ztaps[t].w = 0;
  }
  g.N = zC; g.C = xC; g.cTile = cTile; g.splitK = splitK; g.dw = dw; g.nPass = nPass;
  set_force_wgrad_cta2(backend == 2 ? 1 : 0);
  cudaError_t e = backend == 1 ? launch_wgrad_simt(g, (cudaStream_t)stream)
                               : launch_wgrad_tc(g, (cudaStream_t)stream);
  set_force_wgrad_cta2(-1);
  if (e != cudaSuccess) {
    if (!last_error()[0]) set_error("wgrad launch: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int mcgvc_debug_wgrad_c8(const void* z16, const void* z8h, const void* z8l, int zC, int zX, int zY, int zB,
                         const void* x16, const void* x8h, const void* x8l, int xC, int xX, int xY, int xP,
                         int xB, int pX, int pY, int pB, int nTaps, const int8_t* taps4, const int8_t* ztaps4,
                         float* dw, int cTile, int splitK, int main_bf16, float out_scale, float corr_scale,
                         int backend, void* stream) {
  WgradGeom g{};
  g.dz = ActOperand{z16, nullptr, zC, zX, zY, 1, zB, z8h, z8l};
  g.x = ActOperand{x16, nullptr, xC, xX, xY, xP, xB, x8h, x8l};
  g.pX = pX; g.pY = pY; g.pB = pB;
  if (!choose_box(pB, pY, pX, 64, &g.BX, &g.BY, &g.BB)) { set_error("choose_box failed"); return 1; }
  g.tilesX = (pX + g.BX - 1) / g.BX;
  g.tilesY = (pY + g.BY - 1) / g.BY;
  g.tilesB = (pB + g.BB - 1) / g.BB;
  if (nTaps > kMaxTaps) { set_error("too many taps"); return 1; }
  g.nTaps = nTaps;
  for (int t = 0; t < nTaps; ++t) {
    g.taps[t].dx = taps4[4 * t + 0];
    g.taps[t].dy = taps4[4 * t + 1];
    g.taps[t].plane = (uint8_t)taps4[4 * t + 2];
    g.taps[t].w = (uint8_t)taps4[4 * t + 3];
    g.ztaps[t].dx = ztaps4[4 * t + 0];
    g.ztaps[t].dy = ztaps4[4 * t + 1];
    g.ztaps[t].plane = 0;
    g.ztaps[t].w = 0;
  }
  g.N = zC; g.C = xC; g.cTile = cTile; g.splitK = splitK; g.dw = dw; g.nPass = 2;
  g.mainBf16 = main_bf16; g.c8OutScale = out_scale; g.c8CorrScale = corr_scale;
  g.algoFlops = 2.0 * pB * pY * pX * (double)zC * xC * nTaps;
  cudaError_t e = backend == 1 ? launch_wgrad_c8_simt(g, (cudaStream_t)stream) : launch_wgrad_c8(g, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    if (!last_error()[0]) set_error("wgrad_c8 launch: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

const char* mcgvc_last_error(void) { return last_error(); }

}  // extern "C"
