// Bandwidth-bound layer kernels around the tensor-core convolutions: operand preparation,
// InstanceNorm statistics, normalise + gate/swish (+ PixelShuffle, + residual) forward and
// backward, the 5x15 / 1x3 single-output-channel heads, weight packing and gradient unpacking.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mcgvc {

// Activation tensor in engine layout.  Logical index (img, y, x, c); memory is either plain
// [img][Y][X][C] or parity-split [img][4][ceil(Y/2)][ceil(X/2)][C] (plane = (y&1)*2 + (x&1)), the
// form a stride-2 convolution consumes through TMA without element strides.
// How a value v is written into an operand's planes.  c8 = 0: split-bf16 (hi = bf16(v), lo =
// bf16(v - hi)).  c8 = 1 (conv_c8.cu / wgrad_c8.cu): hi = fp16(v*S); the `lo` storage (2 B/elem) holds
// two e4m3 planes of `elems` bytes each: e4m3(hi*E) and e4m3((v*S - hi)*E*2^11).  c8 = 2: the fp16 plane
// only (dz tensors of the C8H mode, whose backward GEMMs run a single fp16 pass).
struct PlaneFmt {
  int c8;
  float S, E;
  long long elems;
};
struct ActBuf {
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  float* f32;
  int nImg, Y, X, C;
  int parity;
  PlaneFmt fmt;
};

enum ApplyMode {
  kGatedNoNorm = 0,     // a = z[c] * sigmoid(z[C+c])                      (G stem, model.py:242)
  kGatedIN = 1,         // a = IN(z[c]) * sigmoid(IN(z[C+c]))              (model.py:101-103, :71-74)
  kINOnly = 2,          // a = IN(z[c]) (+ residual)                       (model.py:255, :75-76, :267)
  kINSwish = 3,         // y = IN(z[c]); a = y*sigmoid(y)                  (D blocks, model.py:330-337)
  kINSwishShuffle = 5,  // PixelShuffle(2) then IN then swish              (model.py:226-237)
};

struct ApplyArgs {
  int mode;
  const float* z;       // raw conv output, rows (img, zy, zx), Nz columns
  int Nz, zY, zX;
  const float* mean;    // [nImg][Nstat]
  const float* rstd;
  int Nstat;
  const float* gamma;   // engine order, index (img % affPeriod) * Nstat + s
  const float* beta;
  int affPeriod;
  const float* residual;  // plain fp32 [nImg][Y][X][C] or null
  ActBuf out;
};

struct ApplyBwdArgs {
  int mode;
  const float* z;
  int Nz, zY, zX;
  const float* mean;
  const float* rstd;
  int Nstat;
  const float* gamma;
  const float* beta;
  int affPeriod;
  ActBuf dA;            // gradient w.r.t. the layer's activation (fp32 in .f32, same layout as fwd out)
  float* t1;            // [nImg][Nstat] sum dy        (written by reduce, read by apply)
  float* t2;            // [nImg][Nstat] sum dy*xhat
  int prezeroed;        // t1/t2 were zero-filled by the caller (one memset per backward call)
  // C8 output of dz (dzFmt.c8 = 1): pass 1 collects {max |rstd*gamma|, max |dy|, max |xhat|} (float bits,
  // atomicMax, zero-filled by the caller) in mx[0..2]; pass 2 derives the power-of-two scale S with
  // |dz*S| <= 2^14 from the bound |dz| <= mx0*mx1*(2 + mx2), writes the planes with it and publishes
  // the scale record {1/S, 1/E} for the GEMMs that consume dz.
  PlaneFmt dzFmt;
  unsigned int* mx;
  float* dzRec;
  float* dgamma;        // engine order, accumulated (+=) over images
  float* dbeta;
  __nv_bfloat16* dz_hi; // [rows][Nz]
  __nv_bfloat16* dz_lo;
  float* dbias;         // [Nz] accumulated, or null
};

cudaError_t launch_stats(const float* z, int Nz, int P, int nImg, int groups, float* mean,
                         float* rstd, cudaStream_t s);
cudaError_t launch_stats_finalize(const float* sum, const float* sq, int nImg, int Nz, int groups,
                                  int countPerGroup, float* mean, float* rstd, cudaStream_t s);
cudaError_t launch_apply_fwd(const ApplyArgs& a, cudaStream_t s);
cudaError_t launch_apply_bwd_reduce(const ApplyBwdArgs& a, cudaStream_t s);
cudaError_t launch_apply_bwd(const ApplyBwdArgs& a, cudaStream_t s);

// stem operand builders
cudaError_t launch_prep_g(const float* x, const float* mask, int B, int T, __nv_bfloat16* hi,
                          __nv_bfloat16* lo, cudaStream_t s);
// Discriminator stem (3x3, 1 -> 128 channels, K = 9): direct CUDA-core kernels fused with swish
cudaError_t launch_d_stem_fwd(const float* x, int B, int T, const __nv_bfloat16* wh,
                              const __nv_bfloat16* wl, const float* bias, ActBuf out, cudaStream_t s);
cudaError_t launch_d_stem_bwd(const float* x, int B, int T, const __nv_bfloat16* wh,
                              const __nv_bfloat16* wl, const float* bias, ActBuf dA, float* dW,
                              float* dB, float* q, cudaStream_t s);
// heads: P is [B*Y*X][128] fp32 per-tap partial products
cudaError_t launch_head_g_fwd(const float* P, const float* bias, int B, int Y, int X, float* out,
                              cudaStream_t s);
cudaError_t launch_head_g_bwd(const float* dout, int B, int Y, int X, __nv_bfloat16* dP_hi,
                              __nv_bfloat16* dP_lo, float* dbias, cudaStream_t s);
cudaError_t launch_head_d_fwd(const float* P, const float* bias, int B, int Y, int X, float* out,
                              cudaStream_t s);
cudaError_t launch_head_d_bwd(const float* dout, const float* out, int B, int Y, int X,
                              __nv_bfloat16* dP_hi, __nv_bfloat16* dP_lo, float* dbias,
                              cudaStream_t s);
// stem input gradients (col2im of the stem operand gradient)
cudaError_t launch_col2im_g(const float* dX15, const float* mask, int B, int T, float* dx,
                            cudaStream_t s);
cudaError_t launch_col2im_d(const float* q, int B, int T, float* dx, cudaStream_t s);

// Weight packing.  Reference tensor is [N][C][T] fp32 (OIHW with T = KH*KW); `kind` selects how
// engine coordinates (t', n', c') map onto it (see pack_map in layers.cu).
enum PackKind {
  kPackStd = 0,       // t'=t, c'=c, n' = n + nOffset
  kPackShuffle = 1,   // std with n' = (n%4)*(N/4) + n/4        (PixelShuffle channel grouping)
  kPackStemG = 2,     // ref [N][2][5*15]: t'=kh, c' = kw*2+c
  kPack2dTo1d = 3,    // ref [N][256*20][1]: t' = cRef%20, c' = cRef/20
  kPack1dTo2d = 4,    // ref [256*20][C][1]: n' = (n%20)*256 + n/20
  kPackHead = 5,      // ref [1][C][T]: n' = t, t' = 0
  kPackStemD = 6,     // ref [N][1][9]: c' = t, t' = 0
};
struct PackArgs {
  int kind;
  const float* ref;     // reference-layout weights
  int N, C, T;          // reference dims
  int nOffset;          // engine row offset (fused conv||gates)
  int Np, Cp, Tp;       // engine dims of the fprop layout [Tp][Np][Cp]
  __nv_bfloat16* f_hi;  // fprop layout [Tp][Np][Cp]
  __nv_bfloat16* f_lo;
  __nv_bfloat16* d_hi;  // dgrad layout [Tp][Cd][Np] (Cd = Cp rounded up to 64), or null
  __nv_bfloat16* d_lo;
  int Cd;
};

// Small per-channel vectors (bias / gamma / beta): engine[i'] <-> ref[i] permutations.
enum VecKind { kVecIdent = 0, kVecShuffle = 1, kVecHC20 = 2 };

// Table-driven variants: one launch packs / unpacks every tensor of a model (blockIdx.y = entry).
struct PackEntry {
  int kind, N, C, T, nOffset, Np, Cp, Tp, Cd;
  int refOff;             // float offset in the reference-order flat buffer
  int fHi, fLo, dHi, dLo; // bf16 element offsets in the packed blob (dHi < 0: no data-gradient copy)
  int gW;                 // float offset in the engine-layout gradient blob
  // C8 planes (c8 = 1): the fHi/dHi regions hold fp16, the fLo/dLo regions two e4m3 planes of
  // fElems / dElems bytes; rec = float offset (packed fp32 area) of the conv's record
  // {1/S = 1, 1/E, amax bits, E}; E = largest power of two with amax*E <= 224
  int c8, rec, fElems, dElems;
};
struct PackTable {
  int count;
  int actRec;             // float offset of the static activation record {1, 1/2} (c8 mode), or -1
  float scale;            // unpack only: gradFlat += scale * blob (1/world folds the data-parallel average)
  PackEntry e[32];
};
struct VecEntry {
  int kind, n, refOff, engOff;
};
struct VecTable {
  int count;
  float scale;            // unpack only, see PackTable
  VecEntry e[96];
};
cudaError_t launch_pack_weights_table(const PackTable& t, const float* params, __nv_bfloat16* packed,
                                      float* packedF32, cudaStream_t s);
cudaError_t launch_unpack_wgrads_table(const PackTable& t, const float* gblob, float* gradFlat,
                                       cudaStream_t s);
cudaError_t launch_pack_vecs_table(const VecTable& t, const float* params, float* eng, cudaStream_t s);
cudaError_t launch_unpack_vecs_table(const VecTable& t, const float* eng, float* gradFlat,
                                     cudaStream_t s);

cudaError_t launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1,
                        float b2, float eps, int step, cudaStream_t s);
cudaError_t launch_fill_zero(void* p, size_t bytes, cudaStream_t s);

}  // namespace mcgvc
