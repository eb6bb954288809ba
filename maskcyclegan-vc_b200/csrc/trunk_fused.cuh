// Fused forward of the Generator's six gated 1-D residual blocks (reference
// mask_cyclegan_vc/model.py:40-76 ResidualLayer, :258-263): ONE launch instead of 36.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mcgvc {

constexpr int kTrunkBlocks = 6;

// Everything the kernel needs besides the tensor maps.  All activation tensors are rows (b, x) of the
// [B][W2] position grid, channels innermost; block i reads R[i] and writes z4[i], st4[i], H[i], z5[i],
// st5[i], R[i+1] -- exactly the tensors the layer-by-layer path saves for the backward pass.
struct TrunkFwdArgs {
  int B, W2;
  int BX, BB;                  // tile = BB samples x BX positions = 128 rows (BX = pow2 >= W2)
  int nPass;                   // 3 = split-bf16, 1 = bf16
  float* Rf[kTrunkBlocks + 1];                 // [L][256] fp32
  __nv_bfloat16* Rhi[kTrunkBlocks + 1];        // [L][256]
  __nv_bfloat16* Rlo[kTrunkBlocks + 1];
  __nv_bfloat16* Hhi[kTrunkBlocks];            // [L][512]
  __nv_bfloat16* Hlo[kTrunkBlocks];
  float* z4[kTrunkBlocks];                     // [L][1024] raw conv || gate output
  float* z5[kTrunkBlocks];                     // [L][256]
  float* mean4[kTrunkBlocks];                  // [B][1024]
  float* rstd4[kTrunkBlocks];
  float* mean5[kTrunkBlocks];                  // [B][256]
  float* rstd5[kTrunkBlocks];
  const float* biasA[kTrunkBlocks];            // [1024] engine order (conv || gate)
  const float* gammaA[kTrunkBlocks];
  const float* betaA[kTrunkBlocks];
  const float* biasB[kTrunkBlocks];            // [256]
  const float* gammaB[kTrunkBlocks];
  const float* betaB[kTrunkBlocks];
};

// Operand planes for the tensor maps: plane base of block 0 and the (uniform) byte distance between
// consecutive blocks' planes.
struct TrunkFwdMaps {
  const void* Rhi; const void* Rlo; long long RStrideBytes;       // R[i] planes, [B][W2][256]
  const void* Hhi; const void* Hlo; long long HStrideBytes;       // H[i] planes, [B][W2][512]
  const void* Wah; const void* Wal; long long WaStrideBytes;      // conv a weights [3][1024][256]
  const void* Wbh; const void* Wbl; long long WbStrideBytes;      // conv b weights [3][256][512]
};

// true when the fused kernel covers this shape (else the caller runs the layer-by-layer path)
bool trunk_fwd_supported(int B, int W2);
cudaError_t launch_trunk_fwd(const TrunkFwdArgs& a, const TrunkFwdMaps& m, cudaStream_t stream);

// Fused data-gradient chain of the same six blocks (autograd through model.py:71-76, six times): from
// dR[6] (gradient w.r.t. the trunk output) to dR[0], writing on the way the dz operands (bf16 hi/lo) of
// the twelve weight-gradient GEMMs and accumulating the InstanceNorm affine gradients.
struct TrunkBwdArgs {
  int B, W2;
  int BX, BB;
  int nPass;
  const float* dR6;                            // [L][256] gradient w.r.t. R[6]
  float* dR0;                                  // [L][256] gradient w.r.t. R[0]
  const float* z4[kTrunkBlocks];               // saved by the forward pass
  const float* z5[kTrunkBlocks];
  const float* mean4[kTrunkBlocks];
  const float* rstd4[kTrunkBlocks];
  const float* mean5[kTrunkBlocks];
  const float* rstd5[kTrunkBlocks];
  const float* gammaA[kTrunkBlocks];
  const float* betaA[kTrunkBlocks];
  const float* gammaB[kTrunkBlocks];
  float* dgammaA[kTrunkBlocks];                // [1024] accumulated (+=) over samples
  float* dbetaA[kTrunkBlocks];
  float* dgammaB[kTrunkBlocks];                // [256]
  float* dbetaB[kTrunkBlocks];
  __nv_bfloat16* dz5hi[kTrunkBlocks];          // [L][256]  d(loss)/d(z5), operand of wgrad b and dgrad b
  __nv_bfloat16* dz5lo[kTrunkBlocks];
  __nv_bfloat16* dz4hi[kTrunkBlocks];          // [L][1024]
  __nv_bfloat16* dz4lo[kTrunkBlocks];
};
struct TrunkBwdMaps {
  const void* Z5hi; const void* Z5lo; long long Z5StrideBytes;    // dz5[i] planes [B][W2][256]
  const void* Z4hi; const void* Z4lo; long long Z4StrideBytes;    // dz4[i] planes [B][W2][1024]
  const void* Wbh; const void* Wbl; long long WbStrideBytes;      // conv b data-gradient weights [3][512][256]
  const void* Wah; const void* Wal; long long WaStrideBytes;      // conv a data-gradient weights [3][256][1024]
};
bool trunk_bwd_supported(int B, int W2);
cudaError_t launch_trunk_bwd(const TrunkBwdArgs& a, const TrunkBwdMaps& m, cudaStream_t stream);

}  // namespace mcgvc
