// Network-level description of the two models on the hot path and the host-side orchestration of
// their forward / backward passes (one C call = one whole Generator or Discriminator pass on a
// CUDA stream).  Reference: mask_cyclegan_vc/model.py:106-349.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "gemm_types.cuh"
#include "layers.cuh"

namespace mcgvc {

// One tensor-core convolution (possibly the fusion of a conv and its gate conv, which read the
// same input) and where its reference parameters / engine-layout copies / gradients live.
struct ConvDesc {
  std::string name;
  int nParts;                 // 1, or 2 for conv || gates
  long long wOff[2], bOff[2]; // float offsets in the reference-order flat parameter buffer
  int refN, refC, refT;       // reference dims per part ([N][C][T])
  int kind;                   // PackKind
  int biasKind;               // VecKind
  int Np, Cp, Tp, Cd;         // engine dims; Cd = GEMM-N of the data-gradient layout (0: none)
  long long fHi, fLo, dHi, dLo;  // bf16 element offsets in the packed weight blob
  long long biasEng;          // float offset in the packed small-vector area
  long long gW, gB;           // float offsets in the engine-layout gradient blob
  // C8 precision mode: every conv except the stems and heads keeps fp16 + 2 x e4m3 planes in the
  // same storage (fHi/dHi = fp16, fLo/dLo = two e4m3 planes of fElems / dElems bytes)
  int c8Eligible;
  long long fElems, dElems;   // elements per plane of the two layouts
  long long c8Rec;            // float offset (packed fp32 area) of {1/S, 1/E, amax bits, E}
};
struct NormDesc {
  std::string name;
  int nParts;
  long long gOff[2], bOff[2];
  int n;                      // channels per part
  int vecKind;
  int Nstat;                  // nParts * n, or n/20 for the HC20 layer (affPeriod 20)
  int affPeriod;
  long long gammaEng, betaEng;   // float offsets in the packed small-vector area
  long long gGamma, gBeta;       // float offsets in the gradient blob
};

struct ModelDesc {
  std::vector<ConvDesc> convs;
  std::vector<NormDesc> norms;
  long long paramCount;       // floats in the reference-order flat buffer
  long long packedBf16;       // bf16 elements in the packed blob (weights)
  long long packedF32;        // floats in the packed blob (small vectors), placed after the bf16 area
  long long gradFloats;       // floats in the engine-layout gradient blob
  long long actRec;           // float offset (packed fp32 area) of the static activation record {1, 1/2}
  // parameters that never receive a gradient (the Discriminator's downSample4, model.py:316-320 vs
  // :340-349): float range [deadBegin, deadBegin + deadLen) of the reference-order flat buffer.  The
  // "live" gradient layout is the flat layout with this range cut out.
  long long deadBegin, deadLen;
  long long packed_bytes() const { return packedBf16 * 2 + packedF32 * 4; }
};

const ModelDesc& generator_desc();
const ModelDesc& discriminator_desc();

struct RunCfg {
  cudaStream_t stream;
  int backend;   // 0 = tcgen05/TMA kernels, 1 = SIMT checking kernels
  int nPass;     // passes used by the call: 3 = split-bf16, 1 = bf16 (capi picks it per direction)
  cudaStream_t side;  // stream for the weight-gradient GEMMs of a backward pass, or null (same stream)
  cudaEvent_t forkEvent;  // persistent event used to fork to / join from the side stream
  int c8;        // 1 = C8 precision mode: fp16 main pass + two e4m3 correction passes (stems / heads: split-bf16)
  int half16;    // C8H backward: the data-gradient GEMMs of C8 layers run ONE fp16 pass on the 16-bit planes
                 // (and dz keeps only its fp16 plane)
  int wgradHalf16;  // C8H / C8W backward: the weight-gradient GEMMs of C8 layers run ONE fp16 pass on the 16-bit planes
};

// sizes (bytes) of the per-call buffers the caller provides
long long generator_saved_bytes(int B, int T);
long long generator_fwd_ws_bytes(int B, int T);
long long generator_bwd_ws_bytes(int B, int T);
long long discriminator_saved_bytes(int B, int T);
long long discriminator_fwd_ws_bytes(int B, int T);
long long discriminator_bwd_ws_bytes(int B, int T);

int pack_model(const ModelDesc& d, const float* params, void* packed, const RunCfg& rc);
// live = 0: gradFlat has the full reference-order layout; live = 1: the dead range is cut out
int unpack_grads(const ModelDesc& d, const float* gblob, float* gradFlat, const RunCfg& rc, int live = 0,
                 float scale = 1.f);

int generator_forward(const void* packed, const float* x, const float* mask, int B, int T,
                      float* out, void* saved, void* ws, const RunCfg& rc);
int generator_backward(const void* packed, const void* saved, const float* mask, const float* dout,
                       int B, int T, float* dx, float* gblob, int needWgrad, void* ws,
                       const RunCfg& rc);
int discriminator_forward(const void* packed, const float* x, int B, int T, float* out, void* saved,
                          void* ws, const RunCfg& rc);
int discriminator_backward(const void* packed, const void* saved, const float* out,
                           const float* dout, int B, int T, float* dx, float* gblob, int needWgrad,
                           void* ws, const RunCfg& rc);

// debug: named offsets (bytes) of tensors inside the saved blob, for layer-by-layer parity tests
struct SavedEntry {
  std::string name;
  long long offset, bytes;
};
std::vector<SavedEntry> generator_saved_layout(int B, int T);
std::vector<SavedEntry> discriminator_saved_layout(int B, int T);

}  // namespace mcgvc
