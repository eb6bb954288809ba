// TMA tensor-map builders shared by the tensor-core kernels (defined in conv_igemm.cu).
#pragma once
#include "gemm_types.cuh"

namespace mcgvc {

// bf16 / fp16 plane of an activation [B][P][Y][X][C]: box (64 channels, BX, BY, 1, BB), 128B swizzle
bool make_act_tmap(CUtensorMap* m, const void* base, const ActOperand& a, int BX, int BY, int BB);
// bf16 / fp16 weights [T][N][K]: box (64, boxN, 1), 128B swizzle
bool make_wgt_tmap(CUtensorMap* m, const void* base, const WgtOperand& w, int boxN);
// e4m3 plane of an activation: box (boxC bytes, BX, BY, 1, BB); 128-byte boxes use the 128B swizzle,
// 64-byte boxes the 64B swizzle
bool make_plane8_tmap(CUtensorMap* m, const void* base, const ActOperand& a, int boxC, int BX, int BY, int BB);
// e4m3 plane of the weights [T][N][K]: box (64 B, boxN, 1), 64B swizzle
bool make_wgt8_tmap(CUtensorMap* m, const void* base, const WgtOperand& w, int boxN);

// generic 16-bit tiled map (rank <= 5), 128B swizzle, zero OOB fill: dims / box innermost first,
// stridesBytes[i] = byte stride of dimension i+1
bool make_tmap16(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims,
                 const unsigned long long* stridesBytes, const unsigned int* box);

}  // namespace mcgvc
