// Device-side training data feed: crop + frame-in-fill mask generation from a
// device-resident mel pool, replacing the per-item host work of the reference's
// dataset/vc_dataset.py:19-77 (which re-shuffles and re-crops the WHOLE dataset for every item it
// returns) and the pageable host->device copies of mask_cyclegan_vc/train.py:187-190.
//
// The pool holds every utterance as the reference stores it -- a (80, T_u) row-major float array --
// back to back; utt_off[u] is the float offset of utterance u and utt_frames[u] its T_u.  One
// selection per output sample, sel[4][B] = {utterance, crop start, mask start, mask size}, is drawn
// on the host with the reference's distributions (datafeed.py) and costs 16 bytes per sample of
// host->device traffic instead of 2 * 80 * n_frames * 4.
//   x[b, m, f]    = pool[utt_off[u] + m * T_u + start + f]              vc_dataset.py:46-48,55
//   mask[b, m, f] = 0 if mask_start <= f < mask_start + mask_size else 1   vc_dataset.py:51-54
// Pure byte movement: HBM-bound, 4 B read + 8 B written per output element, coalesced along f.
#include "../../include/mcgvc.h"
#include "gemm_types.cuh"

using namespace mcgvc;

namespace {
constexpr int kMel = 80;

__global__ void crop_mask_kernel(const float* __restrict__ pool, const long long* __restrict__ utt_off,
                                 const int* __restrict__ utt_frames, int n_utts, const int* __restrict__ sel,
                                 int B, int n_frames, float* __restrict__ x, float* __restrict__ mask) {
  const long long total = (long long)B * kMel * n_frames;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(idx % n_frames);
    const int m = (int)((idx / n_frames) % kMel);
    const int b = (int)(idx / ((long long)n_frames * kMel));
    const int u = sel[b], start = sel[B + b], ms = sel[2 * B + b], mlen = sel[3 * B + b];
    // an out-of-range selection never reads outside the pool: the sample comes out as zeros
    const bool ok = u >= 0 && u < n_utts && start >= 0 && start + n_frames <= utt_frames[u >= 0 && u < n_utts ? u : 0];
    const int T = ok ? utt_frames[u] : 0;
    x[idx] = ok ? pool[utt_off[u] + (long long)m * T + start + f] : 0.f;
    mask[idx] = (f >= ms && f < ms + mlen) ? 0.f : 1.f;
  }
}

// optional report of how many selections were out of range
__global__ void check_sel_kernel(const int* __restrict__ utt_frames, int n_utts, const int* __restrict__ sel,
                                 int B, int n_frames, int* __restrict__ bad) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int u = sel[b], start = sel[B + b], ms = sel[2 * B + b], mlen = sel[3 * B + b];
  bool ok = u >= 0 && u < n_utts;
  if (ok) ok = start >= 0 && start + n_frames <= utt_frames[u];
  ok = ok && ms >= 0 && mlen >= 0 && ms + mlen <= n_frames;
  if (!ok) atomicAdd(bad, 1);
}
}  // namespace

extern "C" int mcgvc_crop_mask(const float* pool, const long long* utt_off, const int* utt_frames, int n_utts,
                               const int* sel, int batch, int n_frames, float* x, float* mask,
                               int* bad_count, void* stream) {
  if (!pool || !utt_off || !utt_frames || !sel || !x || !mask) { set_error("crop_mask: null pointer"); return 1; }
  if (batch < 1 || n_frames < 1 || n_utts < 1) { set_error("crop_mask: batch, n_frames and n_utts must be >= 1"); return 1; }
  cudaStream_t s = (cudaStream_t)stream;
  if (bad_count) {
    cudaMemsetAsync(bad_count, 0, sizeof(int), s);
    check_sel_kernel<<<(batch + 127) / 128, 128, 0, s>>>(utt_frames, n_utts, sel, batch, n_frames, bad_count);
    if (launched() != cudaSuccess) { set_error("crop_mask: check launch failed"); return 1; }
  }
  const long long total = (long long)batch * kMel * n_frames;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  crop_mask_kernel<<<(unsigned)blocks, 256, 0, s>>>(pool, utt_off, utt_frames, n_utts, sel, batch, n_frames, x, mask);
  cudaError_t e = launched();
  if (e != cudaSuccess) { set_error("crop_mask: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}
