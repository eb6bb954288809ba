// Loss tail of the train step: the L1 cycle / identity terms and the LSGAN
// adversarial terms of mask_cyclegan_vc/train.py:219-237 and :276-294.  The reference evaluates each
// term as 3-4 elementwise aten kernels plus a reduction (about 40 tiny launches per step for the ten
// terms); here a term is ONE reduction kernel forward and ONE elementwise kernel backward, and all
// terms of a phase accumulate, already weighted, into one device scalar.
//   kind 0 (L1):    term = weight * mean |a - b|            train.py:219-224
//   kind 1 (LSGAN): term = weight * mean (target - a)^2     train.py:227-232, :276-288
// HBM-bound: 4-8 B read per element forward, 4-8 B read + 4 B written backward.
#include "../../include/mcgvc.h"
#include "gemm_types.cuh"

using namespace mcgvc;

namespace {
__global__ void loss_term_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, int kind,
                                 float target, float scale, float* __restrict__ out) {
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (kind == 0) {
      acc += fabsf(a[i] - b[i]);
    } else {
      const float d = target - a[i];
      acc = fmaf(d, d, acc);
    }
  }
  __shared__ float red[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    atomicAdd(out, t * scale);
  }
}
// d(term)/da (and, for L1, nothing for b: the reference's b is data); gout = upstream gradient of the total
__global__ void loss_term_grad_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, int kind,
                                      float target, float scale, const float* __restrict__ gout, float* __restrict__ da) {
  const float g = scale * (gout ? *gout : 1.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (kind == 0) {
      const float d = a[i] - b[i];
      da[i] = d > 0.f ? g : (d < 0.f ? -g : 0.f);      // torch: sign(0) = 0
    } else {
      da[i] = -2.f * g * (target - a[i]);
    }
  }
}
int grid_of(long long n) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  return (int)(blocks < 1 ? 1 : blocks);
}
}  // namespace

extern "C" {

int mcgvc_loss_term(const float* a, const float* b, long long n, int kind, float target, float weight,
                    float* total, void* stream) {
  if (!a || !total || n < 1 || (kind != MCGVC_LOSS_L1 && kind != MCGVC_LOSS_LSGAN) || (kind == MCGVC_LOSS_L1 && !b)) {
    set_error("loss_term: bad arguments (kind %d, n %lld)", kind, n);
    return 1;
  }
  loss_term_kernel<<<grid_of(n), 256, 0, (cudaStream_t)stream>>>(a, b, n, kind, target, weight / (float)n, total);
  cudaError_t e = launched();
  if (e != cudaSuccess) { set_error("loss_term: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

int mcgvc_loss_term_grad(const float* a, const float* b, long long n, int kind, float target, float weight,
                         const float* grad_total, float* grad_a, void* stream) {
  if (!a || !grad_a || n < 1 || (kind != MCGVC_LOSS_L1 && kind != MCGVC_LOSS_LSGAN) || (kind == MCGVC_LOSS_L1 && !b)) {
    set_error("loss_term_grad: bad arguments (kind %d, n %lld)", kind, n);
    return 1;
  }
  loss_term_grad_kernel<<<grid_of(n), 256, 0, (cudaStream_t)stream>>>(a, b, n, kind, target, weight / (float)n, grad_total, grad_a);
  cudaError_t e = launched();
  if (e != cudaSuccess) { set_error("loss_term_grad: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // extern "C"
