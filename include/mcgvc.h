/* mcgvc.h -- C ABI of libmcgvc.so, the B200-native (sm_100a) MaskCycleGAN-VC conv engine.
 *
 * The reference (GANtastic3/MaskCycleGAN-VC) has no FFI: its hot path is the pair of Python
 * nn.Modules in mask_cyclegan_vc/model.py whose arithmetic PyTorch dispatches to cuDNN / oneDNN.
 * This ABI is what a binding for that path attaches to; each entry point names the reference
 * interface it replaces.  All pointers are raw device pointers (plain host integers for sizes); no
 * torch types cross the boundary.  Every function returns 0 on success, non-zero on failure with
 * a message available from mcgvc_last_error().  Nothing is allocated behind the caller's back:
 * the caller provides the packed-weight blob, the saved-activation blob and the workspace, sized
 * by the *_bytes queries below.  All work is enqueued on the given cudaStream_t (passed as void*).
 */
#ifndef MCGVC_H
#define MCGVC_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MCGVC_GENERATOR 0
#define MCGVC_DISCRIMINATOR 1
#define MCGVC_BACKEND_TCGEN05 0   /* tcgen05 + TMA kernels (the product path) */
#define MCGVC_BACKEND_SIMT 1      /* plain CUDA checking kernels (tests / debugging only) */
#define MCGVC_PRECISION_PARITY 3  /* split-bf16 x3: Ah*Wh + Ah*Wl + Al*Wh, fp32 accumulate */
#define MCGVC_PRECISION_MIXED 2   /* forward split-bf16 x3 (output parity), backward single bf16 */
#define MCGVC_PRECISION_FAST 1    /* single bf16 pass */
#define MCGVC_PRECISION_C8 4      /* fp16 main pass + two e4m3 correction passes (2 MMA units per MAC instead of 3):
                                     A*W ~= Ah*Wh + 2^-s (A8h*W8l + A8l*W8h); stems and heads stay split-bf16 x3.
                                     Weights must be packed in the mode they are used in. */
#define MCGVC_PRECISION_C8H 5     /* forward exactly as C8 (Generator output inside the 1e-3 gate); the backward GEMMs of the
                                     C8 layers run ONE fp16 pass over the same operands' 16-bit planes (dynamic power-of-two
                                     scale on dz): gradients are TF32-class (~1e-3 relative), i.e. the accuracy class of the
                                     reference's own CUDA default (cudnn.allow_tf32).  Same packed layout as C8. */
#define MCGVC_PRECISION_C8W 6     /* forward AND data-gradient GEMMs exactly as C8; only the weight-gradient GEMMs of the C8
                                     layers run ONE fp16 pass over the operands' 16-bit planes (the C8H weight-gradient
                                     kernels).  A weight gradient is a leaf of the backward graph -- its rounding error is
                                     not propagated through further layers -- so outputs, input gradients and packed
                                     parameter gradients stay inside the same 1e-3 gate as C8 (tests/test_gpu_network.py,
                                     every C8 referee also runs in this mode) at 5/3 instead of 2 MMA units per MAC of
                                     a training step.  Same packed layout as C8. */

const char* mcgvc_last_error(void);
int mcgvc_set_device(int device);
int mcgvc_set_backend(int backend);
int mcgvc_set_precision(int mode);
int mcgvc_get_precision(void);
/* 1 (default): backward passes use the engine's two priority streams (weight-gradient GEMMs overlap the
 * critical path); 0: everything on the caller's stream (used while timing individual kernels). */
int mcgvc_set_overlap(int on);

/* Model geometry.  Replaces Generator.__init__ / Discriminator.__init__ bookkeeping
 * (model.py:110-211, :287-327): number of floats in the reference-order flat parameter buffer
 * (order of nn.Module.parameters()), and sizes of the engine-side blobs. */
long long mcgvc_param_count(int model);
long long mcgvc_packed_bytes(int model);
long long mcgvc_grad_blob_floats(int model);
long long mcgvc_saved_bytes(int model, int batch, int frames);
long long mcgvc_fwd_workspace_bytes(int model, int batch, int frames);
long long mcgvc_bwd_workspace_bytes(int model, int batch, int frames);
int mcgvc_generator_out_frames(int frames); /* model.py:278-279: 4*ceil(ceil(T/2)/2) */
int mcgvc_discriminator_out_frames(int frames); /* model.py:348: ceil(T/8) */

/* Reference-layout fp32 parameters (OIHW conv weights, biases, InstanceNorm affine) -> engine
 * layout (per-tap [N][C] bf16 hi/lo slices for the forward and data-gradient GEMMs, permuted
 * small vectors).  Call after every optimizer step on the module's parameters. */
int mcgvc_pack_weights(int model, const float* params_flat, void* packed, void* stream);
/* grad_flat (reference order) += engine-layout gradient blob.  Replaces autograd's AccumulateGrad
 * for the module's parameters (train.py:241,298). */
int mcgvc_unpack_grads(int model, const float* grad_blob, float* grad_flat, void* stream);
/* Parameters that never receive a gradient: the Discriminator's downSample4 (constructed at
 * model.py:316-320, never called by forward :340-349; train.py's Adam skips it because its .grad stays
 * None).  [begin, begin + len) is its float range in the reference-order flat buffer (len = 0 for the
 * Generator).  The "live" gradient layout is the flat layout with that range cut out: it is what the
 * data-parallel all-reduce moves (24 811 524 instead of 66 766 852 floats for the four discriminators). */
int mcgvc_dead_param_range(int model, long long* begin, long long* len);
/* grad_live (live layout) += scale * engine-layout gradient blob.  scale = 1/world_size folds the
 * data-parallel average into this pass, so the all-reduce that follows is a plain sum. */
int mcgvc_unpack_grads_live(int model, const float* grad_blob, float* grad_live, float scale, void* stream);

/* Generator.forward(x, mask), model.py:239-280.  x, mask: [B][80][T] fp32; out: [B][80][T'] fp32. */
int mcgvc_generator_forward(const void* packed, const float* x, const float* mask, int batch,
                            int frames, float* out, void* saved, void* workspace, void* stream);
/* Backward of the above (autograd through model.py:239-280).  dout: [B][80][T']; dx: [B][80][T] or
 * NULL when the input needs no gradient; grad_blob: engine-layout fp32 accumulator (+=), may be
 * NULL when need_wgrad == 0. */
int mcgvc_generator_backward(const void* packed, const void* saved, const float* mask,
                             const float* dout, int batch, int frames, float* dx, float* grad_blob,
                             int need_wgrad, void* workspace, void* stream);
/* Discriminator.forward(x), model.py:340-349.  x: [B][80][T]; out: [B][1][10][ceil(T/8)]. */
int mcgvc_discriminator_forward(const void* packed, const float* x, int batch, int frames,
                                float* out, void* saved, void* workspace, void* stream);
int mcgvc_discriminator_backward(const void* packed, const void* saved, const float* out,
                                 const float* dout, int batch, int frames, float* dx,
                                 float* grad_blob, int need_wgrad, void* workspace, void* stream);

/* One Adam update (torch.optim.Adam semantics, no weight decay / amsgrad; replaces the optimizer
 * steps at train.py:242,299 for a contiguous 16-byte-aligned range of the flat parameter buffer).
 * step is the 1-based update count used for bias correction. */
int mcgvc_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                    float lr, float beta1, float beta2, float eps, int step, void* stream);

/* Accounting for bench.py: number of kernels this library has launched so far, and optional
 * per-launch CUDA-event timing of the two tensor-core kernels.  mcgvc_profile_collect synchronises
 * and fills out6 = {conv ms, conv algorithmic FLOPs, conv launches, wgrad ms, wgrad FLOPs, wgrad
 * launches} accumulated since the previous collect. */
long long mcgvc_launch_count(void);
/* CUDA-graph replay of repeated identical forward/backward calls (opt-in: MCGVC_GRAPHS=1 or
 * mcgvc_set_graphs(1); helps batch-1 latency, neutral at batch 64).  mcgvc_graph_stats reports graphs captured / replays so far. */
int mcgvc_set_graphs(int on);
int mcgvc_graph_stats(long long* captures, long long* replays);
int mcgvc_profile_enable(int on);
int mcgvc_profile_collect(double* out6);
/* The same per kernel family (conv_c8_kernel<256>, conv_tc2_kernel<256>, the fused trunk kernels, ...): out
 * receives {ms, algorithmic FLOPs, launches} per kind, the return value is the number of kinds written;
 * mcgvc_profile_kind_name names a kind.  bench.py reports the family with the largest time share as the
 * step's dominant kernel. */
int mcgvc_profile_collect_kinds(double* out, int max_kinds);
const char* mcgvc_profile_kind_name(int kind);

/* Device-side data feed: replaces the crop + frame-in-fill mask work of
 * dataset/vc_dataset.py:44-56 and the host->device copies of mask_cyclegan_vc/train.py:187-190.
 * pool: every utterance as a (80, T_u) row-major float array, back to back, on the device; utt_off[u]
 * = float offset of utterance u, utt_frames[u] = T_u.  sel = int[4][batch] on the device: utterance,
 * crop start (0 <= start <= T_u - n_frames), mask start, mask size (mask_start + mask_size <= n_frames).
 * Writes x[batch][80][n_frames] (the crop) and mask[batch][80][n_frames] (0 inside the mask span, else
 * 1).  A selection outside those bounds never reads outside the pool (its sample comes out as zeros);
 * bad_count (device int, may be null) receives how many there were. */
int mcgvc_crop_mask(const float* pool, const long long* utt_off, const int* utt_frames, int n_utts,
                    const int* sel, int batch, int n_frames, float* x, float* mask, int* bad_count,
                    void* stream);

/* Loss tail: one weighted term of the generator / discriminator losses of
 * mask_cyclegan_vc/train.py:219-237 and :276-294, accumulated into the device scalar *total (which the
 * caller zero-fills once per phase).  L1: total += weight * mean|a - b| (cycle / identity terms, :219-224);
 * LSGAN: total += weight * mean (target - a)^2 (adversarial terms, :227-232 and :276-288; b unused).
 * mcgvc_loss_term_grad writes d total / d a for one term (grad_total: device scalar with the upstream
 * gradient, or null for 1). */
#define MCGVC_LOSS_L1 0
#define MCGVC_LOSS_LSGAN 1
int mcgvc_loss_term(const float* a, const float* b, long long n, int kind, float target, float weight,
                    float* total, void* stream);
int mcgvc_loss_term_grad(const float* a, const float* b, long long n, int kind, float target, float weight,
                         const float* grad_total, float* grad_a, void* stream);

/* Introspection for layer-by-layer parity tests: the index-th named tensor inside the saved blob.
 * Returns 0 and fills name/offset/bytes, or 1 when index is past the end. */
int mcgvc_saved_layout(int model, int batch, int frames, int index, char* name, int name_cap,
                       long long* offset, long long* bytes);

/* Kernel-level entry points (tests): one implicit-GEMM convolution / one weight-gradient GEMM on
 * caller-provided bf16 hi/lo operands, through either backend. */
int mcgvc_debug_conv(const void* a_hi, const void* a_lo, int aC, int aX, int aY, int aP, int aB,
                     const void* w_hi, const void* w_lo, int wK, int wN, int wT, int oX, int oY,
                     int oB, int nTaps, const int8_t* taps4, float* out, long long sB,
                     long long sY, long long sX, int nSplit, long long sNhi, const float* bias,
                     const float* addsrc, int nPass, int backend, int blockN, void* stream);
/* Host-side split-K planner for a convolution geometry (output grid oB x oY x oX, C input channels, N
 * output columns, nTaps filter taps) on the split-bf16 kernels; needs no GPU (assumes 148 SMs then). */
int mcgvc_debug_plan_ksplit(int oB, int oY, int oX, int C, int N, int nSplit, int nTaps, double minGain);
/* Host-side tail-split planner of the CTA-pair kernels (blockN = tile width 128 / 256): number of K-slices the
 * tiles of the partial last wave are cut into (0 = none); *tail_tiles receives how many tiles that is. */
int mcgvc_debug_plan_tail(int oB, int oY, int oX, int C, int N, int nSplit, int nTaps, int blockN, int* tail_tiles);
/* split-K factor used by the following mcgvc_debug_conv calls on the tensor-core backends: every tile's
 * k-blocks run as `k` work items that are added into `out` (which the caller zero-fills); 1 = off. */
int mcgvc_debug_set_conv_ksplit(int k);
/* Kernel-level entry of the "16-bit main pass + two e4m3 correction passes" convolution (conv_c8.cu; the
 * network path reaches the same kernel under MCGVC_PRECISION_C8): out = out_scale * (sum a16*w16 + corr_scale * sum (a8h*w8l + a8l*w8h))
 * + bias + addsrc.  backend 1 = SIMT checker, 2 = tcgen05 CTA-pair kernel (blockN 128 / 256). */
int mcgvc_debug_conv_c8(const void* a16, const void* a8h, const void* a8l, int aC, int aX, int aY, int aP,
                        int aB, const void* w16, const void* w8h, const void* w8l, int wK, int wN, int wT,
                        int oX, int oY, int oB, int nTaps, const int8_t* taps4, float* out,
                        const float* bias, const float* addsrc, int main_bf16, float out_scale,
                        float corr_scale, int backend, int blockN, void* stream);
int mcgvc_debug_wgrad(const void* z_hi, const void* z_lo, int zC, int zX, int zY, int zB,
                      const void* x_hi, const void* x_lo, int xC, int xX, int xY, int xP, int xB,
                      int pX, int pY, int pB, int nTaps, const int8_t* taps4,
                      const int8_t* ztaps4, float* dw, int cTile, int splitK, int nPass,
                      int backend, void* stream);

/* Weight-gradient GEMM in the same scheme (wgrad_c8.cu): dw += out_scale * (sum z16*x16 + corr_scale *
 * sum (z8h*x8l + z8l*x8h)) over positions.  backend 1 = SIMT checker, 2 = tcgen05 CTA-pair kernel
 * (zC multiple of 256, cTile 128 or 256). */
int mcgvc_debug_wgrad_c8(const void* z16, const void* z8h, const void* z8l, int zC, int zX, int zY, int zB,
                         const void* x16, const void* x8h, const void* x8l, int xC, int xX, int xY, int xP,
                         int xB, int pX, int pY, int pB, int nTaps, const int8_t* taps4, const int8_t* ztaps4,
                         float* dw, int cTile, int splitK, int main_bf16, float out_scale, float corr_scale,
                         int backend, void* stream);

#ifdef __cplusplus
}
#endif
#endif
