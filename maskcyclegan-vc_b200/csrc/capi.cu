// extern "C" surface of libmcgvc.so (see include/mcgvc.h).
#include "../../include/mcgvc.h"
#include "network.cuh"

#include <cstdlib>
#include <cstring>

using namespace mcgvc;

static int g_backend = MCGVC_BACKEND_TCGEN05;
static int g_precision = MCGVC_PRECISION_PARITY;

// parity: every GEMM split-bf16 x3; fast: every GEMM single bf16; mixed: forward x3, backward x1
// Backward passes run on two engine-owned streams per device: a HIGH-priority main stream for the
// critical path (layer kernels + data-gradient convs) and a low-priority side stream for the
// weight-gradient GEMMs, which only feed the gradient blob and fill whatever SM time the main chain
// leaves.  The caller's stream waits for both before the call returns control of the buffers
// (stream-ordered), so the C ABI contract "work is enqueued on the caller's stream" still holds.
// MCGVC_OVERLAP=0 runs everything on the caller's stream.
struct DevStreams {
  cudaStream_t main = nullptr, side = nullptr;
  cudaEvent_t in = nullptr, out = nullptr;
};
static int g_overlap = -1;   // -1: take MCGVC_OVERLAP (default on)
static DevStreams* dev_streams() {
  static DevStreams ds[64];
  if (g_overlap < 0) { const char* e = getenv("MCGVC_OVERLAP"); g_overlap = e ? atoi(e) : 1; }
  if (!g_overlap) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  DevStreams& d = ds[dev];
  if (!d.main) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi is the numerically smallest = greatest priority
    if (cudaStreamCreateWithPriority(&d.main, cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&d.side, cudaStreamNonBlocking, lo) != cudaSuccess) return nullptr;
    cudaEventCreateWithFlags(&d.in, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&d.out, cudaEventDisableTiming);
  }
  return &d;
}
static RunCfg cfg(void* stream, bool backward = false) {
  int np = g_precision == MCGVC_PRECISION_PARITY ? 3 : (g_precision == MCGVC_PRECISION_FAST ? 1 : (backward ? 1 : 3));
  return RunCfg{(cudaStream_t)stream, g_backend, np, nullptr};
}
// run a backward body on the engine streams, bracketed by event hand-offs with the caller's stream
template <class F>
static int run_backward(void* stream, F body) {
  RunCfg rc = cfg(stream, true);
  DevStreams* d = dev_streams();
  if (!d) return body(rc);
  cudaStream_t caller = (cudaStream_t)stream;
  cudaEventRecord(d->in, caller);
  cudaStreamWaitEvent(d->main, d->in, 0);
  rc.stream = d->main;
  rc.side = d->side;
  const int rv = body(rc);               // joins the side stream into d->main before returning
  cudaEventRecord(d->out, d->main);
  cudaStreamWaitEvent(caller, d->out, 0);
  return rv;
}
static const ModelDesc* desc(int model) {
  if (model == MCGVC_GENERATOR) return &generator_desc();
  if (model == MCGVC_DISCRIMINATOR) return &discriminator_desc();
  set_error("unknown model id %d", model);
  return nullptr;
}
static bool shape_ok(int B, int T) {
  if (B < 1 || T < 1) { set_error("batch and frames must be >= 1 (got %d, %d)", B, T); return false; }
  return true;
}

extern "C" {

int mcgvc_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e)); return 1; }
  return 0;
}
int mcgvc_set_backend(int backend) {
  if (backend != 0 && backend != 1) { set_error("backend must be 0 or 1"); return 1; }
  g_backend = backend;
  return 0;
}
int mcgvc_set_precision(int mode) {
  if (mode != MCGVC_PRECISION_PARITY && mode != MCGVC_PRECISION_FAST && mode != MCGVC_PRECISION_MIXED) {
    set_error("precision must be MCGVC_PRECISION_PARITY (3), _MIXED (2) or _FAST (1)");
    return 1;
  }
  g_precision = mode;
  return 0;
}
int mcgvc_get_precision(void) { return g_precision; }
int mcgvc_set_overlap(int on) { g_overlap = on ? 1 : 0; return 0; }

long long mcgvc_param_count(int model) { const ModelDesc* d = desc(model); return d ? d->paramCount : -1; }
long long mcgvc_packed_bytes(int model) { const ModelDesc* d = desc(model); return d ? d->packed_bytes() : -1; }
long long mcgvc_grad_blob_floats(int model) { const ModelDesc* d = desc(model); return d ? d->gradFloats : -1; }
long long mcgvc_saved_bytes(int model, int B, int T) {
  if (!shape_ok(B, T)) return -1;
  return model == MCGVC_GENERATOR ? generator_saved_bytes(B, T) : discriminator_saved_bytes(B, T);
}
long long mcgvc_fwd_workspace_bytes(int model, int B, int T) {
  if (!shape_ok(B, T)) return -1;
  return model == MCGVC_GENERATOR ? generator_fwd_ws_bytes(B, T) : discriminator_fwd_ws_bytes(B, T);
}
long long mcgvc_bwd_workspace_bytes(int model, int B, int T) {
  if (!shape_ok(B, T)) return -1;
  return model == MCGVC_GENERATOR ? generator_bwd_ws_bytes(B, T) : discriminator_bwd_ws_bytes(B, T);
}
int mcgvc_generator_out_frames(int T) { return 4 * ((((T + 1) / 2) + 1) / 2); }
int mcgvc_discriminator_out_frames(int T) { return ((((T + 1) / 2 + 1) / 2) + 1) / 2; }

int mcgvc_pack_weights(int model, const float* params, void* packed, void* stream) {
  const ModelDesc* d = desc(model);
  if (!d) return 1;
  if (!params || !packed) { set_error("pack_weights: null pointer"); return 1; }
  return pack_model(*d, params, packed, cfg(stream));
}
int mcgvc_unpack_grads(int model, const float* gblob, float* grad_flat, void* stream) {
  const ModelDesc* d = desc(model);
  if (!d) return 1;
  if (!gblob || !grad_flat) { set_error("unpack_grads: null pointer"); return 1; }
  return unpack_grads(*d, gblob, grad_flat, cfg(stream));
}

int mcgvc_generator_forward(const void* packed, const float* x, const float* mask, int B, int T,
                            float* out, void* saved, void* ws, void* stream) {
  if (!shape_ok(B, T)) return 1;
  if (!packed || !x || !mask || !out || !saved || !ws) { set_error("generator_forward: null pointer"); return 1; }
  return generator_forward(packed, x, mask, B, T, out, saved, ws, cfg(stream));
}
int mcgvc_generator_backward(const void* packed, const void* saved, const float* mask,
                             const float* dout, int B, int T, float* dx, float* gblob,
                             int need_wgrad, void* ws, void* stream) {
  if (!shape_ok(B, T)) return 1;
  if (!packed || !saved || !mask || !dout || !ws || (need_wgrad && !gblob)) { set_error("generator_backward: null pointer"); return 1; }
  return run_backward(stream, [&](const RunCfg& rc) { return generator_backward(packed, saved, mask, dout, B, T, dx, gblob, need_wgrad, ws, rc); });
}
int mcgvc_discriminator_forward(const void* packed, const float* x, int B, int T, float* out,
                                void* saved, void* ws, void* stream) {
  if (!shape_ok(B, T)) return 1;
  if (!packed || !x || !out || !saved || !ws) { set_error("discriminator_forward: null pointer"); return 1; }
  return discriminator_forward(packed, x, B, T, out, saved, ws, cfg(stream));
}
int mcgvc_discriminator_backward(const void* packed, const void* saved, const float* out,
                                 const float* dout, int B, int T, float* dx, float* gblob,
                                 int need_wgrad, void* ws, void* stream) {
  if (!shape_ok(B, T)) return 1;
  if (!packed || !saved || !out || !dout || !ws || (need_wgrad && !gblob)) { set_error("discriminator_backward: null pointer"); return 1; }
  return run_backward(stream, [&](const RunCfg& rc) { return discriminator_backward(packed, saved, out, dout, B, T, dx, gblob, need_wgrad, ws, rc); });
}

long long mcgvc_launch_count(void) { return launch_count(); }
int mcgvc_profile_enable(int on) { profile_enable(on != 0); return 0; }
int mcgvc_profile_collect(double* out6) {
  KernelProfile c, w;
  profile_collect(&c, &w);
  out6[0] = c.ms; out6[1] = c.flops; out6[2] = (double)c.launches;
  out6[3] = w.ms; out6[4] = w.flops; out6[5] = (double)w.launches;
  return 0;
}

int mcgvc_saved_layout(int model, int B, int T, int index, char* name, int name_cap,
                       long long* offset, long long* bytes) {
  if (!shape_ok(B, T)) return 1;
  std::vector<SavedEntry> v = model == MCGVC_GENERATOR ? generator_saved_layout(B, T)
                                                       : discriminator_saved_layout(B, T);
  if (index < 0 || index >= (int)v.size()) return 1;
  if (name && name_cap > 0) {
    strncpy(name, v[index].name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (offset) *offset = v[index].offset;
  if (bytes) *bytes = v[index].bytes;
  return 0;
}

}  // extern "C"
