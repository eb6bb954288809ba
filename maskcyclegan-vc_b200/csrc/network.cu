// Host-side orchestration: model descriptions, weight packing, and the forward / backward layer
// schedules of the Generator and Discriminator (reference mask_cyclegan_vc/model.py:239-280 and
// :340-349; backward = what autograd derives for them, train.py:241,298).
#include "network.cuh"
#include "trunk_fused.cuh"

#include <cstdio>
#include <cstdlib>
#include <map>

namespace mcgvc {

// ================================================================================================
// model descriptions
namespace {

struct ParamTable {
  std::map<std::string, long long> off;
  long long total = 0;
  void add(const std::string& name, long long numel) {
    off[name] = total;
    total += numel;
  }
  void conv(const std::string& name, int n, int c, int t) {
    add(name + ".weight", (long long)n * c * t);
    add(name + ".bias", n);
  }
  void norm(const std::string& name, int n) {
    add(name + ".weight", n);
    add(name + ".bias", n);
  }
  long long at(const std::string& name) const { return off.at(name); }
};

long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

struct DescBuilder {
  ModelDesc d;
  const ParamTable& pt;
  explicit DescBuilder(const ParamTable& p) : pt(p) {
    d.packedBf16 = d.gradFloats = 0;
    d.actRec = 0; d.packedF32 = 64;   // first slot of the fp32 area: static activation record
    d.paramCount = p.total;
    d.deadBegin = d.deadLen = 0;
  }
  void conv(const std::string& name, std::vector<std::string> parts, int refN, int refC, int refT,
            int kind, int biasKind, int Np, int Cp, int Tp, int Cd) {
    ConvDesc c{};
    c.name = name;
    c.nParts = (int)parts.size();
    for (int i = 0; i < c.nParts; ++i) {
      c.wOff[i] = pt.at(parts[i] + ".weight");
      c.bOff[i] = pt.at(parts[i] + ".bias");
    }
    c.refN = refN; c.refC = refC; c.refT = refT;
    c.kind = kind; c.biasKind = biasKind;
    c.Np = Np; c.Cp = Cp; c.Tp = Tp; c.Cd = Cd;
    const long long fsz = align_up((long long)Tp * Np * Cp, 128);
    const long long dsz = align_up((long long)Tp * Cd * Np, 128);
    c.fHi = d.packedBf16; d.packedBf16 += fsz;
    c.fLo = d.packedBf16; d.packedBf16 += fsz;
    if (Cd) {
      c.dHi = d.packedBf16; d.packedBf16 += dsz;
      c.dLo = d.packedBf16; d.packedBf16 += dsz;
    } else {
      c.dHi = c.dLo = -1;
    }
    c.fElems = fsz; c.dElems = dsz;
    // C8 planes only where the MMA time dominates: the big 2-D layers.  Stems / heads (padded K or N)
    // and the 1-D trunk with its two flatten layers (latency-bound small GEMMs) stay split-bf16.
    c.c8Eligible = (kind == kPackShuffle || (kind == kPackStd && name.compare(0, 3, "res") != 0)) ? 1 : 0;
    c.c8Rec = d.packedF32; d.packedF32 += 64;
    c.biasEng = d.packedF32; d.packedF32 += align_up(Np, 64);
    c.gW = d.gradFloats; d.gradFloats += align_up((long long)Tp * Np * Cp, 64);
    c.gB = d.gradFloats; d.gradFloats += align_up(Np, 64);
    d.convs.push_back(c);
  }
  void norm(const std::string& name, std::vector<std::string> parts, int n, int vecKind) {
    NormDesc m{};
    m.name = name;
    m.nParts = (int)parts.size();
    for (int i = 0; i < m.nParts; ++i) {
      m.gOff[i] = pt.at(parts[i] + ".weight");
      m.bOff[i] = pt.at(parts[i] + ".bias");
    }
    m.n = n;
    m.vecKind = vecKind;
    m.affPeriod = vecKind == kVecHC20 ? 20 : 1;
    m.Nstat = vecKind == kVecHC20 ? n / 20 : n * m.nParts;
    const long long tot = align_up((long long)n * m.nParts, 64);
    m.gammaEng = d.packedF32; d.packedF32 += tot;
    m.betaEng = d.packedF32; d.packedF32 += tot;
    m.gGamma = d.gradFloats; d.gradFloats += tot;
    m.gBeta = d.gradFloats; d.gradFloats += tot;
    d.norms.push_back(m);
  }
};

// conv indices
enum { G_STEM = 0, G_DS1, G_DS2, G_2DTO1D, G_RES0 /* 2 per block: a, b */, G_1DTO2D = G_RES0 + 12,
       G_UP1, G_UP2, G_HEAD, G_NCONV };
// norm indices
enum { GN_DS1 = 0, GN_DS2, GN_2DTO1D, GN_RES0 /* 2 per block */, GN_1DTO2D = GN_RES0 + 12, GN_UP1,
       GN_UP2, GN_NNORM };
enum { D_STEM = 0, D_DS1, D_DS2, D_DS3, D_HEAD, D_NCONV };
enum { DN_DS1 = 0, DN_DS2, DN_DS3, DN_NNORM };

ModelDesc build_generator() {
  // reference parameter order = Generator.parameters() (model.py:110-211; `convLayer` is the
  // upSample2 Sequential registered under the attribute name written at :227, ahead of upSample1)
  ParamTable pt;
  pt.conv("conv1", 128, 2, 75);
  pt.conv("conv1_gates", 128, 2, 75);
  for (int i = 1; i <= 2; ++i) {
    const std::string p = "downSample" + std::to_string(i);
    const int cin = i == 1 ? 128 : 256;
    pt.conv(p + ".convLayer.0", 256, cin, 25);
    pt.norm(p + ".convLayer.1", 256);
    pt.conv(p + ".convLayer_gates.0", 256, cin, 25);
    pt.norm(p + ".convLayer_gates.1", 256);
  }
  pt.conv("conv2dto1dLayer", 256, 5120, 1);
  pt.norm("conv2dto1dLayer_tfan", 256);
  for (int i = 1; i <= 6; ++i) {
    const std::string p = "residualLayer" + std::to_string(i);
    pt.conv(p + ".conv1d_layer.0", 512, 256, 3);
    pt.norm(p + ".conv1d_layer.1", 512);
    pt.conv(p + ".conv_layer_gates.0", 512, 256, 3);
    pt.norm(p + ".conv_layer_gates.1", 512);
    pt.conv(p + ".conv1d_out_layer.0", 256, 512, 3);
    pt.norm(p + ".conv1d_out_layer.1", 256);
  }
  pt.conv("conv1dto2dLayer", 5120, 256, 1);
  pt.norm("conv1dto2dLayer_tfan", 5120);
  pt.conv("convLayer.0", 512, 256, 25);   // == upSample2.0
  pt.norm("convLayer.2", 128);            // == upSample2.2
  pt.conv("upSample1.0", 1024, 256, 25);
  pt.norm("upSample1.2", 256);
  pt.conv("lastConvLayer", 1, 128, 75);

  DescBuilder b(pt);
  // stem: the 15 horizontal taps AND pairs of vertical taps are folded into the operand's channels
  // (60 of 64 used), leaving 3 vertical taps at rows h-2, h, h+2 (see prep_g_kernel)
  b.conv("stem", {"conv1", "conv1_gates"}, 128, 2, 75, kPackStemG, kVecIdent, 256, 64, 3, 64);
  b.conv("ds1", {"downSample1.convLayer.0", "downSample1.convLayer_gates.0"}, 256, 128, 25,
         kPackStd, kVecIdent, 512, 128, 25, 128);
  b.conv("ds2", {"downSample2.convLayer.0", "downSample2.convLayer_gates.0"}, 256, 256, 25,
         kPackStd, kVecIdent, 512, 256, 25, 256);
  b.conv("2dto1d", {"conv2dto1dLayer"}, 256, 5120, 1, kPack2dTo1d, kVecIdent, 256, 256, 20, 256);
  for (int i = 1; i <= 6; ++i) {
    const std::string p = "residualLayer" + std::to_string(i);
    b.conv("res" + std::to_string(i) + "a", {p + ".conv1d_layer.0", p + ".conv_layer_gates.0"}, 512,
           256, 3, kPackStd, kVecIdent, 1024, 256, 3, 256);
    b.conv("res" + std::to_string(i) + "b", {p + ".conv1d_out_layer.0"}, 256, 512, 3, kPackStd,
           kVecIdent, 256, 512, 3, 512);
  }
  b.conv("1dto2d", {"conv1dto2dLayer"}, 5120, 256, 1, kPack1dTo2d, kVecHC20, 5120, 256, 1, 256);
  b.conv("up1", {"upSample1.0"}, 1024, 256, 25, kPackShuffle, kVecShuffle, 1024, 256, 25, 256);
  b.conv("up2", {"convLayer.0"}, 512, 256, 25, kPackShuffle, kVecShuffle, 512, 256, 25, 256);
  b.conv("head", {"lastConvLayer"}, 1, 128, 75, kPackHead, kVecIdent, 128, 128, 1, 128);

  b.norm("ds1", {"downSample1.convLayer.1", "downSample1.convLayer_gates.1"}, 256, kVecIdent);
  b.norm("ds2", {"downSample2.convLayer.1", "downSample2.convLayer_gates.1"}, 256, kVecIdent);
  b.norm("2dto1d", {"conv2dto1dLayer_tfan"}, 256, kVecIdent);
  for (int i = 1; i <= 6; ++i) {
    const std::string p = "residualLayer" + std::to_string(i);
    b.norm("res" + std::to_string(i) + "a", {p + ".conv1d_layer.1", p + ".conv_layer_gates.1"}, 512,
           kVecIdent);
    b.norm("res" + std::to_string(i) + "b", {p + ".conv1d_out_layer.1"}, 256, kVecIdent);
  }
  b.norm("1dto2d", {"conv1dto2dLayer_tfan"}, 5120, kVecHC20);
  b.norm("up1", {"upSample1.2"}, 256, kVecIdent);
  b.norm("up2", {"convLayer.2"}, 128, kVecIdent);
  return b.d;
}

ModelDesc build_discriminator() {
  ParamTable pt;
  pt.conv("convLayer1.0", 128, 1, 9);
  pt.conv("downSample1.0", 256, 128, 9);
  pt.norm("downSample1.1", 256);
  pt.conv("downSample2.0", 512, 256, 9);
  pt.norm("downSample2.1", 512);
  pt.conv("downSample3.0", 1024, 512, 9);
  pt.norm("downSample3.1", 1024);
  pt.conv("downSample4.0", 1024, 1024, 10);  // never used in forward (model.py:316-320 vs :340-349)
  pt.norm("downSample4.1", 1024);
  pt.conv("outputConvLayer.0", 1, 1024, 3);

  DescBuilder b(pt);
  b.conv("stem", {"convLayer1.0"}, 128, 1, 9, kPackStemD, kVecIdent, 128, 64, 1, 64);
  b.conv("ds1", {"downSample1.0"}, 256, 128, 9, kPackStd, kVecIdent, 256, 128, 9, 128);
  b.conv("ds2", {"downSample2.0"}, 512, 256, 9, kPackStd, kVecIdent, 512, 256, 9, 256);
  b.conv("ds3", {"downSample3.0"}, 1024, 512, 9, kPackStd, kVecIdent, 1024, 512, 9, 512);
  b.conv("head", {"outputConvLayer.0"}, 1, 1024, 3, kPackHead, kVecIdent, 128, 1024, 1, 1024);
  b.norm("ds1", {"downSample1.1"}, 256, kVecIdent);
  b.norm("ds2", {"downSample2.1"}, 512, kVecIdent);
  b.norm("ds3", {"downSample3.1"}, 1024, kVecIdent);
  b.d.deadBegin = pt.at("downSample4.0.weight");
  b.d.deadLen = pt.at("outputConvLayer.0.weight") - b.d.deadBegin;
  return b.d;
}

}  // namespace

const ModelDesc& generator_desc() {
  static const ModelDesc d = build_generator();
  return d;
}
const ModelDesc& discriminator_desc() {
  static const ModelDesc d = build_discriminator();
  return d;
}

// ================================================================================================
// helpers
namespace {

struct Run {
  RunCfg rc;
  bool ok = true;
  bool forked = false;
  float* tailScratch = nullptr;   // kTailScratchFloats floats of workspace for the conv kernels' tail split
  // weight-gradient GEMMs go to a side stream: they only feed the gradient blob, so they overlap
  // with the bandwidth-bound layer kernels and data-gradient convs of the main chain
  cudaStream_t wgrad_stream() {
    if (!rc.side || !rc.forkEvent) return rc.stream;
    cudaEventRecord(rc.forkEvent, rc.stream);   // everything the wgrad reads has been enqueued
    cudaStreamWaitEvent(rc.side, rc.forkEvent, 0);
    forked = true;
    return rc.side;
  }
  void join() {
    if (forked && rc.side) {
      cudaEventRecord(rc.forkEvent, rc.side);
      cudaStreamWaitEvent(rc.stream, rc.forkEvent, 0);
      forked = false;
    }
  }
  void check(cudaError_t e, const char* what) {
    if (e != cudaSuccess && ok) {
      ok = false;
      if (!last_error()[0] || e != cudaErrorInvalidValue)
        set_error("%s: %s", what, cudaGetErrorString(e));
    }
  }
};

struct Arena {
  uint8_t* base;
  long long off = 0;
  std::vector<SavedEntry>* layout = nullptr;
  explicit Arena(void* b) : base(reinterpret_cast<uint8_t*>(b)) {}
  void* take(long long bytes, const char* name = nullptr) {
    off = align_up(off, 256);
    void* p = base ? base + off : nullptr;
    if (layout && name) layout->push_back(SavedEntry{name, off, bytes});
    off += bytes;
    return p;
  }
  template <class T>
  T* takeT(long long count, const char* name = nullptr) {
    return reinterpret_cast<T*>(take(count * (long long)sizeof(T), name));
  }
};

// Conv-epilogue statistics accumulators: every normalised layer of a forward call gets its own
// slice (sums, then sums of squares) of one pool that is zero-filled with a single memset.
struct StatPool {
  float* base;
  long long off = 0;
  explicit StatPool(float* b) : base(b) {}
  float* take(long long imgs, long long nz) {
    float* p = base + off;
    off += 2 * imgs * nz;
    return p;
  }
  float* take_raw(long long n) {
    float* p = base + off;
    off += n;
    return p;
  }
};
constexpr long long kPoolExtra = 256;   // C8: 4 floats of dz-scale maxima per normalised layer
long long gen_stat_pool_floats(int B) { return 2LL * B * (512 + 512 + 256 + 6 * (1024 + 256) + 5120 + 1024 + 512) + kPoolExtra; }
long long dis_stat_pool_floats(int B) { return 2LL * B * (256 + 512 + 1024) + kPoolExtra; }

struct Weights {   // views into the packed blob
  const __nv_bfloat16* bf;
  const float* f32;
  int c8;
  const ModelDesc& d;
  Weights(const void* packed, const ModelDesc& md, int c8mode)
      : bf(reinterpret_cast<const __nv_bfloat16*>(packed)),
        f32(reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(packed) + md.packedBf16 * 2)),
        c8(c8mode), d(md) {}
  bool is_c8(const ConvDesc& c) const { return c8 && c.c8Eligible; }
  // operand over the forward layout [T][N][K] / the data-gradient layout viewed as [T][N=Cd][K=Np]
  // (or any other (K, N, T) view of the same storage, for the two flatten layers)
  WgtOperand view(const ConvDesc& c, bool dgrad, int K, int N, int T) const {
    const long long hi = dgrad ? c.dHi : c.fHi, lo = dgrad ? c.dLo : c.fLo, el = dgrad ? c.dElems : c.fElems;
    WgtOperand w{bf + hi, bf + lo, K, N, T};
    if (is_c8(c)) {
      w.h8 = reinterpret_cast<const uint8_t*>(bf + lo);
      w.l8 = reinterpret_cast<const uint8_t*>(bf + lo) + el;
      w.rec = f32 + c.c8Rec;
    }
    return w;
  }
  WgtOperand fwd(const ConvDesc& c) const { return view(c, false, c.Cp, c.Np, c.Tp); }
  WgtOperand bwd(const ConvDesc& c) const { return view(c, true, c.Np, c.Cd, c.Tp); }
  const float* bias(const ConvDesc& c) const { return f32 + c.biasEng; }
  const float* gamma(const NormDesc& n) const { return f32 + n.gammaEng; }
  const float* beta(const NormDesc& n) const { return f32 + n.betaEng; }
};

struct TapList {
  int n = 0;
  Tap t[kMaxTaps];
  void add(int dx, int dy, int plane, int w) {
    t[n].dx = (int8_t)dx; t[n].dy = (int8_t)dy; t[n].plane = (uint8_t)plane; t[n].w = (uint8_t)w;
    ++n;
  }
};
// stride-1 KHxKW taps (forward, or data-gradient when sign = -1)
TapList taps_s1(int KH, int KW, int padH, int padW, int sign) {
  TapList l;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) l.add(sign * (kw - padW), sign * (kh - padH), 0, kh * KW + kw);
  return l;
}
// Generator stem over the folded operand: 3 taps at input rows -2, 0, +2 (each holds kernel rows 2t, 2t+1)
// (operand row r = input row + 1: the operand has 81 rows, see prep_g_kernel)
TapList taps_stem_g(int sign) {
  TapList l;
  for (int t = 0; t < 3; ++t) l.add(0, sign * (2 * t - 1), 0, t);
  return l;
}
// stride-2 KxK forward taps over a parity-split input
TapList taps_s2_fwd(int K, int pad) {
  TapList l;
  for (int kh = 0; kh < K; ++kh)
    for (int kw = 0; kw < K; ++kw) {
      const int oy = kh - pad, ox = kw - pad;
      const int ph = ((oy % 2) + 2) % 2, pw = ((ox % 2) + 2) % 2;
      l.add((ox - pw) / 2, (oy - ph) / 2, ph * 2 + pw, kh * K + kw);
    }
  return l;
}
// stride-2 data-gradient taps producing input parity plane (ph, pw): reads dz at y' + (ph+pad-kh)/2
TapList taps_s2_bwd(int K, int pad, int ph, int pw) {
  TapList l;
  for (int kh = 0; kh < K; ++kh) {
    if (((ph + pad - kh) % 2 + 2) % 2) continue;
    for (int kw = 0; kw < K; ++kw) {
      if (((pw + pad - kw) % 2 + 2) % 2) continue;
      l.add((pw + pad - kw) / 2, (ph + pad - kh) / 2, 0, kh * K + kw);
    }
  }
  return l;
}
TapList taps_rows(int n, int sign) {  // n taps along Y (2D<->1D flatten), weight slice = row
  TapList l;
  for (int h = 0; h < n; ++h) l.add(0, sign * h, 0, h);
  return l;
}
TapList taps_one() {
  TapList l;
  l.add(0, 0, 0, 0);
  return l;
}

struct OutAddr {
  float* out;
  long long sB, sY, sX;
  int nSplit;
  long long sNhi;
};
OutAddr plain_out(float* out, int Y, int X, int N) {
  return OutAddr{out, (long long)Y * X * N, (long long)X * N, (long long)N, N, 0};
}

ConvGeom conv_geom(Run& r, const ActOperand& a, const WgtOperand& w, const TapList& taps, int oB, int oY,
                   int oX, const OutAddr& o, const float* bias, const float* addsrc, const char* what,
                   double algoFrac) {
  ConvGeom g{};
  g.a = a;
  g.w = w;
  g.oX = oX; g.oY = oY; g.oB = oB;
  if (!choose_box(oB, oY, oX, kTileM, &g.BX, &g.BY, &g.BB)) { r.ok = false; set_error("%s: box", what); return g; }
  g.tilesX = (oX + g.BX - 1) / g.BX;
  g.tilesY = (oY + g.BY - 1) / g.BY;
  g.tilesB = (oB + g.BB - 1) / g.BB;
  g.nTaps = taps.n;
  g.cBlocks = a.C / kBlockK;
  for (int i = 0; i < taps.n; ++i) g.taps[i] = taps.t[i];
  g.nGroups = 1; g.grpTapStart[0] = 0; g.grpTapCount[0] = taps.n; g.grpOutOff[0] = 0;
  g.sB = o.sB; g.sY = o.sY; g.sX = o.sX; g.nSplit = o.nSplit; g.sNhi = o.sNhi;
  g.out = o.out; g.bias = bias; g.addsrc = addsrc;
  g.nPass = r.rc.nPass;
  g.algoFlops = 2.0 * oB * oY * oX * (double)w.N * taps.n * a.C * algoFrac;
  g.statSum = nullptr; g.statSq = nullptr;
  g.statSeg = g.BX * g.BY >= 32 ? 32 : g.BX * g.BY;
  g.kSplit = 1;
  return g;
}

// C8 operands (fp16 + 2 x e4m3 planes on both sides) run on the C8 kernels; everything else on the
// split-bf16 / bf16 kernels.
bool is_c8(const ConvGeom& g) { return g.a.h8 != nullptr && g.w.h8 != nullptr; }
// ... on the C8 kernels (fp16 + 2 x e4m3), unless this is a C8H backward call (single fp16 pass)
bool on_c8_kernel(const Run& r, const ConvGeom& g) { return is_c8(g) && !r.rc.half16; }
// C8 pair kernel tile width: 256-wide tiles when they still fill the chip, else 128-wide (two TMEM
// buffer pairs, epilogue overlapped)
long long c8_pair_tiles(const ConvGeom& g, int bn) {
  return (((long long)g.tilesX * g.tilesY * g.tilesB + 1) / 2) * (g.w.N / bn) * g.nGroups;
}
int c8_block_n(const ConvGeom& g) {
  static int minWide = -1;   // MCGVC_C8_WIDE_MIN: fewest 256-wide pair tiles for which the wide kernel is used
  if (minWide < 0) { const char* e = getenv("MCGVC_C8_WIDE_MIN"); minWide = e ? atoi(e) : 74; }
  const bool wide = g.w.N % 256 == 0 && g.nSplit % 256 == 0 && c8_pair_tiles(g, 256) >= minWide;
  return wide ? 256 : 128;
}
// Tail split of the partial last wave (pair kernels, gemm_types.cuh): tried before the uniform split-K
bool try_tail(Run& r, ConvGeom& g) {
  if (r.rc.backend != 0 || !r.tailScratch || g.kSplit > 1) return false;
  return conv_plan_tail(g, on_c8_kernel(r, g) ? c8_block_n(g) : 0, r.tailScratch);
}

cudaError_t launch_conv_any(Run& r, ConvGeom& g) {
  if (is_c8(g) && r.rc.half16) {
    g.nPass = 1;
    g.half16 = 1;
    g.c8OutScale = 1.f;
    g.c8RecA = g.a.rec;
    g.c8RecW = g.w.rec;
    return r.rc.backend == 0 ? launch_conv_tc(g, r.rc.stream) : launch_conv_simt(g, r.rc.stream);
  }
  if (is_c8(g)) {
    g.c8OutScale = 1.f;
    g.c8CorrScale = 1.f / 2048.f;   // 2^-11: pre-scale of the residual planes
    g.c8RecA = g.a.rec;
    g.c8RecW = g.w.rec;
    g.mainBf16 = 0;
    if (r.rc.backend != 0) { g.kSplit = 1; return launch_conv_c8_simt(g, r.rc.stream); }
    return launch_conv_c8(g, c8_block_n(g), r.rc.stream);
  }
  if ((g.a.h8 != nullptr) != (g.w.h8 != nullptr)) {
    set_error("conv: activation and weight operands disagree on the C8 format");
    return cudaErrorInvalidValue;
  }
  return r.rc.backend == 0 ? launch_conv_tc(g, r.rc.stream) : launch_conv_simt(g, r.rc.stream);
}

// Split-K (tensor-core backend only): when the planner finds that K-slices fill the SM waves
// better, zero-fill the `splitFloats` floats at g.out and let the slices add into it.
void plan_split(Run& r, ConvGeom& g, long long splitFloats, double minGain, const char* what) {
  if (!r.ok || splitFloats <= 0 || r.rc.backend != 0) return;
  // 256-wide C8 tiles cannot overlap their epilogue with the next item's MMAs (D1 + D2 fill TMEM), so
  // the red.add epilogues of K slices land on the critical path: measured 124.9 -> 121.4 ms per step
  // with split-K off in C8 mode.  Only the 128-wide (double-buffered) C8 tiles may split.
  if (on_c8_kernel(r, g) && c8_block_n(g) == 256) return;
  const int s = on_c8_kernel(r, g) ? plan_ksplit_waves(c8_pair_tiles(g, 128), 74, g, minGain) : conv_plan_ksplit(g, minGain);
  if (s <= 1) return;
  r.check(cudaMemsetAsync(g.out, 0, (size_t)splitFloats * sizeof(float), r.rc.stream), what);
  g.kSplit = s;
}

// `splitFloats` > 0: the output is a dense region of that many floats starting at o.out and the
// convolution may run split-K (data gradients: no statistics to fuse).
void run_conv(Run& r, const ActOperand& a, const WgtOperand& w, const TapList& taps, int oB, int oY,
              int oX, const OutAddr& o, const float* bias, const float* addsrc, const char* what,
              double algoFrac = 1.0, float* statSum = nullptr, float* statSq = nullptr,
              long long splitFloats = 0) {
  if (!r.ok) return;
  ConvGeom g = conv_geom(r, a, w, taps, oB, oY, oX, o, bias, addsrc, what, algoFrac);
  if (!r.ok) return;
  g.statSum = r.rc.backend == 0 ? statSum : nullptr;
  g.statSq = statSq;
  if (!try_tail(r, g) && !g.statSum) plan_split(r, g, splitFloats, 0.08, what);
  r.check(launch_conv_any(r, g), what);
}


void run_wgrad(Run& r, const ActOperand& dz, const ActOperand& x, const TapList& xtaps,
               const TapList* ztaps, int pB, int pY, int pX, float* dw, const char* what,
               double algoFrac = 1.0) {
  if (!r.ok) return;
  WgradGeom g{};
  g.dz = dz;
  g.x = x;
  g.pX = pX; g.pY = pY; g.pB = pB;
  if (!choose_box(pB, pY, pX, 64, &g.BX, &g.BY, &g.BB)) { r.ok = false; set_error("%s: box", what); return; }
  g.tilesX = (pX + g.BX - 1) / g.BX;
  g.tilesY = (pY + g.BY - 1) / g.BY;
  g.tilesB = (pB + g.BB - 1) / g.BB;
  g.nTaps = xtaps.n;
  for (int i = 0; i < xtaps.n; ++i) {
    g.taps[i] = xtaps.t[i];
    if (ztaps) g.ztaps[i] = ztaps->t[i];
    else g.ztaps[i] = Tap{0, 0, 0, 0};
  }
  g.N = dz.C;
  g.C = x.C;
  // channel tile: the widest that divides C (256-wide tiles halve the dz re-reads per MMA)
  g.cTile = x.C % 256 == 0 ? 256 : (x.C % 128 == 0 ? 128 : 64);
  {
    static int envct = -1;  // tuning override: MCGVC_WGRAD_CTILE=128 caps the channel tile
    if (envct < 0) { const char* e = getenv("MCGVC_WGRAD_CTILE"); envct = e ? atoi(e) : 0; }
    if (envct == 128 && g.cTile == 256) g.cTile = 128;
  }
  // split-K: pick the split whose CTA count fills whole waves of the 148 SMs best (a partial last
  // wave costs a full wave of time), with a mild penalty per extra split for the atomic merge
  const long long units = (long long)xtaps.n * (g.N / 128) * (g.C / g.cTile);
  const long long posTiles = (long long)g.tilesX * g.tilesY * g.tilesB;
  long long maxSplit = posTiles / 8;
  if (maxSplit > 148) maxSplit = 148;
  if (maxSplit < 1) maxSplit = 1;
  int bestSk = 1;
  double bestCost = 1e30;
  for (long long sk = 1; sk <= maxSplit; ++sk) {
    const double waves = (double)(units * sk) / 148.0;
    const double rounds = (double)((units * sk + 147) / 148);
    const double cost = rounds / waves * (1.0 + 0.01 * (double)sk) + (waves < 1.0 ? 1.0 / waves : 0.0);
    if (cost < bestCost - 1e-9) { bestCost = cost; bestSk = (int)sk; }
  }
  g.splitK = bestSk;
  g.dw = dw;
  g.nPass = r.rc.nPass;
  g.algoFlops = 2.0 * pB * pY * pX * (double)g.N * g.C * xtaps.n * algoFrac;
  cudaStream_t ws = r.wgrad_stream();
  // single-pass pair kernel: two taps per work item share one staged dz k-block (wgrad_gemm.cu)
  auto pair_taps = [&]() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("MCGVC_WGRAD_TAP_PAIR"); on = e ? atoi(e) : 1; }
    if (!on || r.rc.backend != 0 || ztaps || g.nPass != 1 || g.N % 256 || g.cTile < 128 || xtaps.n < 2) return;
    g.tapPair = 1;
    const long long unitsP = (long long)((xtaps.n + 1) / 2) * (g.N / 256) * (g.C / g.cTile);
    int best = 1;
    double bestC = 1e30;
    for (long long sk = 1; sk <= maxSplit; ++sk) {
      const double waves = (double)(unitsP * sk) / 74.0;
      const double rounds = (double)((unitsP * sk + 73) / 74);
      const double cost = rounds / waves * (1.0 + 0.01 * (double)sk) + (waves < 1.0 ? 1.0 / waves : 0.0);
      if (cost < bestC - 1e-9) { bestC = cost; best = (int)sk; }
    }
    g.splitK = best;
  };
  if (dz.h8 && x.h8 && r.rc.wgradHalf16) {   // C8H / C8W: one fp16 pass over the 16-bit planes
    g.nPass = 1;
    g.half16 = 1;
    g.c8OutScale = 1.f;
    g.c8RecZ = dz.rec;
    g.c8RecX = x.rec;
    pair_taps();
    r.check(r.rc.backend == 0 ? launch_wgrad_tc(g, ws) : launch_wgrad_simt(g, ws), what);
    return;
  }
  if (dz.h8 && x.h8) {
    g.cTile = x.C % 256 == 0 ? 256 : 128;
    g.c8OutScale = 1.f;
    g.c8CorrScale = 1.f / 2048.f;
    g.c8RecZ = dz.rec;
    g.c8RecX = x.rec;
    g.mainBf16 = 0;
    if (g.N % 256 || x.C % 128) { r.ok = false; set_error("%s: C8 wgrad needs N %% 256 == 0 and C %% 128 == 0", what); return; }
    // wave-aware split-K for the pair tiling of the C8 kernel
    const long long unitsC8 = (long long)xtaps.n * (g.N / 256) * (g.C / g.cTile);
    int best = 1;
    double bestC = 1e30;
    for (long long sk = 1; sk <= maxSplit; ++sk) {
      const double waves = (double)(unitsC8 * sk) / 74.0;
      const double rounds = (double)((unitsC8 * sk + 73) / 74);
      const double cost = rounds / waves * (1.0 + 0.01 * (double)sk) + (waves < 1.0 ? 1.0 / waves : 0.0);
      if (cost < bestC - 1e-9) { bestC = cost; best = (int)sk; }
    }
    g.splitK = best;
    r.check(r.rc.backend == 0 ? launch_wgrad_c8(g, ws) : launch_wgrad_c8_simt(g, ws), what);
    return;
  }
  if ((dz.h8 != nullptr) != (x.h8 != nullptr)) { r.ok = false; set_error("%s: dz and x disagree on the C8 format", what); return; }
  pair_taps();   // mixed / fast modes (nPass = 1)
  r.check(r.rc.backend == 0 ? launch_wgrad_tc(g, ws) : launch_wgrad_simt(g, ws), what);
}

struct BfPair {
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  long long elems;     // elements per plane
  int c8;              // C8 planes (fp16 in hi, two e4m3 planes in lo) instead of split-bf16
  const float* rec;    // C8: device-side scale record {1/S, 1/E} the consuming GEMMs read
};
BfPair take_pair(Arena& a, long long elems, const char* name, int c8 = 0, const float* rec = nullptr) {
  BfPair p;
  p.hi = a.takeT<__nv_bfloat16>(elems, name);
  p.lo = a.takeT<__nv_bfloat16>(elems);
  p.elems = elems;
  p.c8 = c8;
  p.rec = rec;
  return p;
}
PlaneFmt plane_fmt(const BfPair& p) { return PlaneFmt{p.c8, 1.f, 2.f, p.elems}; }   // activations: S = 1, E = 2
void set_c8(ActOperand& o, const BfPair& p) {
  if (!p.c8) return;
  o.h8 = reinterpret_cast<const uint8_t*>(p.lo);
  o.l8 = reinterpret_cast<const uint8_t*>(p.lo) + p.elems;
  o.rec = p.rec;
}
ActOperand plain_op(const BfPair& p, int B, int Y, int X, int C) {
  ActOperand o{p.hi, p.lo, C, X, Y, 1, B};
  set_c8(o, p);
  return o;
}
ActOperand parity_op(const BfPair& p, int B, int Y, int X, int C) {
  ActOperand o{p.hi, p.lo, C, (X + 1) / 2, (Y + 1) / 2, 4, B};
  set_c8(o, p);
  return o;
}
long long parity_elems(int B, int Y, int X, int C) {
  return (long long)B * 4 * ((Y + 1) / 2) * ((X + 1) / 2) * C;
}

struct Stat {
  float* mean;
  float* rstd;
};
Stat take_stat(Arena& a, long long n) {
  Stat s;
  s.mean = a.takeT<float>(n);
  s.rstd = a.takeT<float>(n);
  return s;
}

// Data gradient of a stride-2 KxK convolution: the four input-parity planes are four stride-1
// convolutions over dz with disjoint tap subsets; they run as ONE launch (tile groups) so that
// their tiles fill the SMs together.  dIn is the parity-split fp32 gradient [B][4][Yp][Xp][Cin].
void run_dgrad_s2(Run& r, const ActOperand& dz, const WgtOperand& w, int K, int pad, int B, int Yp, int Xp,
                  int Cin, float* dIn, const char* what) {
  if (!r.ok) return;
  ConvGeom g{};
  g.a = dz;
  g.w = w;
  g.oX = Xp; g.oY = Yp; g.oB = B;
  if (!choose_box(B, Yp, Xp, kTileM, &g.BX, &g.BY, &g.BB)) { r.ok = false; set_error("%s: box", what); return; }
  g.tilesX = (Xp + g.BX - 1) / g.BX;
  g.tilesY = (Yp + g.BY - 1) / g.BY;
  g.tilesB = (B + g.BB - 1) / g.BB;
  g.cBlocks = dz.C / kBlockK;
  g.nGroups = 4;
  int nt = 0;
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      const int p = ph * 2 + pw;
      const TapList tl = taps_s2_bwd(K, pad, ph, pw);
      g.grpTapStart[p] = nt;
      g.grpTapCount[p] = tl.n;
      g.grpOutOff[p] = (long long)p * Yp * Xp * Cin;
      for (int i = 0; i < tl.n; ++i) g.taps[nt++] = tl.t[i];
    }
  g.nTaps = nt;
  g.sB = (long long)4 * Yp * Xp * Cin; g.sY = (long long)Xp * Cin; g.sX = Cin;
  g.nSplit = Cin; g.sNhi = 0;
  g.out = dIn; g.bias = nullptr; g.addsrc = nullptr;
  g.nPass = r.rc.nPass;
  g.algoFlops = 2.0 * B * Yp * Xp * (double)w.N * nt * dz.C;
  g.statSum = nullptr; g.statSq = nullptr; g.statSeg = 32;
  g.kSplit = 1;
  if (!try_tail(r, g)) plan_split(r, g, (long long)B * 4 * Yp * Xp * Cin, 0.08, what);
  r.check(launch_conv_any(r, g), what);
}

// Convolution whose output feeds an InstanceNorm: the statistics come out of the conv epilogue
// (tcgen05 backend) or from the stand-alone statistics kernel (SIMT checking backend, or
// MCGVC_FUSED_STATS=0).  statImgs x statNz is the statistics grid over the conv's [oB][N] output.
bool fused_stats_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MCGVC_FUSED_STATS"); v = e ? atoi(e) : 1; }
  return v != 0;
}
void run_conv_in(Run& r, const ActOperand& a, const WgtOperand& w, const TapList& taps, int oB, int oY,
                 int oX, const OutAddr& o, const float* bias, const char* what, float* ssum, float* /*unused*/,
                 int statImgs, int statNz, int groups, int planePositions, const Stat& st, const float* z,
                 long long splitFloats = 0) {
  if (!r.ok) return;
  float* ssq = ssum + (size_t)statImgs * statNz;   // sums of squares right behind the sums
  bool fused = r.rc.backend == 0 && fused_stats_enabled();
  if (fused && splitFloats > 0) {
    // layers a little over one SM wave (Discriminator ds3: 80 pair tiles on 74 pairs): split-K plus
    // one stand-alone statistics pass over z beats paying a second, nearly empty round
    ConvGeom g = conv_geom(r, a, w, taps, oB, oY, oX, o, bias, nullptr, what, 1.0);
    // C8 layers and layers whose partial last wave the tail split handles keep the fused statistics
    if (r.ok && !is_c8(g) && !try_tail(r, g) && conv_plan_ksplit(g, 0.2) > 1) fused = false;
  }
  if (!fused && r.rc.backend == 0 && splitFloats > 0) {
    run_conv(r, a, w, taps, oB, oY, oX, o, bias, nullptr, what, 1.0, nullptr, nullptr, splitFloats);
    if (r.ok) r.check(launch_stats(z, statNz, planePositions, statImgs, groups, st.mean, st.rstd, r.rc.stream), what);
    return;
  }
  if (fused) {
    // ssum / ssq: this layer's own slice of the accumulator pool the caller zero-filled up front
    run_conv(r, a, w, taps, oB, oY, oX, o, bias, nullptr, what, 1.0, ssum, ssq);
    if (r.ok) r.check(launch_stats_finalize(ssum, ssq, statImgs, statNz, groups, planePositions, st.mean, st.rstd, r.rc.stream), what);
  } else {
    run_conv(r, a, w, taps, oB, oY, oX, o, bias, nullptr, what);
    if (r.ok) r.check(launch_stats(z, statNz, planePositions, statImgs, groups, st.mean, st.rstd, r.rc.stream), what);
  }
}

ApplyArgs mk_apply(int mode, const float* z, int Nz, int zY, int zX, const Stat& st, int Nstat,
                   const float* gamma, const float* beta, int affPeriod, const float* residual,
                   ActBuf out) {
  ApplyArgs a{};
  a.mode = mode; a.z = z; a.Nz = Nz; a.zY = zY; a.zX = zX;
  a.mean = st.mean; a.rstd = st.rstd; a.Nstat = Nstat;
  a.gamma = gamma; a.beta = beta; a.affPeriod = affPeriod;
  a.residual = residual; a.out = out;
  return a;
}
ApplyBwdArgs mk_bwd(int mode, const float* z, int Nz, int zY, int zX, const Stat& st, int Nstat,
                    const float* gamma, const float* beta, int affPeriod, ActBuf dA, StatPool& tp,
                    float* dgamma, float* dbeta, BfPair dz, float* dbias) {
  ApplyBwdArgs a{};
  a.mode = mode; a.z = z; a.Nz = Nz; a.zY = zY; a.zX = zX;
  a.mean = st.mean; a.rstd = st.rstd; a.Nstat = Nstat;
  a.gamma = gamma; a.beta = beta; a.affPeriod = affPeriod;
  // this layer's slice of the reduction pool (sum dy, then sum dy*xhat), zero-filled once per call
  a.dA = dA; a.t1 = tp.take(dA.nImg, Nstat); a.t2 = a.t1 + (size_t)dA.nImg * Nstat;
  a.prezeroed = 1;
  a.dgamma = dgamma; a.dbeta = dbeta;
  a.dz_hi = dz.hi; a.dz_lo = dz.lo; a.dbias = dbias;
  a.dzFmt = PlaneFmt{dz.c8, 1.f, 1.f, dz.elems};
  a.mx = dz.c8 ? reinterpret_cast<unsigned int*>(tp.take_raw(4)) : nullptr;
  a.dzRec = const_cast<float*>(dz.rec);
  return a;
}
void run_bwd(Run& r, const ApplyBwdArgs& a, const char* what) {
  if (!r.ok) return;
  r.check(launch_apply_bwd_reduce(a, r.rc.stream), what);
  r.check(launch_apply_bwd(a, r.rc.stream), what);
}

}  // namespace

// ================================================================================================
// packing
namespace {
PackTable make_pack_table(const ModelDesc& d, int c8 = 0) {
  PackTable t{};
  t.actRec = c8 ? (int)d.actRec : -1;
  for (const ConvDesc& c : d.convs)
    for (int p = 0; p < c.nParts; ++p) {
      PackEntry& e = t.e[t.count++];
      e.kind = c.kind; e.N = c.refN; e.C = c.refC; e.T = c.refT; e.nOffset = p * c.refN;
      e.Np = c.Np; e.Cp = c.Cp; e.Tp = c.Tp; e.Cd = c.Cd;
      e.refOff = (int)c.wOff[p];
      e.fHi = (int)c.fHi; e.fLo = (int)c.fLo;
      e.dHi = c.Cd ? (int)c.dHi : -1; e.dLo = c.Cd ? (int)c.dLo : -1;
      e.gW = (int)c.gW;
      e.c8 = c8 && c.c8Eligible;
      e.rec = (int)c.c8Rec;
      e.fElems = (int)c.fElems;
      e.dElems = (int)c.dElems;
    }
  return t;
}
// small vectors: engine offsets are into the packed fp32 area (pack) or the gradient blob (unpack)
VecTable make_vec_table(const ModelDesc& d, bool grads) {
  VecTable t{};
  for (const ConvDesc& c : d.convs)
    for (int p = 0; p < c.nParts; ++p)
      t.e[t.count++] = VecEntry{c.biasKind, c.refN, (int)c.bOff[p], (int)((grads ? c.gB : c.biasEng) + p * c.refN)};
  for (const NormDesc& n : d.norms)
    for (int p = 0; p < n.nParts; ++p) {
      t.e[t.count++] = VecEntry{n.vecKind, n.n, (int)n.gOff[p], (int)((grads ? n.gGamma : n.gammaEng) + p * n.n)};
      t.e[t.count++] = VecEntry{n.vecKind, n.n, (int)n.bOff[p], (int)((grads ? n.gBeta : n.betaEng) + p * n.n)};
    }
  return t;
}
}  // namespace

int pack_model(const ModelDesc& d, const float* params, void* packed, const RunCfg& rc) {
  Run r{rc};
  __nv_bfloat16* bf = reinterpret_cast<__nv_bfloat16*>(packed);
  float* f32 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(packed) + d.packedBf16 * 2);
  // padding elements (unused taps / channels) must be zero
  r.check(launch_fill_zero(packed, (size_t)d.packed_bytes(), rc.stream), "pack: zero");
  const PackTable pt = make_pack_table(d, rc.c8);
  const VecTable vt = make_vec_table(d, false);
  if (pt.count > 32 || vt.count > 96) { set_error("pack tables too small"); return 1; }
  r.check(launch_pack_weights_table(pt, params, bf, f32, rc.stream), "pack: weights");
  r.check(launch_pack_vecs_table(vt, params, f32, rc.stream), "pack: vectors");
  return r.ok ? 0 : 1;
}

int unpack_grads(const ModelDesc& d, const float* gblob, float* gradFlat, const RunCfg& rc, int live, float scale) {
  Run r{rc};
  PackTable pt = make_pack_table(d);
  VecTable vt = make_vec_table(d, true);
  if (pt.count > 32 || vt.count > 96) { set_error("pack tables too small"); return 1; }
  pt.scale = vt.scale = scale;
  if (live && d.deadLen > 0) {   // offsets behind the dead range move down by its length
    for (int i = 0; i < pt.count; ++i) if (pt.e[i].refOff >= d.deadBegin) pt.e[i].refOff -= (int)d.deadLen;
    for (int i = 0; i < vt.count; ++i) if (vt.e[i].refOff >= d.deadBegin) vt.e[i].refOff -= (int)d.deadLen;
  }
  r.check(launch_unpack_wgrads_table(pt, gblob, gradFlat, rc.stream), "unpack: weights");
  r.check(launch_unpack_vecs_table(vt, gblob, gradFlat, rc.stream), "unpack: vectors");
  return r.ok ? 0 : 1;
}

// ================================================================================================
// Generator
namespace {

struct GenDims {
  int B, T, W1, W2, X1, X2;  // X1 = 2*W2 (up1 output width), X2 = 4*W2 (output frames)
  GenDims(int b, int t) : B(b), T(t) {
    W1 = (T + 1) / 2;
    W2 = (W1 + 1) / 2;
    X1 = 2 * W2;
    X2 = 4 * W2;
  }
};

struct GenSaved {
  BfPair X15; float* z0;
  BfPair A0; float* z1; Stat st1;
  BfPair A1; float* z2; Stat st2;
  BfPair A2; float* z3; Stat st3;
  float* Rf[7]; BfPair R[7];
  float* z4[6]; Stat st4[6]; BfPair H[6]; float* z5[6]; Stat st5[6];
  float* z6; Stat st6; BfPair U0;
  float* z7; Stat st7; BfPair U1;
  float* z8; Stat st8; BfPair U2;
  long long total;
};

const float* act_rec_of(const void* packed, const ModelDesc& md) {
  return reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(packed) + md.packedBf16 * 2) + md.actRec;
}

// c8: the operands of every layer but the stem and the head are kept as C8 planes (same bytes)
GenSaved plan_gen_saved(const GenDims& d, void* base, std::vector<SavedEntry>* layout, int c8 = 0,
                        const float* rec = nullptr) {
  Arena a(base);
  a.layout = layout;
  GenSaved s{};
  const long long M0 = (long long)d.B * 80 * d.T, M1 = (long long)d.B * 40 * d.W1,
                  M2 = (long long)d.B * 20 * d.W2, L = (long long)d.B * d.W2,
                  M7 = (long long)d.B * 40 * d.X1, M8 = (long long)d.B * 80 * d.X2;
  s.X15 = take_pair(a, (long long)d.B * 81 * d.T * 64, "X15");   // 81 operand rows per image (prep_g_kernel)
  s.z0 = a.takeT<float>(M0 * 256, "z0");
  s.A0 = take_pair(a, parity_elems(d.B, 80, d.T, 128), "A0", c8, rec);
  s.z1 = a.takeT<float>(M1 * 512, "z1");
  s.st1 = take_stat(a, d.B * 512);
  s.A1 = take_pair(a, parity_elems(d.B, 40, d.W1, 256), "A1", c8, rec);
  s.z2 = a.takeT<float>(M2 * 512, "z2");
  s.st2 = take_stat(a, d.B * 512);
  s.A2 = take_pair(a, M2 * 256, "A2");   // feeds the 2D->1D flatten layer (split-bf16, like the whole 1-D trunk)
  s.z3 = a.takeT<float>(L * 256, "z3");
  s.st3 = take_stat(a, d.B * 256);
  for (int i = 0; i < 7; ++i) {
    s.Rf[i] = a.takeT<float>(L * 256, i == 0 ? "R0" : (i == 6 ? "R6" : nullptr));
    s.R[i] = take_pair(a, L * 256, nullptr);
  }
  for (int i = 0; i < 6; ++i) {
    s.z4[i] = a.takeT<float>(L * 1024, i == 0 ? "z4_0" : nullptr);
    s.st4[i] = take_stat(a, d.B * 1024);
    s.H[i] = take_pair(a, L * 512, nullptr);
    s.z5[i] = a.takeT<float>(L * 256, i == 0 ? "z5_0" : nullptr);
    s.st5[i] = take_stat(a, d.B * 256);
  }
  s.z6 = a.takeT<float>(M2 * 256, "z6");
  s.st6 = take_stat(a, (long long)d.B * 20 * 256);
  s.U0 = take_pair(a, M2 * 256, "U0", c8, rec);
  s.z7 = a.takeT<float>(M2 * 1024, "z7");
  s.st7 = take_stat(a, d.B * 256);
  s.U1 = take_pair(a, M7 * 256, "U1", c8, rec);
  s.z8 = a.takeT<float>(M7 * 512, "z8");
  s.st8 = take_stat(a, d.B * 128);
  s.U2 = take_pair(a, M8 * 128, "U2");
  s.total = align_up(a.off, 256);
  return s;
}

ActBuf abuf(BfPair p, float* f32, int nImg, int Y, int X, int C, int parity) {
  return ActBuf{p.hi, p.lo, f32, nImg, Y, X, C, parity, plane_fmt(p)};
}
ActBuf gbuf(float* f32, int nImg, int Y, int X, int C, int parity) {
  return ActBuf{nullptr, nullptr, f32, nImg, Y, X, C, parity, PlaneFmt{0, 1.f, 1.f, 0}};
}

}  // namespace

long long generator_saved_bytes(int B, int T) { return plan_gen_saved(GenDims(B, T), nullptr, nullptr).total; }
std::vector<SavedEntry> generator_saved_layout(int B, int T) {
  std::vector<SavedEntry> v;
  plan_gen_saved(GenDims(B, T), nullptr, &v);
  return v;
}
long long generator_fwd_ws_bytes(int B, int T) {
  GenDims d(B, T);
  return align_up((long long)B * 80 * d.X2 * 128 * 4, 256) + align_up(gen_stat_pool_floats(B) * 4, 256) + 1024 +
         align_up(kTailScratchFloats * 4, 256) + 256;
}

int generator_forward(const void* packed, const float* x, const float* mask, int B, int T,
                      float* out, void* saved, void* ws, const RunCfg& rc) {
  const ModelDesc& md = generator_desc();
  const GenDims d(B, T);
  GenSaved s = plan_gen_saved(d, saved, nullptr, rc.c8, act_rec_of(packed, md));
  Weights W(packed, md, rc.c8);
  Run r{rc};
  cudaStream_t st = rc.stream;
  const auto& cv = md.convs;
  const auto& nm = md.norms;
  Arena wa(ws);
  StatPool sp(wa.takeT<float>(gen_stat_pool_floats(B)));   // conv-epilogue statistics, one memset for all layers
  r.check(cudaMemsetAsync(sp.base, 0, (size_t)gen_stat_pool_floats(B) * sizeof(float), st), "G zero stats");
  r.tailScratch = wa.takeT<float>(kTailScratchFloats);
  float* ssq = nullptr;

  // parity-split buffers have a padding column/row when the extent is odd: keep it zero
  if (d.T & 1) {
    r.check(launch_fill_zero(s.A0.hi, parity_elems(B, 80, d.T, 128) * 2, st), "zero A0");
    r.check(launch_fill_zero(s.A0.lo, parity_elems(B, 80, d.T, 128) * 2, st), "zero A0");
  }
  if (d.W1 & 1) {
    r.check(launch_fill_zero(s.A1.hi, parity_elems(B, 40, d.W1, 256) * 2, st), "zero A1");
    r.check(launch_fill_zero(s.A1.lo, parity_elems(B, 40, d.W1, 256) * 2, st), "zero A1");
  }

  // stem: stack(x*mask, mask) -> 5x15 conv || gates -> a * sigmoid(g)            model.py:241-242
  r.check(launch_prep_g(x, mask, B, T, s.X15.hi, s.X15.lo, st), "prep_g");
  run_conv(r, plain_op(s.X15, B, 81, T, 64), W.fwd(cv[G_STEM]), taps_stem_g(1),
           B, 80, T, plain_out(s.z0, 80, T, 256), W.bias(cv[G_STEM]), nullptr, "G stem conv", 150.0 / 192.0);
  if (r.ok) r.check(launch_apply_fwd(mk_apply(kGatedNoNorm, s.z0, 256, 80, T, Stat{nullptr, nullptr}, 0,
                                              nullptr, nullptr, 1, nullptr,
                                              abuf(s.A0, nullptr, B, 80, T, 128, 1)), st), "G stem glu");
  // downSample1 / downSample2: 5x5 stride 2 conv || gates, IN, gated GLU         model.py:245-246
  run_conv_in(r, parity_op(s.A0, B, 80, T, 128), W.fwd(cv[G_DS1]), taps_s2_fwd(5, 2), B, 40,
              d.W1, plain_out(s.z1, 40, d.W1, 512), W.bias(cv[G_DS1]), "G ds1 conv", sp.take(B, 512), ssq, B, 512, 1,
              40 * d.W1, s.st1, s.z1, (long long)B * 40 * d.W1 * 512);
  if (r.ok) r.check(launch_apply_fwd(mk_apply(kGatedIN, s.z1, 512, 40, d.W1, s.st1, 512, W.gamma(nm[GN_DS1]),
                                              W.beta(nm[GN_DS1]), 1, nullptr,
                                              abuf(s.A1, nullptr, B, 40, d.W1, 256, 1)), st), "G ds1 glu");
  run_conv_in(r, parity_op(s.A1, B, 40, d.W1, 256), W.fwd(cv[G_DS2]), taps_s2_fwd(5, 2), B,
              20, d.W2, plain_out(s.z2, 20, d.W2, 512), W.bias(cv[G_DS2]), "G ds2 conv", sp.take(B, 512), ssq, B, 512, 1,
              20 * d.W2, s.st2, s.z2, (long long)B * 20 * d.W2 * 512);
  if (r.ok) r.check(launch_apply_fwd(mk_apply(kGatedIN, s.z2, 512, 20, d.W2, s.st2, 512, W.gamma(nm[GN_DS2]),
                                              W.beta(nm[GN_DS2]), 1, nullptr,
                                              abuf(s.A2, nullptr, B, 20, d.W2, 256, 0)), st), "G ds2 glu");
  // 2D -> 1D: view (c*20+h), Conv1d k1 5120->256, IN1d                           model.py:249-255
  run_conv_in(r, plain_op(s.A2, B, 20, d.W2, 256), W.fwd(cv[G_2DTO1D]), taps_rows(20, 1), B,
              1, d.W2, plain_out(s.z3, 1, d.W2, 256), W.bias(cv[G_2DTO1D]), "G 2dto1d conv", sp.take(B, 256), ssq, B, 256,
              1, d.W2, s.st3, s.z3);
  if (r.ok) r.check(launch_apply_fwd(mk_apply(kINOnly, s.z3, 256, 1, d.W2, s.st3, 256, W.gamma(nm[GN_2DTO1D]),
                                              W.beta(nm[GN_2DTO1D]), 1, nullptr,
                                              abuf(s.R[0], s.Rf[0], B, 1, d.W2, 256, 0)), st), "G 2dto1d IN");
  // six gated 1-D residual blocks                                                model.py:258-263
  const TapList k3 = taps_s1(1, 3, 0, 1, 1);
  bool trunkFused = false;
  if (r.ok && r.rc.backend == 0 && trunk_fwd_supported(B, d.W2)) {
    // ONE launch for the whole chain (trunk_fused.cu) when the saved-blob / packed-blob layouts are
    // uniform across the six blocks (they are: identical blocks allocated in one loop)
    TrunkFwdArgs ta{};
    TrunkFwdMaps tm{};
    ta.B = B; ta.W2 = d.W2; ta.nPass = r.rc.nPass;
    ta.BX = 4;
    while (ta.BX < d.W2) ta.BX *= 2;
    ta.BB = 128 / ta.BX;
    bool uniform = true;
    auto bytes_between = [](const void* a, const void* b) { return (long long)(reinterpret_cast<const uint8_t*>(b) - reinterpret_cast<const uint8_t*>(a)); };
    tm.Rhi = s.R[0].hi; tm.Rlo = s.R[0].lo; tm.RStrideBytes = bytes_between(s.R[0].hi, s.R[1].hi);
    tm.Hhi = s.H[0].hi; tm.Hlo = s.H[0].lo; tm.HStrideBytes = bytes_between(s.H[0].hi, s.H[1].hi);
    const ConvDesc& a0 = cv[G_RES0];
    const ConvDesc& b0c = cv[G_RES0 + 1];
    tm.Wah = W.bf + a0.fHi; tm.Wal = W.bf + a0.fLo; tm.WaStrideBytes = (cv[G_RES0 + 2].fHi - a0.fHi) * 2;
    tm.Wbh = W.bf + b0c.fHi; tm.Wbl = W.bf + b0c.fLo; tm.WbStrideBytes = (cv[G_RES0 + 3].fHi - b0c.fHi) * 2;
    for (int i = 0; i < 7; ++i) {
      ta.Rf[i] = s.Rf[i]; ta.Rhi[i] = s.R[i].hi; ta.Rlo[i] = s.R[i].lo;
      uniform = uniform && bytes_between(s.R[0].hi, s.R[i].hi) == i * tm.RStrideBytes &&
                bytes_between(s.R[0].lo, s.R[i].lo) == i * tm.RStrideBytes;
    }
    for (int i = 0; i < 6; ++i) {
      const ConvDesc& ca = cv[G_RES0 + 2 * i];
      const ConvDesc& cb = cv[G_RES0 + 2 * i + 1];
      const NormDesc& na = nm[GN_RES0 + 2 * i];
      const NormDesc& nb = nm[GN_RES0 + 2 * i + 1];
      ta.Hhi[i] = s.H[i].hi; ta.Hlo[i] = s.H[i].lo;
      ta.z4[i] = s.z4[i]; ta.z5[i] = s.z5[i];
      ta.mean4[i] = s.st4[i].mean; ta.rstd4[i] = s.st4[i].rstd;
      ta.mean5[i] = s.st5[i].mean; ta.rstd5[i] = s.st5[i].rstd;
      ta.biasA[i] = W.bias(ca); ta.gammaA[i] = W.gamma(na); ta.betaA[i] = W.beta(na);
      ta.biasB[i] = W.bias(cb); ta.gammaB[i] = W.gamma(nb); ta.betaB[i] = W.beta(nb);
      uniform = uniform && bytes_between(s.H[0].hi, s.H[i].hi) == i * tm.HStrideBytes &&
                bytes_between(s.H[0].lo, s.H[i].lo) == i * tm.HStrideBytes &&
                (ca.fHi - a0.fHi) * 2 == i * tm.WaStrideBytes && (ca.fLo - a0.fLo) * 2 == i * tm.WaStrideBytes &&
                (cb.fHi - b0c.fHi) * 2 == i * tm.WbStrideBytes && (cb.fLo - b0c.fLo) * 2 == i * tm.WbStrideBytes;
    }
    if (uniform) {
      // the statistics-pool slices of the layer-by-layer path stay reserved (same pool layout either way)
      for (int i = 0; i < 6; ++i) { sp.take(B, 1024); sp.take(B, 256); }
      r.check(launch_trunk_fwd(ta, tm, st), "G fused trunk");
      trunkFused = true;
    }
  }
  for (int i = 0; i < 6 && !trunkFused; ++i) {
    const ConvDesc& ca = cv[G_RES0 + 2 * i];
    const ConvDesc& cb = cv[G_RES0 + 2 * i + 1];
    const NormDesc& na = nm[GN_RES0 + 2 * i];
    const NormDesc& nb = nm[GN_RES0 + 2 * i + 1];
    run_conv_in(r, plain_op(s.R[i], B, 1, d.W2, 256), W.fwd(ca), k3, B, 1, d.W2,
                plain_out(s.z4[i], 1, d.W2, 1024), W.bias(ca), "G res conv a", sp.take(B, 1024), ssq, B, 1024, 1, d.W2,
                s.st4[i], s.z4[i]);
    if (r.ok) r.check(launch_apply_fwd(mk_apply(kGatedIN, s.z4[i], 1024, 1, d.W2, s.st4[i], 1024, W.gamma(na),
                                                W.beta(na), 1, nullptr,
                                                abuf(s.H[i], nullptr, B, 1, d.W2, 512, 0)), st), "G res glu");
    run_conv_in(r, plain_op(s.H[i], B, 1, d.W2, 512), W.fwd(cb), k3, B, 1, d.W2,
                plain_out(s.z5[i], 1, d.W2, 256), W.bias(cb), "G res conv b", sp.take(B, 256), ssq, B, 256, 1, d.W2,
                s.st5[i], s.z5[i]);
    if (r.ok) r.check(launch_apply_fwd(mk_apply(kINOnly, s.z5[i], 256, 1, d.W2, s.st5[i], 256, W.gamma(nb),
                                                W.beta(nb), 1, s.Rf[i],
                                                abuf(s.R[i + 1], s.Rf[i + 1], B, 1, d.W2, 256, 0)), st), "G res add");
  }
  // 1D -> 2D: Conv1d k1 256->5120, IN1d over time per (c,h) row, view (256,20,W)   model.py:266-271
  {
    OutAddr o{s.z6, (long long)20 * d.W2 * 256, 0, 256, 256, (long long)d.W2 * 256};
    run_conv_in(r, plain_op(s.R[6], B, 1, d.W2, 256), W.fwd(cv[G_1DTO2D]), taps_one(), B, 1,
                d.W2, o, W.bias(cv[G_1DTO2D]), "G 1dto2d conv", sp.take(B * 20, 256), ssq, B * 20, 256, 1, d.W2, s.st6, s.z6);
  }
  if (r.ok) r.check(launch_apply_fwd(mk_apply(kINOnly, s.z6, 256, 1, d.W2, s.st6, 256, W.gamma(nm[GN_1DTO2D]),
                                              W.beta(nm[GN_1DTO2D]), 20, nullptr,
                                              abuf(s.U0, nullptr, B * 20, 1, d.W2, 256, 0)), st), "G 1dto2d IN");
  // upSample1 / upSample2: 5x5 conv, PixelShuffle(2), IN, swish                  model.py:274-275
  const TapList k55 = taps_s1(5, 5, 2, 2, 1);
  run_conv_in(r, plain_op(s.U0, B, 20, d.W2, 256), W.fwd(cv[G_UP1]), k55, B, 20, d.W2,
              plain_out(s.z7, 20, d.W2, 1024), W.bias(cv[G_UP1]), "G up1 conv", sp.take(B, 1024), ssq, B, 1024, 4,
              20 * d.W2, s.st7, s.z7, (long long)B * 20 * d.W2 * 1024);
  if (r.ok) r.check(launch_apply_fwd(mk_apply(kINSwishShuffle, s.z7, 1024, 20, d.W2, s.st7, 256, W.gamma(nm[GN_UP1]),
                                              W.beta(nm[GN_UP1]), 1, nullptr,
                                              abuf(s.U1, nullptr, B, 40, d.X1, 256, 0)), st), "G up1 act");
  run_conv_in(r, plain_op(s.U1, B, 40, d.X1, 256), W.fwd(cv[G_UP2]), k55, B, 40, d.X1,
              plain_out(s.z8, 40, d.X1, 512), W.bias(cv[G_UP2]), "G up2 conv", sp.take(B, 512), ssq, B, 512, 4,
              40 * d.X1, s.st8, s.z8, (long long)B * 40 * d.X1 * 512);
  if (r.ok) r.check(launch_apply_fwd(mk_apply(kINSwishShuffle, s.z8, 512, 40, d.X1, s.st8, 128, W.gamma(nm[GN_UP2]),
                                              W.beta(nm[GN_UP2]), 1, nullptr,
                                              abuf(s.U2, nullptr, B, 80, d.X2, 128, 0)), st), "G up2 act");
  // head: 5x15 conv 128->1 as per-tap GEMM + shifted sum                          model.py:278-279
  float* P = wa.takeT<float>((long long)B * 80 * d.X2 * 128);
  run_conv(r, plain_op(s.U2, B, 80, d.X2, 128), W.fwd(cv[G_HEAD]), taps_one(), B, 80, d.X2,
           plain_out(P, 80, d.X2, 128), nullptr, nullptr, "G head gemm", 75.0 / 128.0);
  if (r.ok) r.check(launch_head_g_fwd(P, W.bias(cv[G_HEAD]), B, 80, d.X2, out, st), "G head sum");
  return r.ok ? 0 : 1;
}

long long generator_bwd_ws_bytes(int B, int T) {
  GenDims d(B, T);
  const long long M0 = (long long)B * 80 * d.T, M2 = (long long)B * 20 * d.W2, L = (long long)B * d.W2,
                  M7 = (long long)B * 40 * d.X1, M8 = (long long)B * 80 * d.X2;
  long long b = 0;
  auto add = [&](long long bytes) { b = align_up(b, 256) + bytes; };
  add(M8 * 128 * 2); add(M8 * 128 * 2);          // dP
  add(M8 * 128 * 4);                             // dU2
  add(M7 * 512 * 2); add(M7 * 512 * 2);          // dz8
  add(M7 * 256 * 4);                             // dU1
  add(M2 * 1024 * 2); add(M2 * 1024 * 2);        // dz7
  add(M2 * 256 * 4);                             // dU0
  add(M2 * 256 * 2); add(M2 * 256 * 2);          // dz6
  for (int i = 0; i < 7; ++i) add(L * 256 * 4);  // dR
  for (int i = 0; i < 6; ++i) { add(L * 256 * 2); add(L * 256 * 2); add(L * 1024 * 2); add(L * 1024 * 2); }  // dz5, dz4 per block
  add(L * 512 * 4);                              // dH
  add(L * 256 * 2); add(L * 256 * 2);            // dz3
  add(M2 * 256 * 4);                             // dA2
  add(M2 * 512 * 2); add(M2 * 512 * 2);          // dz2
  add(parity_elems(B, 40, d.W1, 256) * 4);       // dA1
  add((long long)B * 40 * d.W1 * 512 * 2); add((long long)B * 40 * d.W1 * 512 * 2);  // dz1
  add(parity_elems(B, 80, d.T, 128) * 4);        // dA0
  add(M0 * 256 * 2); add(M0 * 256 * 2);          // dz0
  add((long long)B * 81 * d.T * 64 * 4);         // dX15 (81 operand rows per image)
  add(gen_stat_pool_floats(B) * 4);              // t1 / t2 reduction pool
  add(64 * 4);                                   // dz scale records (C8 mode)
  add(2 * 5120 * 4);                             // affine-grad sink
  add(kTailScratchFloats * 4);                   // conv tail-split partials
  return align_up(b, 256) + 4096;
}

int generator_backward(const void* packed, const void* saved, const float* mask, const float* dout,
                       int B, int T, float* dx, float* gblob, int needWgrad, void* ws,
                       const RunCfg& rc) {
  const ModelDesc& md = generator_desc();
  const GenDims d(B, T);
  GenSaved s = plan_gen_saved(d, const_cast<void*>(saved), nullptr, rc.c8, act_rec_of(packed, md));
  Weights W(packed, md, rc.c8);
  Run r{rc};
  cudaStream_t st = rc.stream;
  const auto& cv = md.convs;
  const auto& nm = md.norms;
  Arena a(ws);
  const long long M0 = (long long)B * 80 * d.T, M2 = (long long)B * 20 * d.W2, L = (long long)B * d.W2,
                  M7 = (long long)B * 40 * d.X1, M8 = (long long)B * 80 * d.X2;
  const long long M1 = (long long)B * 40 * d.W1;
  StatPool tp(a.takeT<float>(gen_stat_pool_floats(B)));   // IN-backward reductions, one memset for all layers
  r.check(cudaMemsetAsync(tp.base, 0, (size_t)gen_stat_pool_floats(B) * sizeof(float), st), "G zero bwd sums");
  float* junk = a.takeT<float>(2 * 5120);  // sink for affine grads when the caller wants none
  r.tailScratch = a.takeT<float>(kTailScratchFloats);
  auto gW = [&](int ci) { return gblob + cv[ci].gW; };
  auto gB = [&](int ci) -> float* { return needWgrad ? gblob + cv[ci].gB : nullptr; };
  auto gGa = [&](int ni) -> float* { return needWgrad ? gblob + nm[ni].gGamma : junk; };
  auto gBe = [&](int ni) -> float* { return needWgrad ? gblob + nm[ni].gBeta : junk + 5120; };
  // C8 mode: the dz of every normalised layer is written as C8 planes with its own scale record
  float* dzRecs = a.takeT<float>(2 * 32);
  int nRec = 0;
  const int dzFmt = rc.c8 ? (rc.half16 ? 2 : 1) : 0;   // C8H: dz keeps only its fp16 plane
  auto dz_pair = [&](long long elems) { return take_pair(a, elems, nullptr, dzFmt, rc.c8 ? dzRecs + 2 * nRec++ : nullptr); };
  const TapList one = taps_one();

  // ---- head                                                                    model.py:278
  BfPair dP = take_pair(a, M8 * 128, nullptr);
  r.check(launch_head_g_bwd(dout, B, 80, d.X2, dP.hi, dP.lo, gB(G_HEAD), st), "G head bwd");
  float* dU2 = a.takeT<float>(M8 * 128);
  run_conv(r, plain_op(dP, B, 80, d.X2, 128), W.bwd(cv[G_HEAD]), one, B, 80, d.X2,
           plain_out(dU2, 80, d.X2, 128), nullptr, nullptr, "G head dgrad", 75.0 / 128.0, nullptr, nullptr, M8 * 128);
  if (needWgrad)
    run_wgrad(r, plain_op(dP, B, 80, d.X2, 128), plain_op(s.U2, B, 80, d.X2, 128),
              one, nullptr, B, 80, d.X2, gW(G_HEAD), "G head wgrad", 75.0 / 128.0);
  // ---- upSample2
  const TapList k55f = taps_s1(5, 5, 2, 2, 1), k55b = taps_s1(5, 5, 2, 2, -1);
  BfPair dz8 = dz_pair(M7 * 512);
  run_bwd(r, mk_bwd(kINSwishShuffle, s.z8, 512, 40, d.X1, s.st8, 128, W.gamma(nm[GN_UP2]), W.beta(nm[GN_UP2]),
                    1, gbuf(dU2, B, 80, d.X2, 128, 0), tp, gGa(GN_UP2), gBe(GN_UP2), dz8, gB(G_UP2)),
          "G up2 bwd");
  float* dU1 = a.takeT<float>(M7 * 256);
  run_conv(r, plain_op(dz8, B, 40, d.X1, 512), W.bwd(cv[G_UP2]), k55b, B, 40, d.X1,
           plain_out(dU1, 40, d.X1, 256), nullptr, nullptr, "G up2 dgrad", 1.0, nullptr, nullptr, M7 * 256);
  if (needWgrad)
    run_wgrad(r, plain_op(dz8, B, 40, d.X1, 512), plain_op(s.U1, B, 40, d.X1, 256),
              k55f, nullptr, B, 40, d.X1, gW(G_UP2), "G up2 wgrad");
  // ---- upSample1
  BfPair dz7 = dz_pair(M2 * 1024);
  run_bwd(r, mk_bwd(kINSwishShuffle, s.z7, 1024, 20, d.W2, s.st7, 256, W.gamma(nm[GN_UP1]), W.beta(nm[GN_UP1]),
                    1, gbuf(dU1, B, 40, d.X1, 256, 0), tp, gGa(GN_UP1), gBe(GN_UP1), dz7, gB(G_UP1)),
          "G up1 bwd");
  float* dU0 = a.takeT<float>(M2 * 256);
  run_conv(r, plain_op(dz7, B, 20, d.W2, 1024), W.bwd(cv[G_UP1]), k55b, B, 20, d.W2,
           plain_out(dU0, 20, d.W2, 256), nullptr, nullptr, "G up1 dgrad", 1.0, nullptr, nullptr, M2 * 256);
  if (needWgrad)
    run_wgrad(r, plain_op(dz7, B, 20, d.W2, 1024), plain_op(s.U0, B, 20, d.W2, 256),
              k55f, nullptr, B, 20, d.W2, gW(G_UP1), "G up1 wgrad");
  // ---- 1D -> 2D
  BfPair dz6 = take_pair(a, M2 * 256, nullptr);
  run_bwd(r, mk_bwd(kINOnly, s.z6, 256, 1, d.W2, s.st6, 256, W.gamma(nm[GN_1DTO2D]), W.beta(nm[GN_1DTO2D]), 20,
                    gbuf(dU0, B * 20, 1, d.W2, 256, 0), tp, gGa(GN_1DTO2D), gBe(GN_1DTO2D), dz6, nullptr),
          "G 1dto2d bwd");
  float* dR[7];
  for (int i = 0; i < 7; ++i) dR[i] = a.takeT<float>(L * 256);
  const TapList rows20 = taps_rows(20, 1);
  {
    // dR6[(b,w), cin] = sum_h sum_c dz6[b,h,w,c] * W[c*20+h][cin]: 20 row taps over the (c,w,h) view
    const ConvDesc& c = cv[G_1DTO2D];
    const WgtOperand wop = W.view(c, true, 256, c.Cd, 20);
    run_conv(r, plain_op(dz6, B, 20, d.W2, 256), wop, rows20, B, 1, d.W2,
             plain_out(dR[6], 1, d.W2, 256), nullptr, nullptr, "G 1dto2d dgrad");
    if (needWgrad) {
      TapList xt;  // x operand (R6) is not shifted; dz operand walks the 20 rows; slice = row
      for (int h = 0; h < 20; ++h) xt.add(0, 0, 0, h);
      run_wgrad(r, plain_op(dz6, B, 20, d.W2, 256), plain_op(s.R[6], B, 1, d.W2, 256),
                xt, &rows20, B, 1, d.W2, gW(G_1DTO2D), "G 1dto2d wgrad");
    }
  }
  // ---- residual blocks, reversed
  const TapList k3f = taps_s1(1, 3, 0, 1, 1), k3b = taps_s1(1, 3, 0, 1, -1);
  float* dH = a.takeT<float>(L * 512);
  // one dz pair per block: the weight-gradient GEMMs read them from the side stream while the main
  // stream already works on the next block
  BfPair dz5[6], dz4[6];
  for (int i = 0; i < 6; ++i) {
    dz5[i] = take_pair(a, L * 256, nullptr);
    dz4[i] = take_pair(a, L * 1024, nullptr);
  }
  bool trunkFused = false;
  if (r.ok && r.rc.backend == 0 && trunk_bwd_supported(B, d.W2)) {
    // ONE launch for the data-gradient chain of all six blocks (trunk_fused.cu)
    TrunkBwdArgs ta{};
    TrunkBwdMaps tm{};
    ta.B = B; ta.W2 = d.W2; ta.nPass = r.rc.nPass;
    ta.BX = 4;
    while (ta.BX < d.W2) ta.BX *= 2;
    ta.BB = 128 / ta.BX;
    ta.dR6 = dR[6]; ta.dR0 = dR[0];
    auto bytes_between = [](const void* p0, const void* p1) { return (long long)(reinterpret_cast<const uint8_t*>(p1) - reinterpret_cast<const uint8_t*>(p0)); };
    const ConvDesc& a0 = cv[G_RES0];
    const ConvDesc& b0c = cv[G_RES0 + 1];
    tm.Z5hi = dz5[0].hi; tm.Z5lo = dz5[0].lo; tm.Z5StrideBytes = bytes_between(dz5[0].hi, dz5[1].hi);
    tm.Z4hi = dz4[0].hi; tm.Z4lo = dz4[0].lo; tm.Z4StrideBytes = bytes_between(dz4[0].hi, dz4[1].hi);
    tm.Wah = W.bf + a0.dHi; tm.Wal = W.bf + a0.dLo; tm.WaStrideBytes = (cv[G_RES0 + 2].dHi - a0.dHi) * 2;
    tm.Wbh = W.bf + b0c.dHi; tm.Wbl = W.bf + b0c.dLo; tm.WbStrideBytes = (cv[G_RES0 + 3].dHi - b0c.dHi) * 2;
    bool uniform = true;
    for (int i = 0; i < 6; ++i) {
      const ConvDesc& ca = cv[G_RES0 + 2 * i];
      const ConvDesc& cb = cv[G_RES0 + 2 * i + 1];
      const int na = GN_RES0 + 2 * i, nb = GN_RES0 + 2 * i + 1;
      ta.z4[i] = s.z4[i]; ta.z5[i] = s.z5[i];
      ta.mean4[i] = s.st4[i].mean; ta.rstd4[i] = s.st4[i].rstd;
      ta.mean5[i] = s.st5[i].mean; ta.rstd5[i] = s.st5[i].rstd;
      ta.gammaA[i] = W.gamma(nm[na]); ta.betaA[i] = W.beta(nm[na]); ta.gammaB[i] = W.gamma(nm[nb]);
      ta.dgammaA[i] = gGa(na); ta.dbetaA[i] = gBe(na); ta.dgammaB[i] = gGa(nb); ta.dbetaB[i] = gBe(nb);
      ta.dz5hi[i] = dz5[i].hi; ta.dz5lo[i] = dz5[i].lo; ta.dz4hi[i] = dz4[i].hi; ta.dz4lo[i] = dz4[i].lo;
      uniform = uniform && bytes_between(dz5[0].hi, dz5[i].hi) == i * tm.Z5StrideBytes &&
                bytes_between(dz5[0].lo, dz5[i].lo) == i * tm.Z5StrideBytes &&
                bytes_between(dz4[0].hi, dz4[i].hi) == i * tm.Z4StrideBytes &&
                bytes_between(dz4[0].lo, dz4[i].lo) == i * tm.Z4StrideBytes &&
                (ca.dHi - a0.dHi) * 2 == i * tm.WaStrideBytes && (ca.dLo - a0.dLo) * 2 == i * tm.WaStrideBytes &&
                (cb.dHi - b0c.dHi) * 2 == i * tm.WbStrideBytes && (cb.dLo - b0c.dLo) * 2 == i * tm.WbStrideBytes;
    }
    if (uniform) {
      for (int i = 0; i < 6; ++i) { tp.take(B, 256); tp.take(B, 1024); }   // keep the reduction-pool layout of the layer path
      r.check(launch_trunk_bwd(ta, tm, st), "G fused trunk bwd");
      trunkFused = true;
      if (needWgrad)
        for (int i = 5; i >= 0; --i) {
          const ConvDesc& ca = cv[G_RES0 + 2 * i];
          const ConvDesc& cb = cv[G_RES0 + 2 * i + 1];
          run_wgrad(r, plain_op(dz5[i], B, 1, d.W2, 256), plain_op(s.H[i], B, 1, d.W2, 512),
                    k3f, nullptr, B, 1, d.W2, gblob + cb.gW, "G res wgrad b");
          run_wgrad(r, plain_op(dz4[i], B, 1, d.W2, 1024), plain_op(s.R[i], B, 1, d.W2, 256),
                    k3f, nullptr, B, 1, d.W2, gblob + ca.gW, "G res wgrad a");
        }
    }
  }
  for (int i = 5; i >= 0 && !trunkFused; --i) {
    const ConvDesc& ca = cv[G_RES0 + 2 * i];
    const ConvDesc& cb = cv[G_RES0 + 2 * i + 1];
    const int na = GN_RES0 + 2 * i, nb = GN_RES0 + 2 * i + 1;
    run_bwd(r, mk_bwd(kINOnly, s.z5[i], 256, 1, d.W2, s.st5[i], 256, W.gamma(nm[nb]), W.beta(nm[nb]), 1,
                      gbuf(dR[i + 1], B, 1, d.W2, 256, 0), tp, gGa(nb), gBe(nb), dz5[i], nullptr), "G res bwd b");
    run_conv(r, plain_op(dz5[i], B, 1, d.W2, 256), W.bwd(cb), k3b, B, 1, d.W2,
             plain_out(dH, 1, d.W2, 512), nullptr, nullptr, "G res dgrad b");
    if (needWgrad)
      run_wgrad(r, plain_op(dz5[i], B, 1, d.W2, 256), plain_op(s.H[i], B, 1, d.W2, 512),
                k3f, nullptr, B, 1, d.W2, gblob + cb.gW, "G res wgrad b");
    run_bwd(r, mk_bwd(kGatedIN, s.z4[i], 1024, 1, d.W2, s.st4[i], 1024, W.gamma(nm[na]), W.beta(nm[na]), 1,
                      gbuf(dH, B, 1, d.W2, 512, 0), tp, gGa(na), gBe(na), dz4[i], nullptr), "G res bwd a");
    // dR[i] = dR[i+1] (skip connection) + dgrad
    run_conv(r, plain_op(dz4[i], B, 1, d.W2, 1024), W.bwd(ca), k3b, B, 1, d.W2,
             plain_out(dR[i], 1, d.W2, 256), nullptr, dR[i + 1], "G res dgrad a");
    if (needWgrad)
      run_wgrad(r, plain_op(dz4[i], B, 1, d.W2, 1024), plain_op(s.R[i], B, 1, d.W2, 256),
                k3f, nullptr, B, 1, d.W2, gblob + ca.gW, "G res wgrad a");
  }
  // ---- 2D -> 1D
  BfPair dz3 = take_pair(a, L * 256, nullptr);
  run_bwd(r, mk_bwd(kINOnly, s.z3, 256, 1, d.W2, s.st3, 256, W.gamma(nm[GN_2DTO1D]), W.beta(nm[GN_2DTO1D]), 1,
                    gbuf(dR[0], B, 1, d.W2, 256, 0), tp, gGa(GN_2DTO1D), gBe(GN_2DTO1D), dz3, nullptr),
          "G 2dto1d bwd");
  float* dA2 = a.takeT<float>(M2 * 256);
  {
    // dA2[b,h,w,c] = sum_n dz3[(b,w), n] * W[n][c*20+h]: one tap, N = (h,c) = 5120 split back over h
    const ConvDesc& c = cv[G_2DTO1D];
    const WgtOperand wop = W.view(c, true, 256, 5120, 1);
    OutAddr o{dA2, (long long)20 * d.W2 * 256, 0, 256, 256, (long long)d.W2 * 256};
    run_conv(r, plain_op(dz3, B, 1, d.W2, 256), wop, one, B, 1, d.W2, o, nullptr, nullptr,
             "G 2dto1d dgrad");
    if (needWgrad)
      run_wgrad(r, plain_op(dz3, B, 1, d.W2, 256), plain_op(s.A2, B, 20, d.W2, 256),
                rows20, nullptr, B, 1, d.W2, gW(G_2DTO1D), "G 2dto1d wgrad");
  }
  // ---- downSample2
  BfPair dz2 = dz_pair(M2 * 512);
  run_bwd(r, mk_bwd(kGatedIN, s.z2, 512, 20, d.W2, s.st2, 512, W.gamma(nm[GN_DS2]), W.beta(nm[GN_DS2]), 1,
                    gbuf(dA2, B, 20, d.W2, 256, 0), tp, gGa(GN_DS2), gBe(GN_DS2), dz2, nullptr), "G ds2 bwd");
  float* dA1 = a.takeT<float>(parity_elems(B, 40, d.W1, 256));
  run_dgrad_s2(r, plain_op(dz2, B, 20, d.W2, 512), W.bwd(cv[G_DS2]), 5, 2, B, 20, (d.W1 + 1) / 2, 256,
               dA1, "G ds2 dgrad");
  if (needWgrad)
    run_wgrad(r, plain_op(dz2, B, 20, d.W2, 512), parity_op(s.A1, B, 40, d.W1, 256),
              taps_s2_fwd(5, 2), nullptr, B, 20, d.W2, gW(G_DS2), "G ds2 wgrad");
  // ---- downSample1
  BfPair dz1 = dz_pair(M1 * 512);
  run_bwd(r, mk_bwd(kGatedIN, s.z1, 512, 40, d.W1, s.st1, 512, W.gamma(nm[GN_DS1]), W.beta(nm[GN_DS1]), 1,
                    gbuf(dA1, B, 40, d.W1, 256, 1), tp, gGa(GN_DS1), gBe(GN_DS1), dz1, nullptr), "G ds1 bwd");
  float* dA0 = a.takeT<float>(parity_elems(B, 80, d.T, 128));
  run_dgrad_s2(r, plain_op(dz1, B, 40, d.W1, 512), W.bwd(cv[G_DS1]), 5, 2, B, 40, (d.T + 1) / 2, 128,
               dA0, "G ds1 dgrad");
  if (needWgrad)
    run_wgrad(r, plain_op(dz1, B, 40, d.W1, 512), parity_op(s.A0, B, 80, d.T, 128),
              taps_s2_fwd(5, 2), nullptr, B, 40, d.W1, gW(G_DS1), "G ds1 wgrad");
  // ---- stem
  BfPair dz0 = take_pair(a, M0 * 256, nullptr);
  run_bwd(r, mk_bwd(kGatedNoNorm, s.z0, 256, 80, d.T, Stat{nullptr, nullptr}, 0, nullptr, nullptr, 1,
                    gbuf(dA0, B, 80, d.T, 128, 1), tp, nullptr, nullptr, dz0, gB(G_STEM)),
          "G stem bwd");
  if (needWgrad)
    run_wgrad(r, plain_op(dz0, B, 80, d.T, 256), plain_op(s.X15, B, 81, d.T, 64),
              taps_stem_g(1), nullptr, B, 80, d.T, gW(G_STEM), "G stem wgrad", 150.0 / 192.0);
  if (dx) {
    float* dX15 = a.takeT<float>((long long)B * 81 * d.T * 64);
    run_conv(r, plain_op(dz0, B, 80, d.T, 256), W.bwd(cv[G_STEM]), taps_stem_g(-1), B, 81,
             d.T, plain_out(dX15, 81, d.T, 64), nullptr, nullptr, "G stem dgrad", 150.0 / 192.0);
    if (r.ok) r.check(launch_col2im_g(dX15, mask, B, d.T, dx, st), "G col2im");
  }
  r.join();
  return r.ok ? 0 : 1;
}

// ================================================================================================
// Discriminator
namespace {

struct DisDims {
  int B, T, W1, W2, W3;
  DisDims(int b, int t) : B(b), T(t) {
    W1 = (T + 1) / 2;
    W2 = (W1 + 1) / 2;
    W3 = (W2 + 1) / 2;
  }
};
struct DisSaved {
  float* xin;     // copy of the input (the stem backward recomputes its pre-activation from it)
  BfPair D0; float* z1; Stat st1;
  BfPair D1; float* z2; Stat st2;
  BfPair D2; float* z3; Stat st3;
  BfPair D3;
  long long total;
};
DisSaved plan_dis_saved(const DisDims& d, void* base, std::vector<SavedEntry>* layout, int c8 = 0,
                        const float* rec = nullptr) {
  Arena a(base);
  a.layout = layout;
  DisSaved s{};
  const long long M0 = (long long)d.B * 80 * d.T, M1 = (long long)d.B * 40 * d.W1,
                  M2 = (long long)d.B * 20 * d.W2, M3 = (long long)d.B * 10 * d.W3;
  s.xin = a.takeT<float>(M0, "xin");
  s.D0 = take_pair(a, parity_elems(d.B, 80, d.T, 128), "D0", c8, rec);
  s.z1 = a.takeT<float>(M1 * 256, "z1");
  s.st1 = take_stat(a, d.B * 256);
  s.D1 = take_pair(a, parity_elems(d.B, 40, d.W1, 256), "D1", c8, rec);
  s.z2 = a.takeT<float>(M2 * 512, "z2");
  s.st2 = take_stat(a, d.B * 512);
  s.D2 = take_pair(a, parity_elems(d.B, 20, d.W2, 512), "D2", c8, rec);
  s.z3 = a.takeT<float>(M3 * 1024, "z3");
  s.st3 = take_stat(a, d.B * 1024);
  s.D3 = take_pair(a, M3 * 1024, "D3");
  s.total = align_up(a.off, 256);
  return s;
}
}  // namespace

long long discriminator_saved_bytes(int B, int T) { return plan_dis_saved(DisDims(B, T), nullptr, nullptr).total; }
std::vector<SavedEntry> discriminator_saved_layout(int B, int T) {
  std::vector<SavedEntry> v;
  plan_dis_saved(DisDims(B, T), nullptr, &v);
  return v;
}
long long discriminator_fwd_ws_bytes(int B, int T) {
  DisDims d(B, T);
  return align_up((long long)B * 10 * d.W3 * 128 * 4, 256) + align_up(dis_stat_pool_floats(B) * 4, 256) + 1024 +
         align_up(kTailScratchFloats * 4, 256) + 256;
}

int discriminator_forward(const void* packed, const float* x, int B, int T, float* out, void* saved,
                          void* ws, const RunCfg& rc) {
  const ModelDesc& md = discriminator_desc();
  const DisDims d(B, T);
  DisSaved s = plan_dis_saved(d, saved, nullptr, rc.c8, act_rec_of(packed, md));
  Weights W(packed, md, rc.c8);
  Run r{rc};
  cudaStream_t st = rc.stream;
  const auto& cv = md.convs;
  const auto& nm = md.norms;
  Arena wa(ws);
  StatPool sp(wa.takeT<float>(dis_stat_pool_floats(B)));
  r.check(cudaMemsetAsync(sp.base, 0, (size_t)dis_stat_pool_floats(B) * sizeof(float), st), "D zero stats");
  r.tailScratch = wa.takeT<float>(kTailScratchFloats);
  float* ssq = nullptr;
  if (d.T & 1) {
    r.check(launch_fill_zero(s.D0.hi, parity_elems(B, 80, d.T, 128) * 2, st), "zero D0");
    r.check(launch_fill_zero(s.D0.lo, parity_elems(B, 80, d.T, 128) * 2, st), "zero D0");
  }
  if (d.W1 & 1) {
    r.check(launch_fill_zero(s.D1.hi, parity_elems(B, 40, d.W1, 256) * 2, st), "zero D1");
    r.check(launch_fill_zero(s.D1.lo, parity_elems(B, 40, d.W1, 256) * 2, st), "zero D1");
  }
  if (d.W2 & 1) {
    r.check(launch_fill_zero(s.D2.hi, parity_elems(B, 20, d.W2, 512) * 2, st), "zero D2");
    r.check(launch_fill_zero(s.D2.lo, parity_elems(B, 20, d.W2, 512) * 2, st), "zero D2");
  }
  // convLayer1: 3x3 conv 1->128 + swish, direct CUDA-core kernel (K = 9)        model.py:290-295,344
  r.check(cudaMemcpyAsync(s.xin, x, (size_t)B * 80 * T * sizeof(float), cudaMemcpyDeviceToDevice, st), "D save x");
  r.check(launch_d_stem_fwd(x, B, T, W.bf + cv[D_STEM].fHi, W.bf + cv[D_STEM].fLo, W.bias(cv[D_STEM]),
                            abuf(s.D0, nullptr, B, 80, T, 128, 1), st), "D stem");
  // downSample1..3: 3x3 stride 2 conv + IN + swish                                 model.py:345-347
  const TapList k33 = taps_s2_fwd(3, 1);
  run_conv_in(r, parity_op(s.D0, B, 80, T, 128), W.fwd(cv[D_DS1]), k33, B, 40, d.W1,
              plain_out(s.z1, 40, d.W1, 256), W.bias(cv[D_DS1]), "D ds1 conv", sp.take(B, 256), ssq, B, 256, 1, 40 * d.W1,
              s.st1, s.z1, (long long)B * 40 * d.W1 * 256);
  if (r.ok) r.check(launch_apply_fwd(mk_apply(kINSwish, s.z1, 256, 40, d.W1, s.st1, 256, W.gamma(nm[DN_DS1]),
                                              W.beta(nm[DN_DS1]), 1, nullptr, abuf(s.D1, nullptr, B, 40, d.W1, 256, 1)), st), "D ds1 act");
  run_conv_in(r, parity_op(s.D1, B, 40, d.W1, 256), W.fwd(cv[D_DS2]), k33, B, 20, d.W2,
              plain_out(s.z2, 20, d.W2, 512), W.bias(cv[D_DS2]), "D ds2 conv", sp.take(B, 512), ssq, B, 512, 1, 20 * d.W2,
              s.st2, s.z2, (long long)B * 20 * d.W2 * 512);
  if (r.ok) r.check(launch_apply_fwd(mk_apply(kINSwish, s.z2, 512, 20, d.W2, s.st2, 512, W.gamma(nm[DN_DS2]),
                                              W.beta(nm[DN_DS2]), 1, nullptr, abuf(s.D2, nullptr, B, 20, d.W2, 512, 1)), st), "D ds2 act");
  run_conv_in(r, parity_op(s.D2, B, 20, d.W2, 512), W.fwd(cv[D_DS3]), k33, B, 10, d.W3,
              plain_out(s.z3, 10, d.W3, 1024), W.bias(cv[D_DS3]), "D ds3 conv", sp.take(B, 1024), ssq, B, 1024, 1, 10 * d.W3,
              s.st3, s.z3, (long long)B * 10 * d.W3 * 1024);
  if (r.ok) r.check(launch_apply_fwd(mk_apply(kINSwish, s.z3, 1024, 10, d.W3, s.st3, 1024, W.gamma(nm[DN_DS3]),
                                              W.beta(nm[DN_DS3]), 1, nullptr, abuf(s.D3, nullptr, B, 10, d.W3, 1024, 0)), st), "D ds3 act");
  // outputConvLayer 1x3 1024->1 + sigmoid                                          model.py:323-327,348
  float* P = wa.takeT<float>((long long)B * 10 * d.W3 * 128);
  run_conv(r, plain_op(s.D3, B, 10, d.W3, 1024), W.fwd(cv[D_HEAD]), taps_one(), B, 10, d.W3,
           plain_out(P, 10, d.W3, 128), nullptr, nullptr, "D head gemm", 3.0 / 128.0);
  if (r.ok) r.check(launch_head_d_fwd(P, W.bias(cv[D_HEAD]), B, 10, d.W3, out, st), "D head sum");
  return r.ok ? 0 : 1;
}

long long discriminator_bwd_ws_bytes(int B, int T) {
  DisDims d(B, T);
  const long long M0 = (long long)B * 80 * d.T, M1 = (long long)B * 40 * d.W1, M2 = (long long)B * 20 * d.W2,
                  M3 = (long long)B * 10 * d.W3;
  long long b = 0;
  auto add = [&](long long bytes) { b = align_up(b, 256) + bytes; };
  add(dis_stat_pool_floats(B) * 4);                            // t1 / t2 reduction pool
  add(64 * 4);                                                 // dz scale records (C8 mode)
  add(2 * 1024 * 4);                                           // affine-grad sink
  add(kTailScratchFloats * 4);                                 // conv tail-split partials
  add(M3 * 128 * 2); add(M3 * 128 * 2);                        // dP
  add(M3 * 1024 * 4);                                          // dD3
  add(M3 * 1024 * 2); add(M3 * 1024 * 2);                      // dz3
  add(parity_elems(B, 20, d.W2, 512) * 4);                     // dD2
  add(M2 * 512 * 2); add(M2 * 512 * 2);                        // dz2
  add(parity_elems(B, 40, d.W1, 256) * 4);                     // dD1
  add(M1 * 256 * 2); add(M1 * 256 * 2);                        // dz1
  add(parity_elems(B, 80, d.T, 128) * 4);                      // dD0
  add(M0 * 12 * 4);                                            // q (stem input-gradient taps)
  return align_up(b, 256) + 4096;
}

int discriminator_backward(const void* packed, const void* saved, const float* out, const float* dout,
                           int B, int T, float* dx, float* gblob, int needWgrad, void* ws,
                           const RunCfg& rc) {
  const ModelDesc& md = discriminator_desc();
  const DisDims d(B, T);
  DisSaved s = plan_dis_saved(d, const_cast<void*>(saved), nullptr, rc.c8, act_rec_of(packed, md));
  Weights W(packed, md, rc.c8);
  Run r{rc};
  cudaStream_t st = rc.stream;
  const auto& cv = md.convs;
  const auto& nm = md.norms;
  Arena a(ws);
  const long long M0 = (long long)B * 80 * d.T, M1 = (long long)B * 40 * d.W1, M2 = (long long)B * 20 * d.W2,
                  M3 = (long long)B * 10 * d.W3;
  StatPool tp(a.takeT<float>(dis_stat_pool_floats(B)));
  r.check(cudaMemsetAsync(tp.base, 0, (size_t)dis_stat_pool_floats(B) * sizeof(float), st), "D zero bwd sums");
  float* junk = a.takeT<float>(2 * 1024);
  r.tailScratch = a.takeT<float>(kTailScratchFloats);
  auto gW = [&](int ci) { return gblob + cv[ci].gW; };
  auto gB = [&](int ci) -> float* { return needWgrad ? gblob + cv[ci].gB : nullptr; };
  auto gGa = [&](int ni) -> float* { return needWgrad ? gblob + nm[ni].gGamma : junk; };
  auto gBe = [&](int ni) -> float* { return needWgrad ? gblob + nm[ni].gBeta : junk + 1024; };
  float* dzRecs = a.takeT<float>(2 * 32);
  int nRec = 0;
  const TapList one = taps_one();
  const TapList k33 = taps_s2_fwd(3, 1);

  BfPair dP = take_pair(a, M3 * 128, nullptr);
  r.check(launch_head_d_bwd(dout, out, B, 10, d.W3, dP.hi, dP.lo, gB(D_HEAD), st), "D head bwd");
  float* dD3 = a.takeT<float>(M3 * 1024);
  run_conv(r, plain_op(dP, B, 10, d.W3, 128), W.bwd(cv[D_HEAD]), one, B, 10, d.W3,
           plain_out(dD3, 10, d.W3, 1024), nullptr, nullptr, "D head dgrad", 3.0 / 128.0, nullptr, nullptr, M3 * 1024);
  if (needWgrad)
    run_wgrad(r, plain_op(dP, B, 10, d.W3, 128), plain_op(s.D3, B, 10, d.W3, 1024), one,
              nullptr, B, 10, d.W3, gW(D_HEAD), "D head wgrad", 3.0 / 128.0);

  struct Lvl { int ci, ni, Nz, Cin, Yo, Xo, Yi, Xi; const float* z; Stat st; BfPair xin; };
  // ds3: z3 [B,10,W3,1024] <- D2 (20 x W2, 512);  ds2: z2 [B,20,W2,512] <- D1 (40 x W1, 256);
  // ds1: z1 [B,40,W1,256] <- D0 (80 x T, 128)
  Lvl lv[3] = {{D_DS3, DN_DS3, 1024, 512, 10, d.W3, 20, d.W2, s.z3, s.st3, s.D2},
               {D_DS2, DN_DS2, 512, 256, 20, d.W2, 40, d.W1, s.z2, s.st2, s.D1},
               {D_DS1, DN_DS1, 256, 128, 40, d.W1, 80, d.T, s.z1, s.st1, s.D0}};
  float* dAct = dD3;   // gradient w.r.t. the level's activation output
  int dActParity = 0;
  for (int l = 0; l < 3; ++l) {
    const Lvl& v = lv[l];
    const long long Mo = (long long)B * v.Yo * v.Xo;
    BfPair dz = take_pair(a, Mo * v.Nz, nullptr, rc.c8 ? (rc.half16 ? 2 : 1) : 0, rc.c8 ? dzRecs + 2 * nRec++ : nullptr);
    run_bwd(r, mk_bwd(kINSwish, v.z, v.Nz, v.Yo, v.Xo, v.st, v.Nz, W.gamma(nm[v.ni]), W.beta(nm[v.ni]), 1,
                      gbuf(dAct, B, v.Yo, v.Xo, v.Nz, dActParity), tp, gGa(v.ni), gBe(v.ni), dz, nullptr),
            "D ds bwd");
    float* dIn = a.takeT<float>(parity_elems(B, v.Yi, v.Xi, v.Cin));
    run_dgrad_s2(r, plain_op(dz, B, v.Yo, v.Xo, v.Nz), W.bwd(cv[v.ci]), 3, 1, B, (v.Yi + 1) / 2,
                 (v.Xi + 1) / 2, v.Cin, dIn, "D ds dgrad");
    if (needWgrad)
      run_wgrad(r, plain_op(dz, B, v.Yo, v.Xo, v.Nz), parity_op(v.xin, B, v.Yi, v.Xi, v.Cin),
                k33, nullptr, B, v.Yo, v.Xo, gW(v.ci), "D ds wgrad");
    dAct = dIn;
    dActParity = 1;
  }
  // stem: direct kernel (recomputes z from x), weight/bias grads straight into the gradient blob
  {
    float* q = dx ? a.takeT<float>(M0 * 12) : nullptr;
    if (r.ok) r.check(launch_d_stem_bwd(s.xin, B, d.T, W.bf + cv[D_STEM].fHi, W.bf + cv[D_STEM].fLo,
                                        W.bias(cv[D_STEM]), gbuf(dAct, B, 80, d.T, 128, 1),
                                        needWgrad ? gW(D_STEM) : nullptr, gB(D_STEM), q, st), "D stem bwd");
    if (dx && r.ok) r.check(launch_col2im_d(q, B, d.T, dx, st), "D col2im");
  }
  (void)M1; (void)M2;
  r.join();
  return r.ok ? 0 : 1;
}

}  // namespace mcgvc
