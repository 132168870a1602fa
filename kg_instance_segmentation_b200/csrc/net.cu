// Host-side runtime of the KGnet forward (KGnet.py:123-350 of the reference): weight store (BN folding,
// repacking), per-shape execution plan, forward_dec and forward_seg.  All arithmetic runs in the kernels of
// net_kernels.cu (CUDA cores) and tc_conv.cu (tcgen05 tensor cores).
#include "net.cuh"
#include "tc_conv.cuh"
#include "tc_shift.cuh"
#include "tc_stem.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace kg {

// stage ids for kg_timing_collect
enum { ST_STEM = 8, ST_BACKBONE = 9, ST_DECODER = 10, ST_HEAD1 = 11, ST_HEAD2 = 12, ST_RESIZE = 13, ST_POOL = 14, ST_EXPORT = 15,
       ST_SEG = 16 };

struct ConvW {
  std::string name;
  int Cout = 0, Cin = 0, R = 0, S = 0;
  std::vector<float> h_w;   // [tap][cin][cout], BN folded
  std::vector<float> h_b;   // [cout]
  float* d_w = nullptr;
  float* d_b = nullptr;
  TcWeights tc;             // fp16 hi/lo [tap][cout_pad][cin] (tensor-core path)
  mutable TcStemWeights stemw;            // swizzled weight image of the tensor-core stem kernel, built on first use
  mutable TcStemWeights stemw_u8;         // same for the uint8-input variant (w / 255, folded bias)
  mutable TcShiftPacked shift1, shift3;   // weights packed for the row-GEMM + shift-add kernel (1 / 3 passes), built on first use
};

struct Tensor {
  size_t off_hi = 0, off_lo = 0;   // byte offsets into the workspace
  int H = 0, W = 0, C = 0;
};

enum OpType { OP_CONV, OP_BILINEAR, OP_MAXPOOL, OP_EXPORT, OP_HEADS2 };

struct Op {
  OpType type = OP_CONV;
  const ConvW* w = nullptr;
  int in0 = -1, in1 = -1, out = -1, res = -1;   // tensor ids
  int in0_coff = 0, C0 = 0, C1 = 0;
  bool x_input = false, in_single = false, out_single = false;
  int out32_ext = -1;                            // index into the external fp32 NCHW outputs (heads 0..11, feats 12..16)
  int stride = 1, pad = 0;
  bool relu = false, sigmoid = false;
  size_t prob_off = 0;                           // index into the device problem array
  int nprob = 0, max_pix = 0;
  int Hout = 0, Wout = 0;
  int stage = ST_BACKBONE;
  int tc_passes = 0;                             // 0: CUDA-core kernel; 1 or 3: tensor-core kernel passes
  int tc_index = -1;                             // index into Plan::tc_ops (OP_CONV) / Plan::shift_ops (OP_HEADS2)
  bool use_shift = false;                        // OP_CONV through the row-GEMM + shift-add kernel (64-channel 3x3 convs)
  int head_scale = -1;                           // OP_HEADS2: the three second-layer head convs of this scale in one launch
};

struct Plan {
  int N = 0, H = 0, W = 0, precision = -1;
  std::vector<Tensor> tensors;
  std::vector<Op> ops;
  std::vector<ConvProb> h_conv_probs;
  std::vector<ResizeProb> h_resize_probs;
  ConvProb* d_conv_probs = nullptr;
  ResizeProb* d_resize_probs = nullptr;
  std::vector<TcConvOp> tc_ops;
  std::vector<TcShiftOp> shift_ops;
  const void* tc_workspace = nullptr;            // workspace base the tensor maps were encoded for
  size_t bytes = 0;
  int feat_ids[5] = {-1, -1, -1, -1, -1};
  int launches = 0;
  ~Plan() {
    if (d_conv_probs) cudaFree(d_conv_probs);
    if (d_resize_probs) cudaFree(d_resize_probs);
  }
};

struct SegBox { int img; int L; int rect[5][4]; int mask_index; };

// Grouped problems of one forward_seg call (built on the host by seg_prepare, uploaded once per call).
struct SegPlan {
  int N = 0, H = 0, W = 0;
  bool valid = false;
  size_t scratch_halfs = 0;      // elements per scratch plane (hi plane, then lo plane)
  long long mask_floats = 0;
  int n_masks = 0, n_boxes = 0;
  struct Step {                  // level l: pre(level l+1) -> bilinear -> up conv -> cat(patch_l, up) -> 1x1
    std::vector<ResizeProb> rs_raw, rs_scr;
    std::vector<ConvProb> up, cat;
    int pix_raw = 0, pix_scr = 0, pix = 0;
  } steps[4];
  std::vector<ConvProb> h0_raw, h0_scr, h1;
  int pix_h0_raw = 0, pix_h0_scr = 0, pix_h1 = 0;
  // ---- tensor-core path: every level's crops packed into one "atlas" image (1-px zero gaps between boxes) ----
  bool atlas = false;
  struct Level {
    int HA = 0, WA = 0;
    size_t P = 0, U = 0, V = 0, Cc = 0, T = 0;     // element offsets of the atlas tensors inside one scratch plane
    size_t mask = 0;                               // byte offset of the uint8 validity mask
    std::vector<ResizeProb> crop, deepest, up;     // feature crop -> P; P -> C for boxes whose deepest level this is; pre -> U
    std::vector<RectProb> rects;
    int pix_crop = 0, pix_deepest = 0, pix_up = 0;   // largest problem of each list, in rows (up: framed rows)
  } lv[5];
  size_t mask_base = 0;                            // byte offset of the mask region inside the seg workspace
  // problem lists of every launch, packed into ONE blob: built by seg_prepare in pinned host memory, copied by forward_seg
  // (asynchronously, from pinned memory) to `blob_base` of the caller's seg workspace
  size_t blob_base = 0, blob_bytes = 0;
  size_t o_rs_raw[4] = {}, o_rs_scr[4] = {}, o_up4[4] = {}, o_cat[4] = {}, o_h0r = 0, o_h0s = 0, o_h1 = 0;   // CUDA-core path
  size_t o_crop[5] = {}, o_deep[5] = {}, o_up[5] = {}, o_rect[5] = {};                                          // atlas path
};

struct Net {
  int blocks[3] = {3, 4, 6};
  std::map<std::string, ConvW> convs;
  bool finalized = false;
  std::unique_ptr<Plan> plan;
  SegPlan seg;
  char* h_blob = nullptr; size_t h_blob_cap = 0;   // pinned staging of the forward_seg problem lists (owned by seg_prepare)
  cudaEvent_t blob_copied = nullptr;               // recorded after forward_seg's H2D copy of the blob: seg_prepare waits on it before reuse
  ~Net() {
    for (auto& kv : convs) {
      if (kv.second.d_w) cudaFree(kv.second.d_w);
      if (kv.second.d_b) cudaFree(kv.second.d_b);
      tc_free_weights(kv.second.tc);
    }
    if (h_blob) cudaFreeHost(h_blob);
    if (blob_copied) cudaEventDestroy(blob_copied);
  }
};

// Appends `bytes` to the pinned blob of the net (16-byte aligned); grows it (allocation: seg_prepare only).
struct BlobWriter {
  Net* net; size_t size = 0; std::vector<char> tmp;
  size_t put(const void* src, size_t bytes) {
    const size_t o = align_up(tmp.size(), 16);
    tmp.resize(o + bytes);
    if (bytes) memcpy(tmp.data() + o, src, bytes);
    return o;
  }
  int commit(SegPlan& sp) {
    if (net->blob_copied == nullptr) KG_CUDA_CHECK(cudaEventCreateWithFlags(&net->blob_copied, cudaEventDisableTiming));
    KG_CUDA_CHECK(cudaEventSynchronize(net->blob_copied));       // a previous forward_seg may still be reading the staging buffer
    if (net->h_blob_cap < tmp.size()) {
      if (net->h_blob) cudaFreeHost(net->h_blob);
      net->h_blob = nullptr; net->h_blob_cap = 0;
      KG_CUDA_CHECK(cudaMallocHost(&net->h_blob, tmp.size() * 2 + 4096));
      net->h_blob_cap = tmp.size() * 2 + 4096;
    }
    if (!tmp.empty()) memcpy(net->h_blob, tmp.data(), tmp.size());
    sp.blob_bytes = tmp.size();
    return KG_OK;
  }
};

// ------------------------------------------------------------------------------------------------
static const int kPlanes[3] = {64, 128, 256};
static const int kFeatC[5] = {64, 64, 256, 512, 1024};
static const int kHeadC[4] = {64, 64, 256, 512};
static const char* kHeadNames[3] = {"kp_head", "short_offset_head", "mid_offset_head"};
static const int kHeadOut[3] = {5, 10, 40};

static std::vector<std::string> expected_convs(const int* blocks) {
  std::vector<std::string> v = {"conv1", "c0_conv.0", "c0_conv.2", "seg_head.0", "seg_head.2", "c4_up_conv.0", "c3_up_conv.0",
                                "c2_up_conv.0", "c1_up_conv.0", "c3_cat_refine.0", "c2_cat_refine.0", "c1_cat_refine.0",
                                "c0_cat_refine.0"};
  for (int l = 0; l < 3; ++l)
    for (int b = 0; b < blocks[l]; ++b) {
      const std::string p = "layer" + std::to_string(l + 1) + "." + std::to_string(b);
      v.push_back(p + ".conv1"); v.push_back(p + ".conv2"); v.push_back(p + ".conv3");
      if (b == 0) v.push_back(p + ".downsample.0");
    }
  for (int l = 0; l < 4; ++l) {
    v.push_back("skip_combine." + std::to_string(l) + ".up.0");
    v.push_back("skip_combine." + std::to_string(l) + ".cat_conv.0");
  }
  for (int s = 0; s < 4; ++s)
    for (int h = 0; h < 3; ++h) {
      v.push_back(std::string(kHeadNames[h]) + "_c" + std::to_string(s) + ".0");
      v.push_back(std::string(kHeadNames[h]) + "_c" + std::to_string(s) + ".2");
    }
  return v;
}

static int upload_conv(ConvW& c) {
  if (c.d_w) { cudaFree(c.d_w); c.d_w = nullptr; }
  if (c.d_b) { cudaFree(c.d_b); c.d_b = nullptr; }
  KG_CUDA_CHECK(cudaMalloc(&c.d_w, c.h_w.size() * sizeof(float)));
  KG_CUDA_CHECK(cudaMalloc(&c.d_b, c.h_b.size() * sizeof(float)));
  KG_CUDA_CHECK(cudaMemcpy(c.d_w, c.h_w.data(), c.h_w.size() * sizeof(float), cudaMemcpyHostToDevice));
  KG_CUDA_CHECK(cudaMemcpy(c.d_b, c.h_b.data(), c.h_b.size() * sizeof(float), cudaMemcpyHostToDevice));
  return KG_OK;
}

// nn.Conv2d weight [Cout,Cin,R,S] (+ bias) followed by an eval-mode nn.BatchNorm2d (KGnet.py:131-132,72-77) folded:
// w' = w * g / sqrt(var + eps), b' = beta + (b - mean) * g / sqrt(var + eps); repacked tap-major [tap][cin][cout].
static int set_conv(Net* net, const char* name, const float* w, int Cout, int Cin, int R, int S, const float* bias,
                    const float* bn_w, const float* bn_b, const float* bn_mean, const float* bn_var, double eps) {
  KG_REQUIRE(net && name && w, "kg_net_set_conv: null argument");
  KG_REQUIRE(Cout > 0 && Cin > 0 && R > 0 && S > 0, "kg_net_set_conv(%s): bad shape", name);
  ConvW& c = net->convs[name];
  c.name = name; c.Cout = Cout; c.Cin = Cin; c.R = R; c.S = S;
  // lazily built packings of the PREVIOUS weights (stem image, shift-add slabs) must not survive a weight update
  c.stemw = TcStemWeights(); c.stemw_u8 = TcStemWeights(); c.shift1 = TcShiftPacked(); c.shift3 = TcShiftPacked();
  c.h_w.assign((size_t)R * S * Cin * Cout, 0.f);
  c.h_b.assign(Cout, 0.f);
  for (int co = 0; co < Cout; ++co) {
    double sc = 1.0, sh = bias ? (double)bias[co] : 0.0;
    if (bn_w) {
      sc = (double)bn_w[co] / std::sqrt((double)bn_var[co] + eps);
      sh = (double)bn_b[co] + (sh - (double)bn_mean[co]) * sc;
    }
    c.h_b[co] = (float)sh;
    for (int ci = 0; ci < Cin; ++ci)
      for (int t = 0; t < R * S; ++t)
        c.h_w[((size_t)t * Cin + ci) * Cout + co] = (float)((double)w[((size_t)co * Cin + ci) * R * S + t] * sc);
  }
  net->finalized = false;
  return KG_OK;
}

static int finalize(Net* net) {
  for (const auto& nme : expected_convs(net->blocks))
    KG_REQUIRE(net->convs.count(nme) == 1, "kg_net_finalize: weights of '%s' were not set", nme.c_str());
  // fused first-layer head convs: the three heads of a scale share their input (KGnet.py:300-316) -> one conv, N = 3C
  for (int s = 0; s < 4; ++s) {
    const int C = kHeadC[s];
    ConvW f;
    f.name = "heads_l1_c" + std::to_string(s);
    f.Cout = 3 * C; f.Cin = C; f.R = 7; f.S = 7;
    f.h_w.assign((size_t)49 * C * 3 * C, 0.f);
    f.h_b.assign(3 * C, 0.f);
    for (int h = 0; h < 3; ++h) {
      const ConvW& src = net->convs.at(std::string(kHeadNames[h]) + "_c" + std::to_string(s) + ".0");
      KG_REQUIRE(src.Cout == C && src.Cin == C && src.R == 7 && src.S == 7, "head %s has an unexpected shape", src.name.c_str());
      for (size_t tc = 0; tc < (size_t)49 * C; ++tc)
        for (int co = 0; co < C; ++co) f.h_w[tc * 3 * C + h * C + co] = src.h_w[tc * C + co];
      for (int co = 0; co < C; ++co) f.h_b[h * C + co] = src.h_b[co];
    }
    ConvW& dst = net->convs[f.name];
    if (dst.d_w) cudaFree(dst.d_w);
    if (dst.d_b) cudaFree(dst.d_b);
    tc_free_weights(dst.tc);
    dst = std::move(f);
  }
  for (auto& kv : net->convs) {
    ConvW& c = kv.second;
    KG_TRY(upload_conv(c));
    if (tc_layer_supported(c.Cin, c.Cout, c.R, c.S)) KG_TRY(tc_pack_weights(c.h_w.data(), c.Cin, c.Cout, c.R, c.S, &c.tc));
  }
  net->plan.reset();
  net->finalized = true;
  return KG_OK;
}

// ------------------------------------------------------------------------------------------------
struct PlanBuilder {
  Net* net; Plan* p; size_t off = 0; int N;
  int alloc(int H, int W, int C, bool lo = true) {
    Tensor t; t.H = H; t.W = W; t.C = C;
    const size_t bytes = align_up((size_t)N * H * W * C * sizeof(__half), 1024);
    t.off_hi = off; off += bytes;
    if (lo) { t.off_lo = off; off += bytes; } else { t.off_lo = (size_t)-1; }
    p->tensors.push_back(t);
    return (int)p->tensors.size() - 1;
  }
  const ConvW* W_(const std::string& n) { return &net->convs.at(n); }

  // generic conv over the batch
  int conv(const std::string& wname, int in0, int in1, int stride, int pad, bool relu, int stage, int res = -1, int out32_ext = -1,
           bool sigmoid = false, int in0_coff = 0, bool x_input = false, int xH = 0, int xW = 0, bool want_out = true,
           bool out_lo = true) {
    const ConvW* w = W_(wname);
    Op op; op.type = OP_CONV; op.w = w; op.in0 = in0; op.in1 = in1; op.res = res; op.stride = stride; op.pad = pad;
    op.relu = relu; op.sigmoid = sigmoid; op.stage = stage; op.out32_ext = out32_ext; op.in0_coff = in0_coff; op.x_input = x_input;
    int Hin, Win;
    if (x_input) { Hin = xH; Win = xW; op.C0 = w->Cin; op.C1 = 0; }
    else {
      const Tensor& t0 = p->tensors[in0];
      Hin = t0.H; Win = t0.W;
      op.C1 = in1 >= 0 ? p->tensors[in1].C : 0;
      op.C0 = w->Cin - op.C1;
    }
    op.Hout = (Hin + 2 * pad - w->R) / stride + 1;
    op.Wout = (Win + 2 * pad - w->S) / stride + 1;
    if (want_out) op.out = alloc(op.Hout, op.Wout, w->Cout, out_lo);
    op.out_single = !out_lo;
    op.prob_off = p->h_conv_probs.size(); op.nprob = N; op.max_pix = op.Hout * op.Wout;
    for (int n = 0; n < N; ++n) {
      ConvProb pb{};
      pb.Hin = Hin; pb.Win = Win; pb.Hout = op.Hout; pb.Wout = op.Wout;
      if (x_input) { pb.in0_off = (long long)n * w->Cin * Hin * Win; }
      else {
        const Tensor& t0 = p->tensors[in0];
        pb.in0_off = (long long)n * Hin * Win * t0.C + in0_coff; pb.in0_pitch = Win * t0.C;
        if (in1 >= 0) { const Tensor& t1 = p->tensors[in1]; pb.in1_off = (long long)n * Hin * Win * t1.C; pb.in1_pitch = Win * t1.C; }
      }
      pb.out_off = (long long)n * op.Hout * op.Wout * w->Cout; pb.out_pitch = op.Wout * w->Cout;
      if (res >= 0) { const Tensor& tr = p->tensors[res]; pb.res_off = (long long)n * tr.H * tr.W * tr.C; pb.res_pitch = tr.W * tr.C; }
      pb.out32_off = (long long)n * w->Cout * op.Hout * op.Wout;
      p->h_conv_probs.push_back(pb);
    }
    p->ops.push_back(op);
    return op.out;
  }
  int bilinear(int in, int Hout, int Wout) {
    const Tensor t = p->tensors[in];
    Op op; op.type = OP_BILINEAR; op.in0 = in; op.out = alloc(Hout, Wout, t.C); op.stage = ST_RESIZE;
    op.prob_off = p->h_resize_probs.size(); op.nprob = N; op.max_pix = Hout * Wout; op.Hout = Hout; op.Wout = Wout;
    for (int n = 0; n < N; ++n) {
      ResizeProb pb{};
      pb.in_off = (long long)n * t.H * t.W * t.C; pb.Hin = t.H; pb.Win = t.W; pb.in_pitch = t.W * t.C;
      pb.out_off = (long long)n * Hout * Wout * t.C; pb.Hout = Hout; pb.Wout = Wout; pb.out_pitch = Wout * t.C;
      p->h_resize_probs.push_back(pb);
    }
    p->ops.push_back(op);
    return op.out;
  }
  int maxpool(int in) {
    const Tensor t = p->tensors[in];
    Op op; op.type = OP_MAXPOOL; op.in0 = in; op.stage = ST_POOL;
    op.Hout = (t.H + 2 - 3) / 2 + 1; op.Wout = (t.W + 2 - 3) / 2 + 1;
    op.out = alloc(op.Hout, op.Wout, t.C);
    p->ops.push_back(op);
    return op.out;
  }
  void exportf(int in, int ext) {
    Op op; op.type = OP_EXPORT; op.in0 = in; op.out32_ext = ext; op.stage = ST_EXPORT;
    p->ops.push_back(op);
  }
};

// Workspace placement by liveness.  The builder above hands out consecutive offsets (21 GB for a bs32 / 512^2 step); every kernel of
// a pass runs on one stream in op order, so a tensor's planes can be reused once its last reader has been enqueued.  First-fit over a
// free list of byte ranges, in op order: a tensor is placed when its producer is reached and released AFTER its last consumer (an op
// never writes into the memory of its own inputs).  The five feature maps forward_seg reads later (and kg_net_import_feats writes)
// are never released.  KG_NO_WS_REUSE=1 keeps the consecutive layout (A/B, debugging).
// Core of the placement, free of plan types (kg_debug_place_by_liveness exposes it to the CPU tests): buffer b of `bytes[b]` bytes is
// written first by op def[b] and read last by op last[b] (last[b] >= n_ops: never released).  Returns the arena size.
static size_t first_fit_by_liveness(int n_ops, const std::vector<int>& def, const std::vector<int>& last, const std::vector<size_t>& bytes,
                                    std::vector<size_t>& off) {
  const int nb = (int)bytes.size();
  struct Range { size_t off, len; };
  std::vector<Range> freel;                                                          // sorted by offset, coalesced
  size_t top = 0;
  auto take = [&](size_t len) -> size_t {
    for (size_t i = 0; i < freel.size(); ++i)
      if (freel[i].len >= len) {
        const size_t o = freel[i].off;
        freel[i].off += len; freel[i].len -= len;
        if (freel[i].len == 0) freel.erase(freel.begin() + (long)i);
        return o;
      }
    const size_t o = top; top += len; return o;
  };
  auto give = [&](size_t o, size_t len) {
    if (len == 0) return;
    size_t i = 0;
    while (i < freel.size() && freel[i].off < o) ++i;
    freel.insert(freel.begin() + (long)i, Range{o, len});
    if (i + 1 < freel.size() && freel[i].off + freel[i].len == freel[i + 1].off) { freel[i].len += freel[i + 1].len; freel.erase(freel.begin() + (long)i + 1); }
    if (i > 0 && freel[i - 1].off + freel[i - 1].len == freel[i].off) { freel[i - 1].len += freel[i].len; freel.erase(freel.begin() + (long)i); }
  };
  std::vector<std::vector<int>> born(n_ops), dies(n_ops + 1);
  for (int b = 0; b < nb; ++b) {
    born[def[b]].push_back(b);
    if (last[b] < n_ops) dies[last[b] + 1].push_back(b);                             // released once its last reader (op last[b]) is enqueued
  }
  off.assign(nb, 0);
  for (int i = 0; i < n_ops; ++i) {
    for (int b : dies[i]) give(off[b], bytes[b]);
    for (int b : born[i]) off[b] = take(bytes[b]);
  }
  return top;
}

static void place_by_liveness(Plan* p) {
  if (getenv("KG_NO_WS_REUSE") != nullptr) return;
  const int nt = (int)p->tensors.size(), nop = (int)p->ops.size();
  std::vector<int> def(nt, -1), last(nt, -1);
  for (int i = 0; i < nop; ++i) {
    const Op& op = p->ops[i];
    if (op.out >= 0 && def[op.out] < 0) def[op.out] = i;
    for (int t : {op.in0, op.in1, op.res})
      if (t >= 0) last[t] = std::max(last[t], i);
  }
  for (int l = 0; l < 5; ++l) last[p->feat_ids[l]] = nop + 1;                       // persistent
  for (int t = 0; t < nt; ++t) { if (def[t] < 0) def[t] = 0; last[t] = std::max(last[t], def[t]); }
  // one buffer per plane: hi of tensor t = 2 t, lo = 2 t + 1 (size 0 when the tensor has no lo plane)
  std::vector<int> bdef(2 * nt), blast(2 * nt);
  std::vector<size_t> bytes(2 * nt), off;
  for (int t = 0; t < nt; ++t) {
    const Tensor& T = p->tensors[t];
    const size_t len = align_up((size_t)p->N * T.H * T.W * T.C * sizeof(__half), 1024);
    bdef[2 * t] = bdef[2 * t + 1] = def[t]; blast[2 * t] = blast[2 * t + 1] = last[t];
    bytes[2 * t] = len; bytes[2 * t + 1] = T.off_lo != (size_t)-1 ? len : 0;
  }
  p->bytes = first_fit_by_liveness(nop, bdef, blast, bytes, off);
  for (int t = 0; t < nt; ++t) {
    Tensor& T = p->tensors[t];
    T.off_hi = off[2 * t];
    if (T.off_lo != (size_t)-1) T.off_lo = off[2 * t + 1];
  }
}

// precision: 0 = CUDA-core fp32 everywhere (on-device reference), 1 = "fast": tensor cores, split-fp16 3-pass in the
// backbone/decoder and single-pass fp16 in the heads, 2 = "exact": tensor cores, 3-pass everywhere.
static void assign_tc(Plan* p, int precision) {
  if (precision == 0 || !tc_available()) return;
  for (auto& op : p->ops) {
    if (op.type != OP_CONV || op.x_input) continue;
    const ConvW* w = op.w;
    if (!w->tc.valid || (op.stride != 1 && !(op.stride == 2 && tc_stride2_enabled()))) continue;
    if (op.C0 % 64 != 0 || (op.C1 % 64) != 0) continue;
    const bool head = op.stage == ST_HEAD1 || op.stage == ST_HEAD2;
    op.tc_passes = (head && precision == 1) ? 1 : 3;
    // "fast": the three 64-channel 3x3 decoder convs keep split activations but single-plane weights (2 passes).  CPU
    // emulation (tools/precision_emulate.py, "<group>:w") puts the keypoint-map error at 6-7e-4 with them, 5.8e-4 without.
    // Measured at 512x512 (tests/test_parity_full_gpu.py): kp error 7.5e-4 with the three 64-channel convs, 8.0e-4 with c3_up / c4_up
    // added (decoder 12.6 -> 10.9 ms); the single-pass heads dominate the error either way.
    if (precision == 1 && getenv("KG_NO_2PASS") == nullptr &&
        (w->name == "c0_conv.2" || w->name == "c1_up_conv.0" || w->name == "c2_up_conv.0" || w->name == "c3_up_conv.0" ||
         w->name == "c4_up_conv.0")) op.tc_passes = 2;
    // experiments: KG_2PASS_EXTRA / KG_1PASS_EXTRA = comma-separated name fragments of further layers to run 2-pass / single-pass in "fast"
    auto listed = [&](const char* var) {
      const char* extra = getenv(var);
      if (extra == nullptr) return false;
      const std::string list(extra);
      size_t pos = 0;
      while (pos <= list.size()) {
        const size_t e = list.find(',', pos);
        const std::string frag = list.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
        if (!frag.empty() && w->name.find(frag) != std::string::npos) return true;
        if (e == std::string::npos) break;
        pos = e + 1;
      }
      return false;
    };
    if (precision == 1 && op.tc_passes == 3 && listed("KG_2PASS_EXTRA")) op.tc_passes = 2;
    if (precision == 1 && op.tc_passes != 1 && listed("KG_1PASS_EXTRA")) op.tc_passes = 1;
  }
}

static int build_plan(Net* net, int N, int H, int W, int precision) {
  KG_REQUIRE(N >= 1 && H >= 16 && W >= 16 && H % 16 == 0 && W % 16 == 0,
             "forward_dec: N=%d H=%d W=%d (H, W must be multiples of 16: the decoder's x2 upsampling assumes it)", N, H, W);
  std::unique_ptr<Plan> p(new Plan());
  p->N = N; p->H = H; p->W = W; p->precision = precision;
  PlanBuilder b{net, p.get(), 0, N};
  const bool fast_heads = precision == 1 && tc_available();
  // KGnet.py:276-286
  int c0a = b.conv("c0_conv.0", -1, -1, 1, 1, true, ST_STEM, -1, -1, false, 0, true, H, W);
  int c0 = b.conv("c0_conv.2", c0a, -1, 1, 1, true, ST_DECODER);
  int c1 = b.conv("conv1", -1, -1, 2, 3, true, ST_STEM, -1, -1, false, 0, true, H, W);
  int x = b.maxpool(c1);
  int feats[3];
  for (int l = 0; l < 3; ++l) {
    for (int blk = 0; blk < net->blocks[l]; ++blk) {
      const std::string pre = "layer" + std::to_string(l + 1) + "." + std::to_string(blk);
      const int stride = (blk == 0 && l > 0) ? 2 : 1;
      int t1 = b.conv(pre + ".conv1", x, -1, 1, 0, true, ST_BACKBONE);
      int t2 = b.conv(pre + ".conv2", t1, -1, stride, 1, true, ST_BACKBONE);
      int idn = x;
      if (blk == 0) idn = b.conv(pre + ".downsample.0", x, -1, stride, 0, false, ST_BACKBONE);
      x = b.conv(pre + ".conv3", t2, -1, 1, 0, true, ST_BACKBONE, idn);
    }
    feats[l] = x;
  }
  const int c2 = feats[0], c3 = feats[1], c4 = feats[2];
  auto T = [&](int id) -> const Tensor& { return p->tensors[id]; };
  // KGnet.py:288-298 (cat order: upsampled first, skip second)
  int u4 = b.bilinear(c4, T(c3).H, T(c3).W);
  int c4u = b.conv("c4_up_conv.0", u4, -1, 1, 1, true, ST_DECODER);
  int c3c = b.conv("c3_cat_refine.0", c4u, c3, 1, 0, true, ST_DECODER);
  int u3 = b.bilinear(c3c, T(c2).H, T(c2).W);
  int c3u = b.conv("c3_up_conv.0", u3, -1, 1, 1, true, ST_DECODER);
  int c2c = b.conv("c2_cat_refine.0", c3u, c2, 1, 0, true, ST_DECODER);
  int u2 = b.bilinear(c2c, T(c1).H, T(c1).W);
  int c2u = b.conv("c2_up_conv.0", u2, -1, 1, 1, true, ST_DECODER);
  int c1c = b.conv("c1_cat_refine.0", c2u, c1, 1, 0, true, ST_DECODER);
  int u1 = b.bilinear(c1c, T(c0).H, T(c0).W);
  int c1u = b.conv("c1_up_conv.0", u1, -1, 1, 1, true, ST_DECODER);
  // c0_cat is read by the first-layer heads only: single-pass heads never touch its lo plane
  int c0c = b.conv("c0_cat_refine.0", c1u, c0, 1, 0, true, ST_DECODER, -1, -1, false, 0, false, 0, 0, true, !fast_heads);
  // KGnet.py:300-316
  const int cats[4] = {c0c, c1c, c2c, c3c};
  for (int s = 0; s < 4; ++s) {
    const int C = kHeadC[s];
    int h1 = b.conv("heads_l1_c" + std::to_string(s), cats[s], -1, 1, 3, true, ST_HEAD1, -1, -1, false, 0, false, 0, 0, true,
                    !fast_heads);
    const int outs[3] = {kHeadOut[0], kHeadOut[1], kHeadOut[2]};
    if (fast_heads && tc_shift_supported(T(h1).H, T(h1).W, 7, 7, 3, C, 3, outs, false)) {
      // the three second-layer head convs of this scale as ONE row-GEMM + shift-add launch (tc_shift.cu)
      Op op; op.type = OP_HEADS2; op.in0 = h1; op.head_scale = s; op.stage = ST_HEAD2; op.C0 = C;
      op.Hout = T(h1).H; op.Wout = T(h1).W; op.tc_passes = 1;
      p->ops.push_back(op);
      continue;
    }
    for (int h = 0; h < 3; ++h) {
      const std::string nme = std::string(kHeadNames[h]) + "_c" + std::to_string(s) + ".2";
      b.conv(nme, h1, -1, 1, 3, false, ST_HEAD2, -1, 3 * s + h, h == 0, h * C, false, 0, 0, false);
      Op& op = p->ops.back();
      op.C0 = C; op.C1 = 0; op.in_single = fast_heads;
    }
  }
  const int fe[5] = {c0, c1, c2, c3, c4};
  for (int l = 0; l < 5; ++l) { p->feat_ids[l] = fe[l]; b.exportf(fe[l], 12 + l); }
  p->bytes = b.off;
  place_by_liveness(p.get());
  assign_tc(p.get(), precision);
  KG_CUDA_CHECK(cudaMalloc(&p->d_conv_probs, p->h_conv_probs.size() * sizeof(ConvProb)));
  KG_CUDA_CHECK(cudaMemcpy(p->d_conv_probs, p->h_conv_probs.data(), p->h_conv_probs.size() * sizeof(ConvProb), cudaMemcpyHostToDevice));
  KG_CUDA_CHECK(cudaMalloc(&p->d_resize_probs, p->h_resize_probs.size() * sizeof(ResizeProb)));
  KG_CUDA_CHECK(cudaMemcpy(p->d_resize_probs, p->h_resize_probs.data(), p->h_resize_probs.size() * sizeof(ResizeProb),
                           cudaMemcpyHostToDevice));
  net->plan = std::move(p);
  net->seg.valid = false;
  return KG_OK;
}

static int ensure_plan(Net* net, int N, int H, int W, int precision) {
  KG_REQUIRE(net != nullptr, "null net handle");
  if (!net->finalized) { set_error("kg_net: call kg_net_finalize first"); return KG_ERR_STATE; }
  KG_REQUIRE(precision >= 0 && precision <= 2, "precision=%d (0 cuda-core fp32, 1 fast, 2 exact)", precision);
  if (precision != 0 && !tc_available()) {
    set_error("precision=%d needs the tcgen05 path, which failed to initialise: %s", precision, tc_status());
    return KG_ERR_STATE;
  }
  if (net->plan && net->plan->N == N && net->plan->H == H && net->plan->W == W && net->plan->precision == precision) return KG_OK;
  return build_plan(net, N, H, W, precision);
}

// One 3x3 conv with 64 (NHWC out) or 1 (fp32 out) output channels as a row-GEMM + shift-add launch, or false if unsupported.
static bool shift_conv_ok(const ConvW* w, int H, int W, int stride, int pad, int Cin_total, int C1, bool nhwc_out, bool has_res) {
  if (stride != 1 || C1 != 0 || has_res || w->R != 3 || w->S != 3 || pad != 1) return false;
  const int no[1] = {w->Cout};
  return tc_shift_supported(H, W, 3, 3, 1, Cin_total, 1, no, nhwc_out);
}

static int shift_conv_fill(TcShiftOp* t, const ConvW* w, int N, int H, int W, int Cin, int passes) {
  t->N = N; t->H = H; t->W = W; t->R = 3; t->S = 3; t->pad = 1; t->Cin = Cin; t->passes = passes; t->n_groups = 1;
  t->g[0].h_w = w->h_w.data(); t->g[0].d_bias = w->d_b; t->g[0].n_out = w->Cout; t->g[0].in_coff = 0;
  TcShiftPacked& pk = passes == 3 ? w->shift3 : w->shift1;      // passes 1 and 2 share the hi-only packing
  if (!pk.valid()) KG_TRY(tc_shift_pack(t, &pk));
  t->packed = pk;
  return KG_OK;
}

struct Ptrs {
  char* ws;
  __half* hi(const Tensor& t) const { return reinterpret_cast<__half*>(ws + t.off_hi); }
  __half* lo(const Tensor& t) const { return t.off_lo == (size_t)-1 ? nullptr : reinterpret_cast<__half*>(ws + t.off_lo); }
};

// d_x: fp32 NCHW input, or nullptr with d_img = the uint8 NHWC image (tensor-core precisions only: the stems normalise on the fly)
static int run_plan(Net* net, const float* d_x, const uint8_t* d_img, float* const* ext, bool want_feats, void* ws, cudaStream_t stream,
                    int* n_launches) {
  Plan* p = net->plan.get();
  Ptrs P{(char*)ws};
  int launches = 0;
  if (p->tc_workspace != ws) {
    // (re-)encode the TMA descriptors of every tensor-core op for this workspace base
    p->tc_ops.clear();
    p->shift_ops.clear();
    for (auto& op : p->ops) {
      if (op.type == OP_HEADS2) {
        const Tensor& t0 = p->tensors[op.in0];
        TcShiftOp t{};
        t.N = p->N; t.H = t0.H; t.W = t0.W; t.R = 7; t.S = 7; t.pad = 3; t.Cin = op.C0; t.in_hi = P.hi(t0); t.in_C = t0.C; t.n_groups = 3;
        for (int h = 0; h < 3; ++h) {
          const ConvW& w = net->convs.at(std::string(kHeadNames[h]) + "_c" + std::to_string(op.head_scale) + ".2");
          t.g[h].h_w = w.h_w.data(); t.g[h].d_bias = w.d_b; t.g[h].n_out = w.Cout; t.g[h].in_coff = h * op.C0; t.g[h].sigmoid = h == 0;
        }
        if (getenv("KG_TC_DEBUG")) fprintf(stderr, "[op %d heads_l2_c%d] ", (int)(&op - p->ops.data()), op.head_scale);
        KG_TRY(tc_shift_prepare(&t));
        op.tc_index = (int)p->shift_ops.size();
        p->shift_ops.push_back(t);
        continue;
      }
      if (op.type != OP_CONV || op.tc_passes == 0) continue;
      const Tensor& t0 = p->tensors[op.in0];
      op.use_shift = false;
      if (op.out >= 0 && op.out32_ext < 0 && !op.sigmoid && op.in0_coff == 0 && t0.C == op.C0 &&
          shift_conv_ok(op.w, op.Hout, op.Wout, op.stride, op.pad, op.C0, op.C1, true, op.res >= 0) && op.w->Cout == 64) {
        TcShiftOp t{};
        t.in_hi = P.hi(t0); t.in_lo = op.in_single ? nullptr : P.lo(t0); t.in_C = t0.C;
        const Tensor& to = p->tensors[op.out];
        t.out_hi = P.hi(to); t.out_lo = op.out_single ? nullptr : P.lo(to); t.relu = op.relu;
        if (getenv("KG_TC_DEBUG")) fprintf(stderr, "[op %d %s] ", (int)(&op - p->ops.data()), op.w->name.c_str());
        KG_TRY(shift_conv_fill(&t, op.w, p->N, op.Hout, op.Wout, op.C0, op.tc_passes));
        KG_TRY(tc_shift_prepare(&t));
        op.use_shift = true;
        op.tc_index = (int)p->shift_ops.size();
        p->shift_ops.push_back(t);
        continue;
      }
      TcConvOp t{};
      t.w = &op.w->tc; t.bias = op.w->d_b;
      t.N = p->N; t.H = op.Hout; t.W = op.Wout; t.R = op.w->R; t.S = op.w->S; t.pad = op.pad;
      t.stride = op.stride; t.Hin = t0.H; t.Win = t0.W;
      t.C0 = op.C0; t.C1 = op.C1; t.Cout = op.w->Cout; t.passes = op.tc_passes;
      t.in0_hi = P.hi(t0); t.in0_lo = op.in_single ? nullptr : P.lo(t0); t.in0_C = t0.C; t.in0_coff = op.in0_coff;
      if (op.in1 >= 0) { const Tensor& t1 = p->tensors[op.in1]; t.in1_hi = P.hi(t1); t.in1_lo = P.lo(t1); t.in1_C = t1.C; }
      if (op.out >= 0) { const Tensor& to = p->tensors[op.out]; t.out_hi = P.hi(to); t.out_lo = op.out_single ? nullptr : P.lo(to); }
      if (op.res >= 0) { const Tensor& tr = p->tensors[op.res]; t.res_hi = P.hi(tr); t.res_lo = P.lo(tr); }
      t.relu = op.relu; t.sigmoid = op.sigmoid;
      if (getenv("KG_TC_DEBUG")) fprintf(stderr, "[op %d %s] ", (int)(&op - p->ops.data()), op.w->name.c_str());
      KG_TRY(tc_conv_prepare(&t));
      op.tc_index = (int)p->tc_ops.size();
      p->tc_ops.push_back(t);
    }
    p->tc_workspace = ws;
  }
  // KG_TIMING_PER_OP=1 (diagnostics): every op of the plan gets its own timing slot 64 + index instead of its stage
  static const bool per_op = getenv("KG_TIMING_PER_OP") != nullptr;
  int op_index = -1;
  for (const Op& op : p->ops) {
    ++op_index;
    if (op.type == OP_EXPORT && !want_feats) continue;
    StageScope ts(per_op && 64 + op_index < KG_MAX_STAGES ? 64 + op_index : op.stage, stream);
    switch (op.type) {
      case OP_CONV: {
        if (op.tc_passes == 0 && !per_op) ts.restage(ST_STEM);   // stage 8 collects every CUDA-core conv, 9..12 are tensor-core only
        float* out32 = op.out32_ext >= 0 ? ext[op.out32_ext] : nullptr;
        if (op.tc_passes > 0 && op.use_shift) {
          KG_TRY(tc_shift_launch(&p->shift_ops[op.tc_index], nullptr, stream));
          ++launches;
          break;
        }
        if (op.tc_passes > 0) {
          KG_TRY(tc_conv_launch(&p->tc_ops[op.tc_index], out32, stream));
          ++launches;
          break;
        }
        const ConvW* w = op.w;
        if (op.x_input && w->Cin == 3 && w->Cout == 64 && op.relu && op.out >= 0 && !op.out_single && out32 == nullptr &&
            ((w->R == 3 && op.stride == 1) || (w->R == 7 && op.stride == 2)) && w->R == w->S && op.pad == w->R / 2) {
          const Tensor& to = p->tensors[op.out];
          if (d_img != nullptr) {
            KG_REQUIRE(p->precision != 0 && tc_stem_supported(w->R, op.stride),
                       "kg_net_forward_dec_u8: uint8 input needs a tensor-core precision (normalise with kg_preprocess_u8 for precision 0)");
            if (!w->stemw_u8.d_img) KG_TRY(tc_stem_pack_u8(w->h_w.data(), w->h_b.data(), w->R, &w->stemw_u8));
            KG_TRY(tc_stem_launch_u8(d_img, &w->stemw_u8, P.hi(to), P.lo(to), p->N, p->H, p->W, w->R, op.stride, stream));
            ++launches;
            break;
          }
          if (p->precision != 0 && tc_stem_supported(w->R, op.stride)) {
            if (!w->stemw.d_img) KG_TRY(tc_stem_pack(w->h_w.data(), w->R, &w->stemw));
            KG_TRY(tc_stem_launch(d_x, &w->stemw, w->d_b, P.hi(to), P.lo(to), p->N, p->H, p->W, w->R, op.stride, stream));
            ++launches;
            break;
          }
          KG_TRY(launch_stem_conv(d_x, w->d_w, w->d_b, P.hi(to), P.lo(to), p->N, p->H, p->W, w->R, op.stride, stream));
          ++launches;
          break;
        }
        ConvArgs a{};
        if (op.x_input) { KG_REQUIRE(d_x != nullptr, "forward_dec: this layer needs the fp32 input (uint8 input covers the tensor-core stems only)"); a.x32 = d_x; }
        else {
          const Tensor& t0 = p->tensors[op.in0];
          a.in0_hi = P.hi(t0); a.in0_lo = op.in_single ? nullptr : P.lo(t0); a.in0_ps = t0.C;
          if (op.in1 >= 0) { const Tensor& t1 = p->tensors[op.in1]; a.in1_hi = P.hi(t1); a.in1_lo = P.lo(t1); a.in1_ps = t1.C; }
        }
        a.C0 = op.C0; a.C1 = op.C1; a.w = w->d_w; a.bias = w->d_b; a.Cout = w->Cout; a.R = w->R; a.S = w->S;
        a.stride = op.stride; a.pad = op.pad; a.relu = op.relu; a.sigmoid = op.sigmoid;
        if (op.out >= 0) { const Tensor& to = p->tensors[op.out]; a.out_hi = P.hi(to); a.out_lo = op.out_single ? nullptr : P.lo(to); a.out_ps = to.C; }
        if (op.res >= 0) { const Tensor& tr = p->tensors[op.res]; a.res_hi = P.hi(tr); a.res_lo = P.lo(tr); a.res_ps = tr.C; }
        a.out32 = out32;
        a.probs = p->d_conv_probs + op.prob_off;
        KG_TRY(launch_conv_ffma(a, op.nprob, op.max_pix, stream));
        ++launches;
        break;
      }
      case OP_HEADS2: {
        float* outs[3] = {ext[3 * op.head_scale], ext[3 * op.head_scale + 1], ext[3 * op.head_scale + 2]};
        KG_TRY(tc_shift_launch(&p->shift_ops[op.tc_index], outs, stream));
        ++launches;
        break;
      }
      case OP_BILINEAR: {
        const Tensor& ti = p->tensors[op.in0]; const Tensor& to = p->tensors[op.out];
        static const bool no2x = getenv("KG_NO_BILINEAR2X") != nullptr;
        if (!no2x && to.H == 2 * ti.H && to.W == 2 * ti.W && to.C == ti.C) {          // the decoder's exact x2 steps (KGnet.py:288-297)
          KG_TRY(launch_bilinear2x(P.hi(ti), P.lo(ti), P.hi(to), P.lo(to), p->N, ti.H, ti.W, ti.C, stream));
          ++launches;
          break;
        }
        KG_TRY(launch_bilinear(P.hi(ti), P.lo(ti), ti.C, P.hi(to), P.lo(to), to.C, ti.C, p->d_resize_probs + op.prob_off, op.nprob,
                               op.max_pix, stream));
        ++launches;
        break;
      }
      case OP_MAXPOOL: {
        const Tensor& ti = p->tensors[op.in0]; const Tensor& to = p->tensors[op.out];
        KG_TRY(launch_maxpool3x3s2(P.hi(ti), P.lo(ti), P.hi(to), P.lo(to), p->N, ti.H, ti.W, ti.C, stream));
        ++launches;
        break;
      }
      case OP_EXPORT: {
        const Tensor& ti = p->tensors[op.in0];
        KG_TRY(launch_export_nchw(P.hi(ti), P.lo(ti), ext[op.out32_ext], p->N, ti.H * ti.W, ti.C, stream));
        ++launches;
        break;
      }
    }
  }
  if (n_launches) *n_launches = launches;
  return KG_OK;
}

// ------------------------------------------------------------------------------------------------
// forward_seg (KGnet.py:246-267,321-350): per-box crops at up to 5 levels, variable-size top-down mini decoder.
// All boxes of the batch run in grouped ("ragged") launches: one launch per layer, blockIdx.z = box.
static bool patch_rect(float y1n, float x1n, float y2n, float x2n, int h, int w, int* r) {
  // get_patches (KGnet.py:246-256): np.round is half-to-even, float32 arithmetic under NumPy 2 promotion rules
  int y1 = (int)rintf(y1n * (float)h), x1 = (int)rintf(x1n * (float)w);
  int y2 = (int)rintf(y2n * (float)h), x2 = (int)rintf(x2n * (float)w);
  y1 = y1 > 0 ? y1 : 0; x1 = x1 > 0 ? x1 : 0;
  y2 = y2 < h - 1 ? y2 : h - 1; x2 = x2 < w - 1 ? x2 : w - 1;
  if (y2 < y1 || x2 < x1 || y2 - y1 < 2 || x2 - x1 < 2) return false;
  r[0] = y1; r[1] = x1; r[2] = y2; r[3] = x2;
  return true;
}

static const int kSegUpIn[4] = {64, 256, 512, 1024}, kSegOut[4] = {64, 64, 256, 512};

static int seg_prepare_atlas(Net* net, int N, int H, int W, const int* counts, const double* boxes, size_t* ws_bytes, long long* mask_floats,
                             int* n_masks, int* mask_index, int* mask_hw, long long* mask_off, int* mask_pitch);

static int seg_prepare(Net* net, int N, int H, int W, const int* counts, const double* boxes, size_t* ws_bytes, long long* mask_floats,
                       int* n_masks, int* mask_index, int* mask_hw, long long* mask_off, int* mask_pitch) {
  KG_REQUIRE(net && counts && ws_bytes && mask_floats && n_masks, "kg_net_seg_prepare: null argument");
  if (net->plan && net->plan->precision != 0 && tc_available() && getenv("KG_SEG_FFMA") == nullptr)
    return seg_prepare_atlas(net, N, H, W, counts, boxes, ws_bytes, mask_floats, n_masks, mask_index, mask_hw, mask_off, mask_pitch);
  KG_REQUIRE(net->plan && net->plan->N == N && net->plan->H == H && net->plan->W == W,
             "kg_net_seg_prepare: forward_dec (or kg_net_import_feats) must run first for N=%d H=%d W=%d", N, H, W);
  Plan* p = net->plan.get();
  SegPlan& sp = net->seg;
  sp = SegPlan();
  sp.N = N; sp.H = H; sp.W = W;
  int fh[5], fw[5];
  for (int l = 0; l < 5; ++l) { fh[l] = p->tensors[p->feat_ids[l]].H; fw[l] = p->tensors[p->feat_ids[l]].W; }
  size_t off = 0;   // elements in one scratch plane
  auto alloc = [&](int h, int w, int C) { const size_t o = off; off += align_up((size_t)h * w * C, 128); return (long long)o; };
  std::vector<SegBox> sboxes;
  int bi = 0, mi = 0;
  long long moff = 0;
  for (int n = 0; n < N; ++n) {
    for (int k = 0; k < counts[n]; ++k, ++bi) {
      KG_REQUIRE(boxes != nullptr, "kg_net_seg_prepare: null boxes");
      const double* bx = boxes + (size_t)bi * 5;
      SegBox sb{}; sb.img = n; sb.L = 0; sb.mask_index = -1;
      const float y1 = (float)bx[0], x1 = (float)bx[1], y2 = (float)bx[2], x2 = (float)bx[3];   // np.asarray(box, np.float32) (KGnet.py:331)
      const float h0 = (float)fh[0], w0 = (float)fw[0];
      for (int l = 0; l < 5; ++l) {
        if (!patch_rect(y1 / h0, x1 / w0, y2 / h0, x2 / w0, fh[l], fw[l], sb.rect[l])) break;   // (:333-340)
        sb.L = l + 1;
      }
      if (sb.L > 0) {
        sb.mask_index = mi;
        const int mh = sb.rect[0][2] - sb.rect[0][0], mw = sb.rect[0][3] - sb.rect[0][1];
        if (mask_hw) { mask_hw[2 * mi] = mh; mask_hw[2 * mi + 1] = mw; }
        if (mask_pitch) mask_pitch[mi] = mw;
        if (mask_off) mask_off[mi] = moff;
        moff += (long long)mh * mw;
        ++mi;
      }
      if (mask_index) mask_index[bi] = sb.mask_index;
      sboxes.push_back(sb);
    }
  }
  sp.n_boxes = bi;
  auto raw_off = [&](const SegBox& sb, int l) {
    return (((long long)sb.img * fh[l] + sb.rect[l][0]) * fw[l] + sb.rect[l][1]) * kFeatC[l];
  };
  struct Pre { bool raw; int level; long long off; int h, w, C; };
  std::vector<Pre> pre(sboxes.size());
  for (size_t i = 0; i < sboxes.size(); ++i) {
    const SegBox& sb = sboxes[i];
    if (sb.L == 0) continue;
    const int l = sb.L - 1;
    pre[i] = Pre{true, l, raw_off(sb, l), sb.rect[l][2] - sb.rect[l][0], sb.rect[l][3] - sb.rect[l][1], kFeatC[l]};
  }
  for (int l = 3; l >= 0; --l) {       // mask_forward (KGnet.py:258-267): deepest level first
    SegPlan::Step& st = sp.steps[l];
    for (size_t i = 0; i < sboxes.size(); ++i) {
      const SegBox& sb = sboxes[i];
      if (sb.L - 1 <= l) continue;      // this box has no level l+1
      const int ph = sb.rect[l][2] - sb.rect[l][0], pw = sb.rect[l][3] - sb.rect[l][1];
      const Pre pr = pre[i];
      const long long u = alloc(ph, pw, kSegUpIn[l]), v = alloc(ph, pw, kSegOut[l]), c = alloc(ph, pw, kSegOut[l]);
      ResizeProb rp{};                  // F.interpolate(inputs2, inputs1.shape[2:]) (:110)
      rp.in_off = pr.off; rp.in_pitch = pr.raw ? fw[pr.level] * kFeatC[pr.level] : pr.w * pr.C;
      rp.Hin = pr.h; rp.Win = pr.w; rp.Hout = ph; rp.Wout = pw; rp.out_off = u; rp.out_pitch = pw * kSegUpIn[l];
      if (pr.raw) { st.rs_raw.push_back(rp); st.pix_raw = std::max(st.pix_raw, ph * pw); }
      else { st.rs_scr.push_back(rp); st.pix_scr = std::max(st.pix_scr, ph * pw); }
      ConvProb up{};                    // self.up: conv3x3 + ReLU
      up.in0_off = u; up.in0_pitch = pw * kSegUpIn[l]; up.Hin = ph; up.Win = pw; up.Hout = ph; up.Wout = pw;
      up.out_off = v; up.out_pitch = pw * kSegOut[l];
      st.up.push_back(up);
      ConvProb ct{};                    // self.cat_conv(torch.cat((inputs1, outputs2), 1)): patch first, upsampled second (:111)
      ct.in0_off = raw_off(sb, l); ct.in0_pitch = fw[l] * kFeatC[l];
      ct.in1_off = v; ct.in1_pitch = pw * kSegOut[l];
      ct.Hin = ph; ct.Win = pw; ct.Hout = ph; ct.Wout = pw; ct.out_off = c; ct.out_pitch = pw * kSegOut[l];
      st.cat.push_back(ct);
      st.pix = std::max(st.pix, ph * pw);
      pre[i] = Pre{false, l, c, ph, pw, kSegOut[l]};
    }
  }
  for (size_t i = 0; i < sboxes.size(); ++i) {   // seg_head (KGnet.py:145-147,344-345)
    const SegBox& sb = sboxes[i];
    if (sb.L == 0) continue;
    const Pre pr = pre[i];
    const int ph = pr.h, pw = pr.w;
    const long long t = alloc(ph, pw, 64);
    ConvProb a{};
    a.in0_off = pr.off; a.in0_pitch = pr.raw ? fw[0] * kFeatC[0] : pw * 64; a.Hin = ph; a.Win = pw; a.Hout = ph; a.Wout = pw;
    a.out_off = t; a.out_pitch = pw * 64;
    if (pr.raw) { sp.h0_raw.push_back(a); sp.pix_h0_raw = std::max(sp.pix_h0_raw, ph * pw); }
    else { sp.h0_scr.push_back(a); sp.pix_h0_scr = std::max(sp.pix_h0_scr, ph * pw); }
    ConvProb b{};
    b.in0_off = t; b.in0_pitch = pw * 64; b.Hin = ph; b.Win = pw; b.Hout = ph; b.Wout = pw;
    b.out32_off = mask_off ? mask_off[sb.mask_index] : 0;
    sp.h1.push_back(b);
    sp.pix_h1 = std::max(sp.pix_h1, ph * pw);
  }
  if (!mask_off) {   // offsets are needed internally even when the caller does not ask for them
    long long o = 0; size_t k = 0;
    for (size_t i = 0; i < sboxes.size(); ++i) {
      if (sboxes[i].L == 0) continue;
      sp.h1[k++].out32_off = o;
      o += (long long)(sboxes[i].rect[0][2] - sboxes[i].rect[0][0]) * (sboxes[i].rect[0][3] - sboxes[i].rect[0][1]);
    }
  }
  sp.scratch_halfs = off;
  sp.mask_floats = moff;
  sp.n_masks = mi;
  BlobWriter bw{net};
  for (int l = 0; l < 4; ++l) {
    auto& st = sp.steps[l];
    sp.o_rs_raw[l] = bw.put(st.rs_raw.data(), st.rs_raw.size() * sizeof(ResizeProb));
    sp.o_rs_scr[l] = bw.put(st.rs_scr.data(), st.rs_scr.size() * sizeof(ResizeProb));
    sp.o_up4[l] = bw.put(st.up.data(), st.up.size() * sizeof(ConvProb));
    sp.o_cat[l] = bw.put(st.cat.data(), st.cat.size() * sizeof(ConvProb));
  }
  sp.o_h0r = bw.put(sp.h0_raw.data(), sp.h0_raw.size() * sizeof(ConvProb));
  sp.o_h0s = bw.put(sp.h0_scr.data(), sp.h0_scr.size() * sizeof(ConvProb));
  sp.o_h1 = bw.put(sp.h1.data(), sp.h1.size() * sizeof(ConvProb));
  KG_TRY(bw.commit(sp));
  sp.blob_base = align_up(off * sizeof(__half), 256) * 2;
  sp.valid = true;
  *ws_bytes = sp.blob_base + align_up(sp.blob_bytes, 256) + 256;
  *mask_floats = moff;
  *n_masks = mi;
  return KG_OK;
}

static int seg_run_atlas(Net* net, void* dec_ws, void* seg_ws, size_t seg_bytes, float* d_masks, cudaStream_t stream, int* n_launches);

static int seg_run(Net* net, void* dec_ws, void* seg_ws, size_t seg_bytes, float* d_masks, cudaStream_t stream, int* n_launches) {
  SegPlan& sp = net->seg;
  if (!sp.valid || !net->plan) { set_error("kg_net_forward_seg: call kg_net_seg_prepare first"); return KG_ERR_STATE; }
  if (sp.atlas) return seg_run_atlas(net, dec_ws, seg_ws, seg_bytes, d_masks, stream, n_launches);
  const size_t plane_bytes = align_up(sp.scratch_halfs * sizeof(__half), 256);
  if (seg_bytes < sp.blob_base + sp.blob_bytes) { set_error("kg_net_forward_seg: workspace too small (%zu < %zu)", seg_bytes, sp.blob_base + sp.blob_bytes); return KG_ERR_WORKSPACE; }
  if (sp.n_masks == 0) { if (n_launches) *n_launches = 0; return KG_OK; }
  KG_REQUIRE(dec_ws && seg_ws && d_masks, "kg_net_forward_seg: null buffer");
  Plan* p = net->plan.get();
  Ptrs P{(char*)dec_ws};
  __half* s_hi = reinterpret_cast<__half*>(seg_ws);
  __half* s_lo = reinterpret_cast<__half*>((char*)seg_ws + plane_bytes);
  // the problem descriptors: one asynchronous copy from the pinned staging buffer seg_prepare filled
  KG_CUDA_CHECK(cudaMemcpyAsync((char*)seg_ws + sp.blob_base, net->h_blob, sp.blob_bytes, cudaMemcpyHostToDevice, stream));
  KG_CUDA_CHECK(cudaEventRecord(net->blob_copied, stream));
  const size_t *o_rs_raw = sp.o_rs_raw, *o_rs_scr = sp.o_rs_scr, *o_up = sp.o_up4, *o_cat = sp.o_cat;
  const size_t o_h0r = sp.o_h0r, o_h0s = sp.o_h0s, o_h1 = sp.o_h1;
  char* dp = (char*)seg_ws + sp.blob_base;
  int launches = 0;
  StageScope ts(ST_SEG, stream);
  auto feat = [&](int l) -> const Tensor& { return p->tensors[p->feat_ids[l]]; };
  auto conv = [&](const std::string& wname, const __half* i0h, const __half* i0l, int C0, int ps0, const __half* i1h, const __half* i1l, int C1,
                  int ps1, __half* oh, __half* ol, int ops, float* o32, bool relu, bool sig, const void* probs, int nprob, int pix) -> int {
    if (nprob == 0) return KG_OK;
    const ConvW& w = net->convs.at(wname);
    ConvArgs a{};
    a.in0_hi = i0h; a.in0_lo = i0l; a.C0 = C0; a.in0_ps = ps0; a.in1_hi = i1h; a.in1_lo = i1l; a.C1 = C1; a.in1_ps = ps1;
    a.w = w.d_w; a.bias = w.d_b; a.Cout = w.Cout; a.R = w.R; a.S = w.S; a.stride = 1; a.pad = w.R / 2;
    a.out_hi = oh; a.out_lo = ol; a.out_ps = ops; a.out32 = o32; a.relu = relu; a.sigmoid = sig;
    a.probs = reinterpret_cast<const ConvProb*>(probs);
    ++launches;
    return launch_conv_ffma(a, nprob, pix, stream);
  };
  for (int l = 3; l >= 0; --l) {
    auto& st = sp.steps[l];
    if (st.up.empty()) continue;
    const std::string pre = "skip_combine." + std::to_string(l);
    const Tensor& fd = feat(l + 1);
    if (!st.rs_raw.empty()) {
      KG_TRY(launch_bilinear(P.hi(fd), P.lo(fd), fd.C, s_hi, s_lo, kSegUpIn[l], kSegUpIn[l],
                             reinterpret_cast<const ResizeProb*>(dp + o_rs_raw[l]), (int)st.rs_raw.size(), st.pix_raw, stream));
      ++launches;
    }
    if (!st.rs_scr.empty()) {
      KG_TRY(launch_bilinear(s_hi, s_lo, kSegUpIn[l], s_hi, s_lo, kSegUpIn[l], kSegUpIn[l],
                             reinterpret_cast<const ResizeProb*>(dp + o_rs_scr[l]), (int)st.rs_scr.size(), st.pix_scr, stream));
      ++launches;
    }
    KG_TRY(conv(pre + ".up.0", s_hi, s_lo, kSegUpIn[l], kSegUpIn[l], nullptr, nullptr, 0, 0, s_hi, s_lo, kSegOut[l], nullptr, true, false,
                dp + o_up[l], (int)st.up.size(), st.pix));
    const Tensor& fl = feat(l);
    KG_TRY(conv(pre + ".cat_conv.0", P.hi(fl), P.lo(fl), fl.C, fl.C, s_hi, s_lo, kSegOut[l], kSegOut[l], s_hi, s_lo, kSegOut[l], nullptr, true,
                false, dp + o_cat[l], (int)st.cat.size(), st.pix));
  }
  const Tensor& f0 = feat(0);
  KG_TRY(conv("seg_head.0", P.hi(f0), P.lo(f0), 64, 64, nullptr, nullptr, 0, 0, s_hi, s_lo, 64, nullptr, true, false, dp + o_h0r,
              (int)sp.h0_raw.size(), sp.pix_h0_raw));
  KG_TRY(conv("seg_head.0", s_hi, s_lo, 64, 64, nullptr, nullptr, 0, 0, s_hi, s_lo, 64, nullptr, true, false, dp + o_h0s,
              (int)sp.h0_scr.size(), sp.pix_h0_scr));
  KG_TRY(conv("seg_head.2", s_hi, s_lo, 64, 64, nullptr, nullptr, 0, 0, nullptr, nullptr, 0, d_masks, false, true, dp + o_h1, (int)sp.h1.size(),
              sp.pix_h1));
  if (n_launches) *n_launches = launches;
  return KG_OK;
}

// ---- forward_seg on the tensor cores --------------------------------------------------------------------------
// The per-box mini decoders all share weights, so the crops of every box are packed, level by level, into one "atlas"
// image per level (shelf packing, >= 1 pixel of zeros around every box = the zero padding of the per-box 3x3 convs).
// Each layer then is ONE dense tcgen05 conv over the atlas; an uint8 validity mask zeroes the gap pixels in the
// epilogue so that the next 3x3 conv again sees zero padding.  Mask patches are strided windows of the fp32 atlas.
static int seg_prepare_atlas(Net* net, int N, int H, int W, const int* counts, const double* boxes, size_t* ws_bytes, long long* mask_floats,
                             int* n_masks, int* mask_index, int* mask_hw, long long* mask_off, int* mask_pitch) {
  KG_REQUIRE(net->plan && net->plan->N == N && net->plan->H == H && net->plan->W == W,
             "kg_net_seg_prepare: forward_dec (or kg_net_import_feats) must run first for N=%d H=%d W=%d", N, H, W);
  Plan* p = net->plan.get();
  SegPlan& sp = net->seg;
  sp = SegPlan();
  sp.N = N; sp.H = H; sp.W = W; sp.atlas = true;
  int fh[5], fw[5];
  for (int l = 0; l < 5; ++l) { fh[l] = p->tensors[p->feat_ids[l]].H; fw[l] = p->tensors[p->feat_ids[l]].W; }
  std::vector<SegBox> sboxes;
  int bi = 0, mi = 0;
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < counts[n]; ++k, ++bi) {
      KG_REQUIRE(boxes != nullptr, "kg_net_seg_prepare: null boxes");
      const double* bx = boxes + (size_t)bi * 5;
      SegBox sb{}; sb.img = n; sb.L = 0; sb.mask_index = -1;
      const float y1 = (float)bx[0], x1 = (float)bx[1], y2 = (float)bx[2], x2 = (float)bx[3];
      const float h0 = (float)fh[0], w0 = (float)fw[0];
      for (int l = 0; l < 5; ++l) {
        if (!patch_rect(y1 / h0, x1 / w0, y2 / h0, x2 / w0, fh[l], fw[l], sb.rect[l])) break;
        sb.L = l + 1;
      }
      if (sb.L > 0) sb.mask_index = mi++;
      if (mask_index) mask_index[bi] = sb.mask_index;
      sboxes.push_back(sb);
    }
  sp.n_boxes = bi; sp.n_masks = mi;
  const int nb = (int)sboxes.size();
  // shelf packing per level (boxes sorted by height, tallest first; stable => deterministic)
  std::vector<int> px[5], py[5];
  for (int l = 0; l < 5; ++l) {
    px[l].assign(nb, -1); py[l].assign(nb, -1);
    std::vector<int> idx;
    int maxw = 0; long long area = 0;
    for (int i = 0; i < nb; ++i)
      if (sboxes[i].L > l) {
        idx.push_back(i);
        const int h = sboxes[i].rect[l][2] - sboxes[i].rect[l][0], w = sboxes[i].rect[l][3] - sboxes[i].rect[l][1];
        maxw = std::max(maxw, w); area += (long long)(h + 1) * (w + 1);
      }
    SegPlan::Level& L = sp.lv[l];
    if (idx.empty()) continue;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
      return sboxes[a].rect[l][2] - sboxes[a].rect[l][0] > sboxes[b].rect[l][2] - sboxes[b].rect[l][0];
    });
    int WA = 128;
    while ((long long)WA * WA < area && WA < 4096) WA *= 2;      // roughly square atlas, power-of-two width >= 128
    while (WA < maxw + 2) WA *= 2;
    int x = 1, y = 1, shelf_h = 0;
    for (int i : idx) {
      const int h = sboxes[i].rect[l][2] - sboxes[i].rect[l][0], w = sboxes[i].rect[l][3] - sboxes[i].rect[l][1];
      if (x + w + 1 > WA) { x = 1; y += shelf_h + 1; shelf_h = 0; }
      px[l][i] = x; py[l][i] = y;
      x += w + 1; shelf_h = std::max(shelf_h, h);
    }
    L.WA = WA; L.HA = y + shelf_h + 1;
  }
  // scratch layout: one plane of halfs (x2 for hi/lo), then the masks
  size_t off = 0;
  auto alloc = [&](const SegPlan::Level& L, int C) { const size_t o = off; off += align_up((size_t)L.HA * L.WA * C, 512); return o; };
  for (int l = 0; l < 5; ++l) {
    SegPlan::Level& L = sp.lv[l];
    if (L.HA == 0) continue;
    L.P = alloc(L, kFeatC[l]);
    if (l < 4) { L.U = alloc(L, kSegUpIn[l]); L.V = alloc(L, kSegOut[l]); L.Cc = alloc(L, kSegOut[l]); }
    if (l == 0) L.T = alloc(L, 64);
  }
  sp.scratch_halfs = off;
  size_t moff = 0;
  for (int l = 0; l < 5; ++l) { sp.lv[l].mask = moff; moff += align_up((size_t)sp.lv[l].HA * sp.lv[l].WA, 256); }
  sp.mask_base = align_up(off * sizeof(__half), 256) * 2;
  // problem lists
  for (int i = 0; i < nb; ++i) {
    const SegBox& sb = sboxes[i];
    for (int l = 0; l < sb.L; ++l) {
      SegPlan::Level& L = sp.lv[l];
      const int h = sb.rect[l][2] - sb.rect[l][0], w = sb.rect[l][3] - sb.rect[l][1];
      const long long apos = (long long)py[l][i] * L.WA + px[l][i];
      ResizeProb cp{};      // exact copy (same size) of the feature crop into the atlas
      cp.in_off = (((long long)sb.img * fh[l] + sb.rect[l][0]) * fw[l] + sb.rect[l][1]) * kFeatC[l]; cp.in_pitch = fw[l] * kFeatC[l];
      cp.Hin = h; cp.Win = w; cp.Hout = h; cp.Wout = w; cp.out_off = (long long)L.P + apos * kFeatC[l]; cp.out_pitch = L.WA * kFeatC[l];
      L.crop.push_back(cp); L.pix_crop = std::max(L.pix_crop, h);
      RectProb rp{}; rp.off = apos; rp.h = h; rp.w = w; rp.pitch = L.WA;
      L.rects.push_back(rp);
      if (l == sb.L - 1 && l < 4) {     // deepest level of this box: its running tensor is the raw crop (mask_forward, KGnet.py:261-262)
        ResizeProb dp = cp;
        dp.in_off = (long long)L.P + apos * kFeatC[l]; dp.in_pitch = L.WA * kFeatC[l];
        dp.out_off = (long long)L.Cc + apos * kSegOut[l]; dp.out_pitch = L.WA * kSegOut[l];
        L.deepest.push_back(dp); L.pix_deepest = std::max(L.pix_deepest, h);
      }
      if (l + 1 < sb.L) {               // pre(level l+1) -> bilinear -> U_l (KGnet.py:110)
        SegPlan::Level& D = sp.lv[l + 1];
        const int dh = sb.rect[l + 1][2] - sb.rect[l + 1][0], dw = sb.rect[l + 1][3] - sb.rect[l + 1][1];
        const long long dpos = (long long)py[l + 1][i] * D.WA + px[l + 1][i];
        ResizeProb up{};
        const int Cs = kSegUpIn[l];
        up.in_off = (long long)(l + 1 == 4 ? D.P : D.Cc) + dpos * Cs; up.in_pitch = D.WA * Cs; up.Hin = dh; up.Win = dw;
        up.Hout = h; up.Wout = w; up.out_off = (long long)L.U + apos * Cs; up.out_pitch = L.WA * Cs;
        up.frame = 1;                   // + the ring of zeros the 3x3 `up` conv reads around the box (the gaps are >= 1 px wide)
        L.up.push_back(up); L.pix_up = std::max(L.pix_up, h + 2);
      }
    }
    if (sb.L > 0) {
      const int m = sb.mask_index;
      const int h = sb.rect[0][2] - sb.rect[0][0], w = sb.rect[0][3] - sb.rect[0][1];
      if (mask_hw) { mask_hw[2 * m] = h; mask_hw[2 * m + 1] = w; }
      if (mask_pitch) mask_pitch[m] = sp.lv[0].WA;
      if (mask_off) mask_off[m] = (long long)py[0][i] * sp.lv[0].WA + px[0][i];
    }
  }
  sp.mask_floats = (long long)sp.lv[0].HA * sp.lv[0].WA;
  BlobWriter bw{net};
  for (int l = 0; l < 5; ++l) {
    auto& L = sp.lv[l];
    sp.o_crop[l] = bw.put(L.crop.data(), L.crop.size() * sizeof(ResizeProb));
    sp.o_deep[l] = bw.put(L.deepest.data(), L.deepest.size() * sizeof(ResizeProb));
    sp.o_up[l] = bw.put(L.up.data(), L.up.size() * sizeof(ResizeProb));
    sp.o_rect[l] = bw.put(L.rects.data(), L.rects.size() * sizeof(RectProb));
  }
  KG_TRY(bw.commit(sp));
  sp.blob_base = align_up(sp.mask_base + moff, 256);
  sp.valid = true;
  *ws_bytes = sp.blob_base + align_up(sp.blob_bytes, 256) + 256;
  *mask_floats = sp.mask_floats;
  *n_masks = mi;
  return KG_OK;
}

static int seg_run_atlas(Net* net, void* dec_ws, void* seg_ws, size_t seg_bytes, float* d_masks, cudaStream_t stream, int* n_launches) {
  SegPlan& sp = net->seg;
  Plan* p = net->plan.get();
  const size_t need = sp.blob_base + sp.blob_bytes;
  if (seg_bytes < need) { set_error("kg_net_forward_seg: workspace too small (%zu < %zu)", seg_bytes, need); return KG_ERR_WORKSPACE; }
  if (sp.n_masks == 0) { if (n_launches) *n_launches = 0; return KG_OK; }
  KG_REQUIRE(dec_ws && seg_ws && d_masks, "kg_net_forward_seg: null buffer");
  Ptrs P{(char*)dec_ws};
  const size_t plane_bytes = align_up(sp.scratch_halfs * sizeof(__half), 256);
  __half* s_hi = reinterpret_cast<__half*>(seg_ws);
  __half* s_lo = reinterpret_cast<__half*>((char*)seg_ws + plane_bytes);
  uint8_t* masks = reinterpret_cast<uint8_t*>((char*)seg_ws + sp.mask_base);
  // the problem lists: one asynchronous copy from the pinned staging buffer seg_prepare filled (no allocation, no sync here)
  KG_CUDA_CHECK(cudaMemcpyAsync((char*)seg_ws + sp.blob_base, net->h_blob, sp.blob_bytes, cudaMemcpyHostToDevice, stream));
  KG_CUDA_CHECK(cudaEventRecord(net->blob_copied, stream));
  const size_t *o_crop = sp.o_crop, *o_deep = sp.o_deep, *o_up = sp.o_up, *o_rect = sp.o_rect;
  const char* dp = (const char*)seg_ws + sp.blob_base;
  int launches = 0;
  StageScope ts(ST_SEG, stream);
  // precision "fast": the whole mask branch runs single-pass fp16 on hi planes only (KG_SEG_1PASS_LEVELS=n keeps the levels
  // >= n in split-fp16 3-pass).  Measured against the oracle: 3e-3 on the sigmoid output with O(1) logits, and identical
  // thresholded masks with raw Kaiming weights (tests/test_net_gpu.py); the 3-pass deep levels cost 2.4 ms for no visible gain.
  const char* sp1 = getenv("KG_SEG_1PASS_LEVELS");
  const int one_pass_levels = p->precision == 1 ? (sp1 ? atoi(sp1) : 5) : 0;   // levels [0, one_pass_levels) are single-pass
  size_t mask_total = 0;
  for (int l = 0; l < 5; ++l) mask_total += align_up((size_t)sp.lv[l].HA * sp.lv[l].WA, 256);
  KG_CUDA_CHECK(cudaMemsetAsync(masks, 0, mask_total, stream));
  // single-pass levels keep only the hi plane of every atlas tensor (their convs never read lo): half the HBM traffic
  auto lo_of = [&](int l) -> __half* { return l < one_pass_levels ? nullptr : s_lo; };
  // crops -> P_l, validity masks
  for (int l = 0; l < 5; ++l) {
    auto& L = sp.lv[l];
    if (L.crop.empty()) continue;
    const Tensor& f = p->tensors[p->feat_ids[l]];
    KG_TRY(launch_copy_rects(P.hi(f), lo_of(l) ? P.lo(f) : nullptr, f.C, s_hi, lo_of(l), f.C, f.C,
                             reinterpret_cast<const ResizeProb*>(dp + o_crop[l]), (int)L.crop.size(), L.pix_crop, stream));
    KG_TRY(launch_fill_rects(masks + L.mask, reinterpret_cast<const RectProb*>(dp + o_rect[l]), (int)L.rects.size(), stream));
    launches += 2;
  }
  auto conv = [&](const std::string& wname, const SegPlan::Level& L, size_t in0, int C0, size_t in1, int C1, size_t out, int Cout_t,
                  float* out32, bool relu, bool sig) -> int {
    const int lvl = (int)(&L - &sp.lv[0]);
    const int passes = lvl < one_pass_levels ? 1 : 3;
    const ConvW& w = net->convs.at(wname);
    __half* lo = lo_of(lvl);
    if (shift_conv_ok(&w, L.HA, L.WA, 1, w.R / 2, C0, C1, out32 == nullptr, false) && (w.Cout == 64 || w.Cout == 1)) {
      TcShiftOp t{};
      t.in_hi = s_hi + in0; t.in_lo = lo ? lo + in0 : nullptr; t.in_C = C0;
      if (out32 == nullptr) { t.out_hi = s_hi + out; t.out_lo = lo ? lo + out : nullptr; t.mask = masks + L.mask; }
      t.relu = relu; t.g[0].sigmoid = sig;
      KG_TRY(shift_conv_fill(&t, &w, 1, L.HA, L.WA, C0, passes));
      t.g[0].sigmoid = sig;
      KG_TRY(tc_shift_prepare(&t));
      ++launches;
      float* outs[1] = {out32};
      return tc_shift_launch(&t, outs, stream);
    }
    TcConvOp t{};
    t.w = &w.tc; t.bias = w.d_b; t.N = 1; t.H = L.HA; t.W = L.WA; t.R = w.R; t.S = w.S; t.pad = w.R / 2;
    t.C0 = C0; t.C1 = C1; t.Cout = w.Cout; t.passes = passes;
    t.in0_hi = s_hi + in0; t.in0_lo = lo ? lo + in0 : nullptr; t.in0_C = C0;
    if (C1 > 0) { t.in1_hi = s_hi + in1; t.in1_lo = lo ? lo + in1 : nullptr; t.in1_C = C1; }
    if (out32 == nullptr) { t.out_hi = s_hi + out; t.out_lo = lo ? lo + out : nullptr; }
    (void)Cout_t;
    t.relu = relu; t.sigmoid = sig; t.mask = masks + L.mask;
    KG_TRY(tc_conv_prepare(&t));
    ++launches;
    return tc_conv_launch(&t, out32, stream);
  };
  for (int l = 3; l >= 0; --l) {
    auto& L = sp.lv[l];
    if (L.HA == 0) continue;
    __half* lo = lo_of(l);
    if (!L.up.empty()) {
      // U_l is written inside the boxes that have a deeper level plus a one-pixel frame of zeros around each (ResizeProb::frame); every
      // other pixel of U_l stays uninitialised: valid conv outputs never read it and the outputs computed there are masked to zero
      KG_TRY(launch_bilinear_rows(s_hi, lo_of(l + 1), kSegUpIn[l], s_hi, lo, kSegUpIn[l], kSegUpIn[l],
                                  reinterpret_cast<const ResizeProb*>(dp + o_up[l]), (int)L.up.size(), L.pix_up, stream));
      ++launches;
      const std::string pre = "skip_combine." + std::to_string(l);
      KG_TRY(conv(pre + ".up.0", L, L.U, kSegUpIn[l], 0, 0, L.V, kSegOut[l], nullptr, true, false));
      KG_TRY(conv(pre + ".cat_conv.0", L, L.P, kFeatC[l], L.V, kSegOut[l], L.Cc, kSegOut[l], nullptr, true, false));   // cat((patch, up), 1) (:111)
    }
    if (L.up.empty()) {   // no conv wrote C_l: clear it so that the gaps read by the next 3x3 conv are zeros
      const size_t cbytes = (size_t)L.HA * L.WA * kSegOut[l] * sizeof(__half);
      KG_CUDA_CHECK(cudaMemsetAsync(s_hi + L.Cc, 0, cbytes, stream));
      if (lo) KG_CUDA_CHECK(cudaMemsetAsync(lo + L.Cc, 0, cbytes, stream));
    }
    if (!L.deepest.empty()) {
      KG_TRY(launch_copy_rects(s_hi, lo, kFeatC[l], s_hi, lo, kSegOut[l], kFeatC[l], reinterpret_cast<const ResizeProb*>(dp + o_deep[l]),
                               (int)L.deepest.size(), L.pix_deepest, stream));
      ++launches;
    }
  }
  const auto& L0 = sp.lv[0];
  KG_TRY(conv("seg_head.0", L0, L0.Cc, 64, 0, 0, L0.T, 64, nullptr, true, false));
  KG_TRY(conv("seg_head.2", L0, L0.T, 64, 0, 0, 0, 1, d_masks, false, true));
  if (n_launches) *n_launches = launches;
  return KG_OK;
}

// Single-layer entry used by the unit tests (one nn.Conv2d [+ReLU] [+residual], KGnet.py:45-61 style).
static int conv2d_nchw(const float* d_x, int N, int Cin, int H, int W, const float* h_w, const float* h_bias, int Cout, int R, int S,
                       int stride, int pad, int relu, const float* d_res, int mode, float* d_y, cudaStream_t stream) {
  KG_REQUIRE(d_x && h_w && d_y && N > 0 && Cin > 0 && Cout > 0, "kg_conv2d_nchw: bad arguments");
  KG_REQUIRE(mode == 0 || mode == 1 || mode == 2 || mode == 3 || mode == 11 || mode == 12 || mode == 13 || mode == 21,
             "kg_conv2d_nchw: mode must be 0 (cuda cores), 1 / 2 / 3 (tensor-core passes) or 11 / 12 / 13 (row-GEMM + shift-add kernel, 1 / 2 / 3 passes) or 21 (single pass, hi-plane output: CTA-pair kernel where eligible)");
  Net tmp;
  KG_TRY(set_conv(&tmp, "c", h_w, Cout, Cin, R, S, h_bias, nullptr, nullptr, nullptr, nullptr, 0.0));
  ConvW& w = tmp.convs.at("c");
  KG_TRY(upload_conv(w));
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
  const size_t in_e = (size_t)N * H * W * Cin, out_e = (size_t)N * Ho * Wo * Cout;
  __half *xh = nullptr, *xl = nullptr, *yh = nullptr, *yl = nullptr, *rh = nullptr, *rl = nullptr;
  ConvProb* d_probs = nullptr;
  auto cleanup = [&]() { cudaFree(xh); cudaFree(xl); cudaFree(yh); cudaFree(yl); cudaFree(rh); cudaFree(rl); cudaFree(d_probs); };
  int rc = KG_OK;
  do {
    if (cudaMalloc(&xh, in_e * 2) || cudaMalloc(&xl, in_e * 2) || cudaMalloc(&yh, out_e * 2) || cudaMalloc(&yl, out_e * 2)) { rc = KG_ERR_CUDA; break; }
    if ((rc = launch_import_nchw(d_x, xh, xl, N, H * W, Cin, stream)) != KG_OK) break;
    if (d_res) {
      if (cudaMalloc(&rh, out_e * 2) || cudaMalloc(&rl, out_e * 2)) { rc = KG_ERR_CUDA; break; }
      if ((rc = launch_import_nchw(d_res, rh, rl, N, Ho * Wo, Cout, stream)) != KG_OK) break;
    }
    if (mode == 0) {
      std::vector<ConvProb> probs(N);
      for (int n = 0; n < N; ++n) {
        ConvProb pb{};
        pb.Hin = H; pb.Win = W; pb.Hout = Ho; pb.Wout = Wo; pb.in0_off = (long long)n * H * W * Cin; pb.in0_pitch = W * Cin;
        pb.out_off = (long long)n * Ho * Wo * Cout; pb.out_pitch = Wo * Cout; pb.res_off = pb.out_off; pb.res_pitch = pb.out_pitch;
        probs[n] = pb;
      }
      if (cudaMalloc(&d_probs, N * sizeof(ConvProb)) || cudaMemcpy(d_probs, probs.data(), N * sizeof(ConvProb), cudaMemcpyHostToDevice)) { rc = KG_ERR_CUDA; break; }
      ConvArgs a{};
      a.in0_hi = xh; a.in0_lo = xl; a.C0 = Cin; a.in0_ps = Cin; a.w = w.d_w; a.bias = w.d_b; a.Cout = Cout; a.R = R; a.S = S;
      a.stride = stride; a.pad = pad; a.out_hi = yh; a.out_lo = yl; a.out_ps = Cout; a.res_hi = rh; a.res_lo = rl; a.res_ps = Cout;
      a.relu = relu; a.probs = d_probs;
      if ((rc = launch_conv_ffma(a, N, Ho * Wo, stream)) != KG_OK) break;
    } else if (mode >= 10 && mode < 20) {
      const bool nhwc = Cout != 1;
      if (d_res != nullptr || !shift_conv_ok(&w, Ho, Wo, stride, pad, Cin, 0, nhwc, false)) {
        set_error("kg_conv2d_nchw: shape not supported by the shift-add kernel"); rc = KG_ERR_INVALID; break;
      }
      TcShiftOp t{};
      t.in_hi = xh; t.in_lo = xl; t.in_C = Cin; t.relu = relu != 0;
      if (nhwc) { t.out_hi = yh; t.out_lo = yl; }
      if ((rc = shift_conv_fill(&t, &w, N, Ho, Wo, Cin, mode - 10)) != KG_OK) break;
      if ((rc = tc_shift_prepare(&t)) != KG_OK) break;
      float* outs[1] = {d_y};
      if ((rc = tc_shift_launch(&t, outs, stream)) != KG_OK) break;
      if (!nhwc) {
        if (cudaStreamSynchronize(stream) != cudaSuccess) { set_error("kg_conv2d_nchw: %s", cudaGetErrorString(cudaGetLastError())); rc = KG_ERR_CUDA; }
        break;
      }
    } else {
      if ((stride != 1 && stride != 2) || !tc_layer_supported(Cin, Cout, R, S) || Cout % 16 != 0 || 2 * pad != R - 1 || R != S) {
        set_error("kg_conv2d_nchw: shape not supported by the tensor-core path"); rc = KG_ERR_INVALID; break;
      }
      if ((rc = tc_pack_weights(w.h_w.data(), Cin, Cout, R, S, &w.tc)) != KG_OK) break;
      TcConvOp t{};
      t.w = &w.tc; t.bias = w.d_b; t.N = N; t.H = Ho; t.W = Wo; t.R = R; t.S = S; t.pad = pad; t.C0 = Cin; t.C1 = 0; t.Cout = Cout;
      t.stride = stride; t.Hin = H; t.Win = W;
      t.passes = mode == 21 ? 1 : mode; t.in0_hi = xh; t.in0_lo = xl; t.in0_C = Cin; t.out_hi = yh; t.out_lo = mode == 21 ? nullptr : yl;
      t.res_hi = rh; t.res_lo = rl; t.relu = relu != 0;
      if ((rc = tc_conv_prepare(&t)) != KG_OK) break;
      if ((rc = tc_conv_launch(&t, nullptr, stream)) != KG_OK) break;
    }
    if ((rc = launch_export_nchw(yh, mode == 21 ? nullptr : yl, d_y, N, Ho * Wo, Cout, stream)) != KG_OK) break;
    if (cudaStreamSynchronize(stream) != cudaSuccess) { set_error("kg_conv2d_nchw: %s", cudaGetErrorString(cudaGetLastError())); rc = KG_ERR_CUDA; }
  } while (0);
  if (rc == KG_ERR_CUDA && cudaPeekAtLastError() != cudaSuccess) set_error("kg_conv2d_nchw: CUDA error %s", cudaGetErrorString(cudaGetLastError()));
  cleanup();
  return rc;
}

// Unit-test entry of the row-GEMM + shift-add head kernel (tc_shift.cu): the three second-layer head convs of one scale
// (7x7, Cin -> 5 / 10 / 40; KGnet.py:161-209) on an fp32 NCHW input with 3*Cin channels (head h reads channels [h*Cin, (h+1)*Cin)).
static int heads_l2_nchw(const float* d_x, int N, int Cin, int H, int W, const float* const* h_w, const float* const* h_bias,
                         float* const* d_y, cudaStream_t stream) {
  KG_REQUIRE(d_x && h_w && h_bias && d_y && N > 0 && Cin > 0, "kg_heads_l2_nchw: bad arguments");
  const int outs[3] = {kHeadOut[0], kHeadOut[1], kHeadOut[2]};
  if (!tc_shift_supported(H, W, 7, 7, 3, Cin, 3, outs, false)) { set_error("kg_heads_l2_nchw: shape not supported by the shift-add kernel"); return KG_ERR_INVALID; }
  Net tmp;
  const size_t in_e = (size_t)N * H * W * 3 * Cin;
  __half* xh = nullptr;
  int rc = KG_OK;
  do {
    if (cudaMalloc(&xh, in_e * 2)) { rc = KG_ERR_CUDA; break; }
    if ((rc = launch_import_nchw(d_x, xh, nullptr, N, H * W, 3 * Cin, stream)) != KG_OK) break;
    TcShiftOp t{};
    t.N = N; t.H = H; t.W = W; t.R = 7; t.S = 7; t.pad = 3; t.Cin = Cin; t.in_hi = xh; t.in_C = 3 * Cin; t.n_groups = 3;
    for (int h = 0; h < 3; ++h) {
      const std::string nme = "h" + std::to_string(h);
      if ((rc = set_conv(&tmp, nme.c_str(), h_w[h], kHeadOut[h], Cin, 7, 7, h_bias[h], nullptr, nullptr, nullptr, nullptr, 0.0)) != KG_OK) break;
      ConvW& w = tmp.convs.at(nme);
      if ((rc = upload_conv(w)) != KG_OK) break;
      t.g[h].h_w = w.h_w.data(); t.g[h].d_bias = w.d_b; t.g[h].n_out = kHeadOut[h]; t.g[h].in_coff = h * Cin; t.g[h].sigmoid = h == 0;
    }
    if (rc != KG_OK) break;
    if ((rc = tc_shift_prepare(&t)) != KG_OK) break;
    if ((rc = tc_shift_launch(&t, d_y, stream)) != KG_OK) break;
    if (cudaStreamSynchronize(stream) != cudaSuccess) { set_error("kg_heads_l2_nchw: %s", cudaGetErrorString(cudaGetLastError())); rc = KG_ERR_CUDA; }
  } while (0);
  if (rc == KG_ERR_CUDA && cudaPeekAtLastError() != cudaSuccess) set_error("kg_heads_l2_nchw: CUDA error %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(xh);
  return rc;
}

}  // namespace kg

using namespace kg;

extern "C" {

int kg_net_create(kg_net** out, const int* blocks) {
  KG_REQUIRE(out != nullptr, "kg_net_create: null out");
  Net* n = new Net();
  if (blocks) for (int i = 0; i < 3; ++i) { KG_REQUIRE(blocks[i] >= 1 && blocks[i] <= 64, "kg_net_create: blocks[%d]=%d", i, blocks[i]); n->blocks[i] = blocks[i]; }
  *out = reinterpret_cast<kg_net*>(n);
  return KG_OK;
}

void kg_net_destroy(kg_net* h) { delete reinterpret_cast<Net*>(h); }

int kg_net_set_conv(kg_net* h, const char* name, const float* h_w, int Cout, int Cin, int R, int S, const float* h_bias,
                    const float* h_bn_weight, const float* h_bn_bias, const float* h_bn_mean, const float* h_bn_var, double bn_eps) {
  return set_conv(reinterpret_cast<Net*>(h), name, h_w, Cout, Cin, R, S, h_bias, h_bn_weight, h_bn_bias, h_bn_mean, h_bn_var, bn_eps);
}

int kg_net_finalize(kg_net* h) {
  KG_REQUIRE(h != nullptr, "kg_net_finalize: null handle");
  return finalize(reinterpret_cast<Net*>(h));
}

int kg_debug_place_by_liveness(int n_ops, int n_buffers, const int* def, const int* last, const unsigned long long* bytes,
                               unsigned long long* offsets, unsigned long long* total) {
  KG_REQUIRE(n_ops > 0 && n_buffers >= 0 && def && last && bytes && offsets && total, "kg_debug_place_by_liveness: bad arguments");
  std::vector<int> d(def, def + n_buffers), l(last, last + n_buffers);
  std::vector<size_t> by(n_buffers), off;
  for (int b = 0; b < n_buffers; ++b) {
    KG_REQUIRE(d[b] >= 0 && d[b] < n_ops && l[b] >= d[b], "kg_debug_place_by_liveness: buffer %d has def=%d last=%d", b, d[b], l[b]);
    by[b] = (size_t)bytes[b];
  }
  *total = first_fit_by_liveness(n_ops, d, l, by, off);
  for (int b = 0; b < n_buffers; ++b) offsets[b] = off[b];
  return KG_OK;
}

size_t kg_net_workspace_bytes(kg_net* h, int N, int H, int W, int precision) {
  Net* net = reinterpret_cast<Net*>(h);
  if (ensure_plan(net, N, H, W, precision) != KG_OK) return 0;
  return net->plan->bytes;
}

int kg_net_forward_dec(kg_net* h, const float* d_x, int N, int H, int W, float* const* d_heads, float* const* d_feats, int precision,
                       void* d_workspace, size_t workspace_bytes, void* stream, int* n_launches) {
  Net* net = reinterpret_cast<Net*>(h);
  KG_REQUIRE(d_x && d_heads && d_workspace, "kg_net_forward_dec: null argument");
  KG_TRY(ensure_plan(net, N, H, W, precision));
  if (workspace_bytes < net->plan->bytes) {
    set_error("kg_net_forward_dec: workspace too small (%zu < %zu bytes)", workspace_bytes, net->plan->bytes);
    return KG_ERR_WORKSPACE;
  }
  float* ext[17];
  for (int i = 0; i < 12; ++i) { KG_REQUIRE(d_heads[i] != nullptr, "kg_net_forward_dec: d_heads[%d] is null", i); ext[i] = d_heads[i]; }
  for (int i = 0; i < 5; ++i) ext[12 + i] = d_feats ? d_feats[i] : nullptr;
  return run_plan(net, d_x, nullptr, ext, d_feats != nullptr, d_workspace, (cudaStream_t)stream, n_launches);
}

int kg_net_forward_dec_u8(kg_net* h, const uint8_t* d_img, int N, int H, int W, float* const* d_heads, float* const* d_feats, int precision,
                          void* d_workspace, size_t workspace_bytes, void* stream, int* n_launches) {
  Net* net = reinterpret_cast<Net*>(h);
  KG_REQUIRE(d_img && d_heads && d_workspace, "kg_net_forward_dec_u8: null argument");
  KG_REQUIRE(precision != 0, "kg_net_forward_dec_u8: uint8 input needs a tensor-core precision (1 fast / 2 exact)");
  KG_TRY(ensure_plan(net, N, H, W, precision));
  if (workspace_bytes < net->plan->bytes) {
    set_error("kg_net_forward_dec_u8: workspace too small (%zu < %zu bytes)", workspace_bytes, net->plan->bytes);
    return KG_ERR_WORKSPACE;
  }
  float* ext[17];
  for (int i = 0; i < 12; ++i) { KG_REQUIRE(d_heads[i] != nullptr, "kg_net_forward_dec_u8: d_heads[%d] is null", i); ext[i] = d_heads[i]; }
  for (int i = 0; i < 5; ++i) ext[12 + i] = d_feats ? d_feats[i] : nullptr;
  return run_plan(net, nullptr, d_img, ext, d_feats != nullptr, d_workspace, (cudaStream_t)stream, n_launches);
}

int kg_net_import_feats(kg_net* h, const float* const* d_feats, int N, int H, int W, int precision, void* d_workspace, size_t workspace_bytes,
                        void* stream) {
  Net* net = reinterpret_cast<Net*>(h);
  KG_REQUIRE(d_feats && d_workspace, "kg_net_import_feats: null argument");
  KG_TRY(ensure_plan(net, N, H, W, precision));
  if (workspace_bytes < net->plan->bytes) { set_error("kg_net_import_feats: workspace too small"); return KG_ERR_WORKSPACE; }
  Ptrs P{(char*)d_workspace};
  for (int l = 0; l < 5; ++l) {
    const Tensor& t = net->plan->tensors[net->plan->feat_ids[l]];
    KG_REQUIRE(d_feats[l] != nullptr, "kg_net_import_feats: d_feats[%d] is null", l);
    KG_TRY(launch_import_nchw(d_feats[l], P.hi(t), P.lo(t), N, t.H * t.W, t.C, (cudaStream_t)stream));
  }
  return KG_OK;
}

int kg_net_seg_prepare(kg_net* h, int N, int H, int W, const int* box_counts, const double* h_boxes, size_t* seg_workspace_bytes,
                       long long* mask_floats, int* n_masks, int* h_mask_index, int* h_mask_hw, long long* h_mask_off, int* h_mask_pitch) {
  return seg_prepare(reinterpret_cast<Net*>(h), N, H, W, box_counts, h_boxes, seg_workspace_bytes, mask_floats, n_masks, h_mask_index,
                     h_mask_hw, h_mask_off, h_mask_pitch);
}

int kg_net_forward_seg(kg_net* h, void* d_dec_workspace, void* d_seg_workspace, size_t seg_workspace_bytes, float* d_masks, void* stream,
                       int* n_launches) {
  KG_REQUIRE(h != nullptr, "kg_net_forward_seg: null handle");
  return seg_run(reinterpret_cast<Net*>(h), d_dec_workspace, d_seg_workspace, seg_workspace_bytes, d_masks, (cudaStream_t)stream, n_launches);
}

int kg_conv2d_nchw(const float* d_x, int N, int Cin, int H, int W, const float* h_w, const float* h_bias, int Cout, int R, int S, int stride,
                   int pad, int relu, const float* d_res, int mode, float* d_y, void* stream) {
  return conv2d_nchw(d_x, N, Cin, H, W, h_w, h_bias, Cout, R, S, stride, pad, relu, d_res, mode, d_y, (cudaStream_t)stream);
}

int kg_heads_l2_nchw(const float* d_x, int N, int Cin, int H, int W, const float* const* h_w, const float* const* h_bias, float* const* d_y,
                     void* stream) {
  return heads_l2_nchw(d_x, N, Cin, H, W, h_w, h_bias, d_y, (cudaStream_t)stream);
}

/* out[0] = algorithmic FLOPs (2*MACs, real channel counts, one pass) of the tensor-core convs of the current plan,
 * out[1] = same for the CUDA-core convs, out[2] / out[3] = their launch counts, out[4..7] = tensor-core FLOPs by
 * stage (backbone, decoder, head layer 1, head layer 2). */
int kg_net_plan_info(kg_net* h, double* out, int n) {
  Net* net = reinterpret_cast<Net*>(h);
  KG_REQUIRE(net && net->plan && out && n >= 8, "kg_net_plan_info: no plan (run forward_dec first)");
  for (int i = 0; i < n; ++i) out[i] = 0.0;
  for (const Op& op : net->plan->ops) {
    if (op.type == OP_HEADS2) {
      const double f = 2.0 * net->plan->N * op.Hout * op.Wout * (double)(kHeadOut[0] + kHeadOut[1] + kHeadOut[2]) * op.C0 * 49;
      out[0] += f; out[2] += 1; out[4 + ST_HEAD2 - ST_BACKBONE] += f;
      continue;
    }
    if (op.type != OP_CONV) continue;
    const double f = 2.0 * net->plan->N * op.Hout * op.Wout * (double)op.w->Cout * (op.C0 + op.C1) * op.w->R * op.w->S;
    if (op.tc_passes > 0) { out[0] += f; out[2] += 1; if (op.stage >= ST_BACKBONE && op.stage <= ST_HEAD2) out[4 + op.stage - ST_BACKBONE] += f; }
    else { out[1] += f; out[3] += 1; }
  }
  return KG_OK;
}

int kg_tc_available(void) { return tc_available() ? 1 : 0; }
const char* kg_tc_status(void) { return tc_status(); }

}  // extern "C"
