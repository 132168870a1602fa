// "Row-GEMM + shift-add" convolution on tcgen05 for layers with FEW output channels per tap: the second-layer 7x7
// head convs of KGnet (Cout = 5 / 10 / 40, KGnet.py:161-209) and the 64-channel 3x3 convs (KGnet.py:139-158,109-120).
// See tc_shift.cu.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include <memory>

#include "common.cuh"

namespace kg {

constexpr int SH_MAX_GROUPS = 3;

// One conv of the fused launch: its own weights / bias / output, reading `Cin` channels at offset `in_coff` of the
// shared NHWC input tensor.
struct TcShiftGroup {
  const float* h_w = nullptr;     // host, [tap][cin][n_out] fp32 (BN folded)
  const float* d_bias = nullptr;  // device, [n_out]
  int n_out = 0, in_coff = 0;
  bool sigmoid = false;
};

// Packed device weights of one launch configuration (fp16 hi / lo planes, [R][rows][Cin]) + the per-conv 1/scale.
struct TcShiftPacked {
  std::shared_ptr<void> d_hi, d_lo;
  float inv_scale[SH_MAX_GROUPS] = {1.f, 1.f, 1.f};
  int passes = 0;
  bool valid() const { return d_hi != nullptr; }
};

struct TcShiftOp {
  int N = 0, H = 0, W = 0, R = 0, S = 0, pad = 0, Cin = 0;   // stride-1 "same" convs: output size == input size
  const __half* in_hi = nullptr;                              // NHWC fp16, pixel stride in_C channels
  const __half* in_lo = nullptr;                              // second plane (split fp16), needed when passes >= 2
  int in_C = 0;
  int passes = 1;                                             // 1: fp16 x fp16; 2: (hi + lo activations) x hi weights; 3: split-fp16 (hi*hi + lo*hi + hi*lo)
  int n_groups = 0;
  TcShiftGroup g[SH_MAX_GROUPS];
  // output: fp32 NCHW per conv (given at launch) unless out_hi is set: then conv 0 writes split-fp16 NHWC (+ReLU)
  __half* out_hi = nullptr;
  __half* out_lo = nullptr;
  bool relu = false;
  const uint8_t* mask = nullptr;                              // NHWC mode: outputs of pixels with mask == 0 are zero
  TcShiftPacked packed;                                       // optional: weights packed earlier by tc_shift_pack
  // filled by tc_shift_prepare
  std::shared_ptr<void> params;
  unsigned grid = 0, smem_bytes = 0;
  bool wres = false;                                          // weights resident in shared memory
};

// configurations with a compiled kernel: the three KGnet heads (5, 10, 40; 7x7; fp32 NCHW out), one 64-channel 3x3 conv
// (NHWC out) and one 1-channel 3x3 conv (fp32 out)
bool tc_shift_supported(int H, int W, int R, int S, int pad, int Cin, int n_groups, const int* n_out, bool nhwc_out);
int tc_shift_pack(const TcShiftOp* op, TcShiftPacked* out);
int tc_shift_prepare(TcShiftOp* op);
// out32[g]: fp32 NCHW output of conv g, [N, n_out, H, W] (ignored in NHWC mode)
int tc_shift_launch(const TcShiftOp* op, float* const* out32, cudaStream_t stream);

}  // namespace kg
