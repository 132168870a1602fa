// "Row-GEMM + shift-add" convolution on tcgen05 for layers with FEW output channels (the second-layer 7x7 head
// convs of KGnet, Cout = 5 / 10 / 40: KGnet.py:161-209).  See tc_shift.cu.
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

struct TcShiftOp {
  int N = 0, H = 0, W = 0, R = 0, S = 0, pad = 0, Cin = 0;   // stride-1 "same" convs: output size == input size
  const __half* in_hi = nullptr;                              // NHWC fp16, pixel stride in_C channels
  int in_C = 0;
  int n_groups = 0;
  TcShiftGroup g[SH_MAX_GROUPS];
  // filled by tc_shift_prepare
  std::shared_ptr<void> params, d_weights;
  unsigned grid = 0, smem_bytes = 0;
};

bool tc_shift_supported(int H, int W, int R, int S, int pad, int Cin, int n_groups, const int* n_out);
int tc_shift_prepare(TcShiftOp* op);
// out32[g]: fp32 NCHW output of group g, [N, n_out, H, W]
int tc_shift_launch(const TcShiftOp* op, float* const* out32, cudaStream_t stream);

}  // namespace kg
