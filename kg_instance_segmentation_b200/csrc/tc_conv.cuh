// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (interface).  See tc_conv.cu.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include <memory>

#include "common.cuh"

namespace kg {

// Weights of one conv repacked for the tensor-core path: fp16 hi/lo planes, [tap][cout_pad][cin] (K-major B
// operand), multiplied by a power-of-two `scale` so that small weights stay in fp16's normal range.
struct TcWeights {
  bool valid = false;
  __half* d_hi = nullptr;
  __half* d_lo = nullptr;
  int cin = 0, cout = 0, cout_pad = 0, taps = 0;
  float inv_scale = 1.f;
};

struct TcParams;   // kernel parameter block (tensor maps + scalars), defined in tc_conv.cu

// One tensor-core conv op over a batch of N dense NHWC images (stride 1, "same" padding given by pad).
struct TcConvOp {
  const TcWeights* w = nullptr;
  const float* bias = nullptr;
  int N = 0, H = 0, W = 0, R = 0, S = 0, pad = 0;   // H, W: output spatial size
  int stride = 1, Hin = 0, Win = 0;                  // input spatial size (== H, W when stride is 1)
  int C0 = 0, C1 = 0, Cout = 0, passes = 1;
  const __half *in0_hi = nullptr, *in0_lo = nullptr, *in1_hi = nullptr, *in1_lo = nullptr;
  int in0_C = 0, in0_coff = 0, in1_C = 0;            // channel counts (pixel strides) of the source tensors
  __half *out_hi = nullptr, *out_lo = nullptr;
  const __half *res_hi = nullptr, *res_lo = nullptr;
  bool relu = false, sigmoid = false;
  const uint8_t* mask = nullptr;                     // optional per-pixel validity mask: outputs of masked-out pixels are 0
  // filled by tc_conv_prepare
  std::shared_ptr<void> params;                      // host copy of TcParams (tensor maps + scalars)
  unsigned grid_x = 0, grid_y = 0, smem_bytes = 0;
};

bool tc_available();
bool tc_stride2_enabled();
const char* tc_status();
bool tc_layer_supported(int cin, int cout, int R, int S);
int tc_pack_weights(const float* w_tap_cin_cout, int cin, int cout, int R, int S, TcWeights* out);
void tc_free_weights(TcWeights& w);
void* tc_encode_tiled_fn();   // cuTensorMapEncodeTiled entry point (null when the tensor-core path is unavailable)
int tc_num_sms();
int tc_conv_prepare(TcConvOp* op);
int tc_conv_launch(const TcConvOp* op, float* out32, cudaStream_t stream);

}  // namespace kg
