// Tensor-core kernels for the two convs that read the raw 3-channel image (c0_conv.0, conv1 + bn1).  See tc_stem.cu.
#pragma once
#include <cuda_fp16.h>

#include <memory>

#include "common.cuh"

namespace kg {

struct TcStemWeights {
  std::shared_ptr<void> d_img;   // device: fp16 hi / lo planes in the swizzled shared-memory layout of the kernel
  int K = 0;
  std::shared_ptr<void> d_bias_u8;   // uint8-input variant only: bias + 0.5 * sum(w / 255), 64 floats
  float inv_scale = 1.f;             // uint8-input variant only: the weight image is (w / 255) * 2^e, inv_scale = 2^-e
};

bool tc_stem_supported(int K, int stride);
int tc_stem_pack(const float* w_tap_cin_cout, int K, TcStemWeights* out);   // weights [tap][3][64] fp32 (BN folded)
// uint8 NHWC image in (the normalisation x / 255 - 0.5 of test.py:92 folded into weights and bias; see tc_stem.cu)
int tc_stem_pack_u8(const float* w_tap_cin_cout, const float* bias64, int K, TcStemWeights* out);
int tc_stem_launch_u8(const uint8_t* img, const TcStemWeights* w, __half* out_hi, __half* out_lo, int N, int H, int W, int K, int stride,
                      cudaStream_t s);
int tc_stem_launch(const float* x, const TcStemWeights* w, const float* bias, __half* out_hi, __half* out_lo, int N, int H, int W, int K,
                   int stride, cudaStream_t s);

}  // namespace kg
