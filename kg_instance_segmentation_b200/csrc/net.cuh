// KGnet forward on device: shared structures of the conv / resize / pool kernels and the host-side plan.
// Activations are NHWC "split fp16": value = float(hi) + float(lo) (two planes of __half); tensors that are
// only consumed by single-pass tensor-core layers keep just the hi plane.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace kg {

// One convolution problem: an image of the batch, or one ROI crop of forward_seg.  Offsets are in ELEMENTS
// from the base pointers of ConvArgs; windows are (H, W) rectangles of a parent NHWC tensor with a row pitch.
struct ConvProb {
  long long in0_off, in1_off, out_off, res_off, out32_off;
  int Hin, Win, Hout, Wout;
  int in0_pitch, in1_pitch, out_pitch, res_pitch;
};

struct ConvArgs {
  const __half *in0_hi, *in0_lo, *in1_hi, *in1_lo;   // *_lo may be null (single-plane input)
  const float* x32;                                  // fp32 NCHW input (stem convs); in0_off = image offset
  int C0, C1, in0_ps, in1_ps;                        // channels taken from each source; pixel strides (elements)
  const float* w;                                    // [R*S][Cin][Cout] fp32, BN folded
  const float* bias;                                 // [Cout]
  int Cout, R, S, stride, pad;
  __half *out_hi, *out_lo;                           // NHWC outputs (either may be null)
  int out_ps;
  float* out32;                                      // fp32 NCHW output: out32[out32_off + (co*Hout + oy)*Wout + ox]
  const __half *res_hi, *res_lo;                     // residual added before the activation
  int res_ps;
  int relu, sigmoid;
  const ConvProb* probs;
};

struct ResizeProb {
  long long in_off, out_off;
  int Hin, Win, in_pitch, Hout, Wout, out_pitch;
  int frame;       // 1: additionally write a one-pixel frame of zeros around the Hout x Wout output rectangle (forward_seg atlases: the
  int pad_;        //    zero padding the following 3x3 conv reads), instead of clearing the whole atlas beforehand
};

struct RectProb { long long off; int h, w, pitch; };

// launchers (net_kernels.cu)
int launch_fill_rects(uint8_t* mask, const RectProb* rects, int nrect, cudaStream_t s);
int launch_conv_ffma(const ConvArgs& a, int nprob, int max_pix, cudaStream_t s);
int launch_stem_conv(const float* x, const float* w, const float* bias, __half* out_hi, __half* out_lo, int N, int H, int W, int K,
                     int stride, cudaStream_t s);
int launch_bilinear(const __half* in_hi, const __half* in_lo, int in_ps, __half* out_hi, __half* out_lo, int out_ps, int C,
                    const ResizeProb* probs, int nprob, int max_pix, cudaStream_t s);
// same resize, row-tiled (one CTA per two output rows of a problem): max_rows = the largest Hout (+ 2 for framed problems)
int launch_bilinear_rows(const __half* in_hi, const __half* in_lo, int in_ps, __half* out_hi, __half* out_lo, int out_ps, int C,
                         const ResizeProb* probs, int nprob, int max_rows, cudaStream_t s);
// plain copies of rectangles between NHWC tensors (forward_seg: feature crops into the atlases): Hout x Wout pixels x C channels each
int launch_copy_rects(const __half* in_hi, const __half* in_lo, int in_ps, __half* out_hi, __half* out_lo, int out_ps, int C,
                      const ResizeProb* probs, int nprob, int max_h, cudaStream_t s);
int launch_bilinear2x(const __half* in_hi, const __half* in_lo, __half* out_hi, __half* out_lo, int N, int Hin, int Win, int C, cudaStream_t s);
int launch_maxpool3x3s2(const __half* in_hi, const __half* in_lo, __half* out_hi, __half* out_lo, int N, int Hin, int Win,
                        int C, cudaStream_t s);
// NHWC split fp16 -> fp32 NCHW and back
int launch_export_nchw(const __half* in_hi, const __half* in_lo, float* out, int N, int HW, int C, cudaStream_t s);
int launch_import_nchw(const float* in, __half* out_hi, __half* out_lo, int N, int HW, int C, cudaStream_t s);

// tensor-core path (tc_conv.cu)
struct TcConvDesc;   // opaque per-op state (tensor maps)
bool tc_available();

}  // namespace kg
