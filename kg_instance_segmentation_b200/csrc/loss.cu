// Losses of the reference on the device: the forward passes (the arithmetic of its validation loop, train.py:165-177: model.eval(),
// forward, loss_dec x 4 + loss_seg) and their gradients with respect to the PREDICTIONS (what `loss.backward()` of train.py:150
// hands to the network's own backward pass, which is not part of this library -- SURVEY.md 8f-1).
//   DetectionLossAll.forward (loss.py:12-49): BCE on the keypoint maps + gt-masked L1 / KP_RADIUS on the short and mid offsets,
//                                             total = kp + short + 0.25 * mid.  ONE pass over the 55 prediction / target channels.
//   SEG_loss.forward         (seg_loss.py:31-97): per matched (prediction, ground-truth object) pair the mean BCE between the mask
//                                             patch and the ground-truth mask cropped to the rounded box and resized to the patch
//                                             (cv2.INTER_NEAREST); the IoU matching itself is list logic on the host.
#include "common.cuh"

#include <algorithm>

namespace kg {

__constant__ int c_loss_from_kp[20] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3, 1, 2, 3, 4, 2, 3, 4, 3, 4, 4};   // source keypoint of directed edge m (config.EDGES + reversed)

// F.binary_cross_entropy element: -(t * log(p) + (1 - t) * log(1 - p)) with both logs clamped at -100 (PyTorch), fp32 like torch
__device__ __forceinline__ float bce_term(float p, float t) {
  const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
  return -(t * lp + (1.f - t) * lq);
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.;
  if (warp == 0) {
    t = lane < (int)(blockDim.x >> 5) ? s_red[lane] : 0.;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;      // valid in thread 0
}

// acc[0] BCE sum, [1] short L1 sum, [2] short mask sum, [3] mid L1 sum, [4] mid mask sum
__global__ void __launch_bounds__(256) detection_loss_kernel(const float* __restrict__ pr_kp, const float* __restrict__ pr_short,
                                                             const float* __restrict__ pr_mid, const float* __restrict__ gt, long long total,
                                                             int HW, float inv_radius, double* __restrict__ acc) {
  __shared__ double s_red[8];
  double a_bce = 0., a_sh = 0., a_shm = 0., a_mid = 0., a_midm = 0.;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / HW;
    const int q = (int)(p - n * HW);
    const float* g = gt + n * 55 * HW + q;
    const float* pk = pr_kp + n * 5 * HW + q;
    const float* ps = pr_short + n * 10 * HW + q;
    const float* pm = pr_mid + n * 40 * HW + q;
    float gk[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      gk[i] = __ldg(g + (size_t)i * HW);
      a_bce += (double)bce_term(__ldg(pk + (size_t)i * HW), gk[i]);
      const float l0 = fabsf(__ldg(ps + (size_t)(2 * i) * HW) - __ldg(g + (size_t)(5 + 2 * i) * HW)) * inv_radius;
      const float l1 = fabsf(__ldg(ps + (size_t)(2 * i + 1) * HW) - __ldg(g + (size_t)(6 + 2 * i) * HW)) * inv_radius;
      a_sh += (double)(l0 * gk[i]) + (double)(l1 * gk[i]);
      a_shm += 2. * (double)gk[i];
    }
#pragma unroll
    for (int m = 0; m < 20; ++m) {
      const float w = gk[c_loss_from_kp[m]];
      const float l0 = fabsf(__ldg(pm + (size_t)(2 * m) * HW) - __ldg(g + (size_t)(15 + 2 * m) * HW)) * inv_radius;
      const float l1 = fabsf(__ldg(pm + (size_t)(2 * m + 1) * HW) - __ldg(g + (size_t)(16 + 2 * m) * HW)) * inv_radius;
      a_mid += (double)(l0 * w) + (double)(l1 * w);
      a_midm += 2. * (double)w;
    }
  }
  double v;
  v = block_sum(a_bce, s_red); if (threadIdx.x == 0) atomicAdd(acc + 0, v);
  v = block_sum(a_sh, s_red); if (threadIdx.x == 0) atomicAdd(acc + 1, v);
  v = block_sum(a_shm, s_red); if (threadIdx.x == 0) atomicAdd(acc + 2, v);
  v = block_sum(a_mid, s_red); if (threadIdx.x == 0) atomicAdd(acc + 3, v);
  v = block_sum(a_midm, s_red); if (threadIdx.x == 0) atomicAdd(acc + 4, v);
}

// out[0] kp loss, [1] short, [2] mid, [3] total = kp + short + 0.25 * mid (loss.py:44-48)
__global__ void detection_loss_finish_kernel(const double* __restrict__ acc, double n_kp_elems, float* __restrict__ out) {
  const float kp = (float)(acc[0] / n_kp_elems);
  const float sh = (float)(acc[1] / (acc[2] + 1e-10));
  const float mid = (float)(acc[3] / (acc[4] + 1e-10));
  out[0] = kp; out[1] = sh; out[2] = mid; out[3] = kp + sh + 0.25f * mid;
}

// d loss / d predictions of DetectionLossAll (autograd of loss.py:12-49): BCE -> g * (p - t) / max(p * (1 - p), 1e-12) / numel
// (PyTorch's binary_cross_entropy_backward), masked L1 -> g * sign(p - t) * mask / radius / (sum(mask) + 1e-10), mid term x 0.25.
// acc = the 5 sums of the forward pass (the mask sums are its denominators); g = upstream gradient (device scalar, or 1).
__global__ void __launch_bounds__(256) detection_loss_backward_kernel(const float* __restrict__ pr_kp, const float* __restrict__ pr_short,
                                                                      const float* __restrict__ pr_mid, const float* __restrict__ gt,
                                                                      long long total, int HW, float inv_radius, const double* __restrict__ acc,
                                                                      const float* __restrict__ grad_out, float* __restrict__ g_kp,
                                                                      float* __restrict__ g_short, float* __restrict__ g_mid) {
  const float g = grad_out != nullptr ? __ldg(grad_out) : 1.f;
  const float c_kp = g / (float)((double)total * 5.);
  const float c_sh = g * inv_radius / (float)(acc[2] + 1e-10);
  const float c_mid = 0.25f * g * inv_radius / (float)(acc[4] + 1e-10);
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / HW;
    const int q = (int)(p - n * HW);
    const float* gp = gt + n * 55 * HW + q;
    const long long ok = n * 5 * HW + q, os = n * 10 * HW + q, om = n * 40 * HW + q;
    float gk[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      gk[i] = __ldg(gp + (size_t)i * HW);
      const float pv = __ldg(pr_kp + ok + (size_t)i * HW);
      g_kp[ok + (size_t)i * HW] = c_kp * (pv - gk[i]) / fmaxf((1.f - pv) * pv, 1e-12f);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float d = __ldg(pr_short + os + (size_t)(2 * i + j) * HW) - __ldg(gp + (size_t)(5 + 2 * i + j) * HW);
        g_short[os + (size_t)(2 * i + j) * HW] = (d > 0.f ? c_sh : d < 0.f ? -c_sh : 0.f) * gk[i];
      }
    }
#pragma unroll
    for (int m = 0; m < 20; ++m) {
      const float w = gk[c_loss_from_kp[m]];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float d = __ldg(pr_mid + om + (size_t)(2 * m + j) * HW) - __ldg(gp + (size_t)(15 + 2 * m + j) * HW);
        g_mid[om + (size_t)(2 * m + j) * HW] = (d > 0.f ? c_mid : d < 0.f ? -c_mid : 0.f) * w;
      }
    }
  }
}

struct SegPair { long long patch_off; int pitch, h, w; int gt_index; int y1, x1, y2, x2; };

// one CTA per matched pair: mean BCE(patch, nearest-resized crop of the ground-truth mask) (seg_loss.py:62-86)
__global__ void __launch_bounds__(256) seg_loss_kernel(const float* __restrict__ masks, const SegPair* __restrict__ pairs,
                                                       const float* __restrict__ gt_masks, int H, int W, float* __restrict__ pair_loss) {
  __shared__ double s_red[8];
  const SegPair pr = pairs[blockIdx.x];
  const int ch = pr.y2 - pr.y1, cw = pr.x2 - pr.x1;
  double a = 0.;
  if (ch > 0 && cw > 0) {
    // cv2.resize(..., (w1, h1), INTER_NEAREST): src index = min(floor(dst * (1 / (dst_size / src_size))), src_size - 1)
    const double ify = 1.0 / ((double)pr.h / (double)ch), ifx = 1.0 / ((double)pr.w / (double)cw);
    const float* g = gt_masks + (size_t)pr.gt_index * H * W;
    for (int e = threadIdx.x; e < pr.h * pr.w; e += blockDim.x) {
      const int r = e / pr.w, c = e - r * pr.w;
      const int sy = min((int)floor(__dmul_rn((double)r, ify)), ch - 1), sx = min((int)floor(__dmul_rn((double)c, ifx)), cw - 1);
      const float t = __ldg(g + (size_t)(pr.y1 + sy) * W + pr.x1 + sx);
      a += (double)bce_term(__ldg(masks + pr.patch_off + (long long)r * pr.pitch + c), t);
    }
  }
  const double v = block_sum(a, s_red);
  if (threadIdx.x == 0) pair_loss[blockIdx.x] = (float)(v / (double)max(1, pr.h * pr.w));
}

// d (sum_k coeff[k] * pair_loss[k]) / d masks: one CTA per pair, scatter-add (a patch can be matched with several objects)
__global__ void __launch_bounds__(256) seg_loss_backward_kernel(const float* __restrict__ masks, const SegPair* __restrict__ pairs,
                                                                const float* __restrict__ gt_masks, int H, int W,
                                                                const float* __restrict__ coeff, float* __restrict__ grad_masks) {
  const SegPair pr = pairs[blockIdx.x];
  const int ch = pr.y2 - pr.y1, cw = pr.x2 - pr.x1;
  if (ch <= 0 || cw <= 0) return;
  const float c = __ldg(coeff + blockIdx.x) / (float)(pr.h * pr.w);
  const double ify = 1.0 / ((double)pr.h / (double)ch), ifx = 1.0 / ((double)pr.w / (double)cw);
  const float* g = gt_masks + (size_t)pr.gt_index * H * W;
  for (int e = threadIdx.x; e < pr.h * pr.w; e += blockDim.x) {
    const int r = e / pr.w, cc = e - r * pr.w;
    const int sy = min((int)floor(__dmul_rn((double)r, ify)), ch - 1), sx = min((int)floor(__dmul_rn((double)cc, ifx)), cw - 1);
    const float t = __ldg(g + (size_t)(pr.y1 + sy) * W + pr.x1 + sx);
    const long long o = pr.patch_off + (long long)r * pr.pitch + cc;
    const float pv = __ldg(masks + o);
    atomicAdd(grad_masks + o, c * (pv - t) / fmaxf((1.f - pv) * pv, 1e-12f));
  }
}

}  // namespace kg

using namespace kg;

extern "C" int kg_detection_loss(const float* d_pr_kp, const float* d_pr_short, const float* d_pr_mid, const float* d_gt, int N, int H, int W,
                                 float kp_radius, double* d_scratch5, float* d_out4, void* stream_) {
  KG_REQUIRE(d_pr_kp && d_pr_short && d_pr_mid && d_gt && d_scratch5 && d_out4 && N > 0 && H > 0 && W > 0 && kp_radius > 0.f,
             "kg_detection_loss: bad arguments");
  cudaStream_t stream = (cudaStream_t)stream_;
  KG_CUDA_CHECK(cudaMemsetAsync(d_scratch5, 0, 5 * sizeof(double), stream));
  const long long total = (long long)N * H * W;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 8);
  detection_loss_kernel<<<grid, 256, 0, stream>>>(d_pr_kp, d_pr_short, d_pr_mid, d_gt, total, H * W, 1.f / kp_radius, d_scratch5);
  detection_loss_finish_kernel<<<1, 1, 0, stream>>>(d_scratch5, (double)total * 5., d_out4);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

extern "C" int kg_seg_loss_pairs(const float* d_masks, const void* d_pairs, int n_pairs, const float* d_gt_masks, int H, int W,
                                 float* d_pair_loss, void* stream) {
  KG_REQUIRE(n_pairs >= 0 && H > 0 && W > 0, "kg_seg_loss_pairs: bad sizes");
  if (n_pairs == 0) return KG_OK;
  KG_REQUIRE(d_masks && d_pairs && d_gt_masks && d_pair_loss, "kg_seg_loss_pairs: null argument");
  seg_loss_kernel<<<(unsigned)n_pairs, 256, 0, (cudaStream_t)stream>>>(d_masks, reinterpret_cast<const SegPair*>(d_pairs), d_gt_masks, H, W,
                                                                      d_pair_loss);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

extern "C" int kg_detection_loss_backward(const float* d_pr_kp, const float* d_pr_short, const float* d_pr_mid, const float* d_gt, int N,
                                          int H, int W, float kp_radius, const double* d_scratch5, const float* d_grad_out,
                                          float* d_grad_kp, float* d_grad_short, float* d_grad_mid, void* stream) {
  KG_REQUIRE(d_pr_kp && d_pr_short && d_pr_mid && d_gt && d_scratch5 && d_grad_kp && d_grad_short && d_grad_mid && N > 0 && H > 0 &&
             W > 0 && kp_radius > 0.f, "kg_detection_loss_backward: bad arguments");
  const long long total = (long long)N * H * W;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 8);
  detection_loss_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_pr_kp, d_pr_short, d_pr_mid, d_gt, total, H * W, 1.f / kp_radius,
                                                                         d_scratch5, d_grad_out, d_grad_kp, d_grad_short, d_grad_mid);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

extern "C" int kg_seg_loss_pairs_backward(const float* d_masks, const void* d_pairs, int n_pairs, const float* d_gt_masks, int H, int W,
                                          const float* d_pair_coeff, float* d_grad_masks, void* stream) {
  KG_REQUIRE(n_pairs >= 0 && H > 0 && W > 0, "kg_seg_loss_pairs_backward: bad sizes");
  if (n_pairs == 0) return KG_OK;
  KG_REQUIRE(d_masks && d_pairs && d_gt_masks && d_pair_coeff && d_grad_masks, "kg_seg_loss_pairs_backward: null argument");
  seg_loss_backward_kernel<<<(unsigned)n_pairs, 256, 0, (cudaStream_t)stream>>>(d_masks, reinterpret_cast<const SegPair*>(d_pairs), d_gt_masks,
                                                                               H, W, d_pair_coeff, d_grad_masks);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}
