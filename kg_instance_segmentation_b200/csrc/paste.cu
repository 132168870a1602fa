// InstanceHeat.post_processing of the reference (test.py:127-157) on the device: every mask patch is resized to its
// rounded detection box, pasted into an input-sized canvas, the canvas is resized to the original image size and
// thresholded.  One thread per output pixel of one detection; nothing is materialised in between (the canvas and the
// resized patch are evaluated on the fly), so the per-box D2H + 2 x cv2.resize loop of the reference becomes one launch
// and the result leaves the device as one uint8 tensor.
//
// cv2.resize(..., INTER_LINEAR) on float32 (OpenCV's generic path, imgproc/resize.cpp): for destination index d,
//   f = (float)((d + 0.5) * (src / dst) - 0.5);  s = floor(f);  f -= s;  s < 0 -> (s, f) = (0, 0);  s >= src - 1 -> (src - 1, 0)
//   horizontal pass first: row[d] = S[s] * (1 - f) + S[s + 1] * f, then vertical: out = row0 * (1 - g) + row1 * g  (fp32, no FMA);
//   equal sizes are a plain copy.  (OpenCV builds with IPP differ from this in the last ulp; the mask is thresholded.)
#include "common.cuh"

namespace kg {

struct AxisTap { int s0, s1; float w0, w1; };

__device__ __forceinline__ AxisTap linear_tap(int d, int dst, int src) {
  AxisTap t;
  if (dst == src) { t.s0 = t.s1 = d; t.w0 = 1.f; t.w1 = 0.f; return t; }
  const double scale = (double)src / (double)dst;
  float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);     // no FMA contraction: the host computes mul, then sub
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { s = 0; f = 0.f; }
  if (s >= src - 1) { s = src - 1; f = 0.f; }
  t.s0 = s; t.s1 = min(s + 1, src - 1);
  t.w0 = __fsub_rn(1.f, f); t.w1 = f;
  return t;
}

__device__ __forceinline__ float lerp_cv(float a, float wa, float b, float wb) {
  return __fadd_rn(__fmul_rn(a, wa), __fmul_rn(b, wb));
}

struct PasteBox { const float* patch; int pitch, ph, pw; int y1, x1, th, tw; };

// value of the resized patch at (i, j) of the th x tw target rectangle
__device__ __forceinline__ float patch_at(const PasteBox& b, int i, int j) {
  if (b.th == b.ph && b.tw == b.pw) return __ldg(b.patch + (long long)i * b.pitch + j);
  const AxisTap ty = linear_tap(i, b.th, b.ph), tx = linear_tap(j, b.tw, b.pw);
  const float* r0 = b.patch + (long long)ty.s0 * b.pitch;
  const float* r1 = b.patch + (long long)ty.s1 * b.pitch;
  const float h0 = lerp_cv(__ldg(r0 + tx.s0), tx.w0, __ldg(r0 + tx.s1), tx.w1);
  const float h1 = lerp_cv(__ldg(r1 + tx.s0), tx.w0, __ldg(r1 + tx.s1), tx.w1);
  return lerp_cv(h0, ty.w0, h1, ty.w1);
}

__device__ __forceinline__ float canvas_at(const PasteBox& b, int y, int x) {
  const int i = y - b.y1, j = x - b.x1;
  if (i < 0 || j < 0 || i >= b.th || j >= b.tw) return 0.f;
  return patch_at(b, i, j);
}

__global__ void __launch_bounds__(256) paste_masks_kernel(const float* __restrict__ masks, const long long* __restrict__ mask_off,
                                                          const int* __restrict__ mask_pitch, const int* __restrict__ mask_hw,
                                                          const float* __restrict__ dets, int input_h, int input_w, int image_h,
                                                          int image_w, float seg_thresh, uint8_t* __restrict__ out_masks,
                                                          float* __restrict__ out_dets) {
  const int k = blockIdx.y;
  const float* d = dets + (size_t)k * 5;
  PasteBox b;
  // test.py:137-141: np.round on float32 = half-to-even, clipped to the network input
  b.y1 = max(0, (int)rintf(d[0])); b.x1 = max(0, (int)rintf(d[1]));
  const int y2 = min((int)rintf(d[2]), input_h - 1), x2 = min((int)rintf(d[3]), input_w - 1);
  b.th = y2 - b.y1; b.tw = x2 - b.x1;
  b.patch = masks + mask_off[k]; b.pitch = mask_pitch[k]; b.ph = mask_hw[2 * k]; b.pw = mask_hw[2 * k + 1];
  if (out_dets != nullptr && blockIdx.x == 0 && threadIdx.x < 5) {
    const int t = threadIdx.x;
    float v = d[4];
    if (t == 0) v = (float)((double)b.y1 / (double)input_h * (double)image_h);     // test.py:150-153 (Python floats)
    if (t == 1) v = (float)((double)b.x1 / (double)input_w * (double)image_w);
    if (t == 2) v = (float)((double)y2 / (double)input_h * (double)image_h);
    if (t == 3) v = (float)((double)x2 / (double)input_w * (double)image_w);
    out_dets[(size_t)k * 5 + t] = v;
  }
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= image_h * image_w) return;
  const int Y = p / image_w, X = p - Y * image_w;
  float v = 0.f;
  if (b.th > 0 && b.tw > 0) {
    if (image_h == input_h && image_w == input_w) {
      v = canvas_at(b, Y, X);
    } else {
      const AxisTap ty = linear_tap(Y, image_h, input_h), tx = linear_tap(X, image_w, input_w);
      // the canvas is zero outside the pasted rectangle: skip the patch reads when the 2 x 2 footprint misses it
      if (ty.s1 >= b.y1 && ty.s0 < b.y1 + b.th && tx.s1 >= b.x1 && tx.s0 < b.x1 + b.tw) {
        const float h0 = lerp_cv(canvas_at(b, ty.s0, tx.s0), tx.w0, canvas_at(b, ty.s0, tx.s1), tx.w1);
        const float h1 = lerp_cv(canvas_at(b, ty.s1, tx.s0), tx.w0, canvas_at(b, ty.s1, tx.s1), tx.w1);
        v = lerp_cv(h0, ty.w0, h1, ty.w1);
      }
    }
  }
  out_masks[(size_t)k * image_h * image_w + p] = v >= seg_thresh ? 1 : 0;      // np.where(mask >= seg_thresh, 1, 0) (test.py:148)
}

// test.py:92: torch.FloatTensor(HWC uint8 -> CHW) / 255 - 0.5, for a batch of already resized images: uint8 NHWC -> fp32 NCHW.
// One thread per pixel (3 bytes in, 3 coalesced fp32 stores out); IEEE division like torch's.
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ img, float* __restrict__ x, int HW, long long total) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const long long n = p / HW;
  const int q = (int)(p - n * HW);
  const uint8_t* s = img + p * 3;
  float* d = x + n * 3 * HW + q;
#pragma unroll
  for (int c = 0; c < 3; ++c) d[(long long)c * HW] = __fsub_rn(__fdiv_rn((float)s[c], 255.f), 0.5f);
}

}  // namespace kg

using namespace kg;

extern "C" int kg_preprocess_u8(const uint8_t* d_img, int N, int H, int W, float* d_x, void* stream) {
  KG_REQUIRE(d_img && d_x && N > 0 && H > 0 && W > 0, "kg_preprocess_u8: bad arguments");
  const long long total = (long long)N * H * W;
  preprocess_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_img, d_x, H * W, total);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}

extern "C" int kg_paste_masks(const float* d_masks, const long long* d_mask_off, const int* d_mask_pitch, const int* d_mask_hw,
                              const float* d_dets, int n, int input_h, int input_w, int image_h, int image_w, float seg_thresh,
                              uint8_t* d_out_masks, float* d_out_dets, void* stream) {
  KG_REQUIRE(n >= 0 && input_h > 0 && input_w > 0 && image_h > 0 && image_w > 0, "kg_paste_masks: bad sizes");
  if (n == 0) return KG_OK;
  KG_REQUIRE(d_masks && d_mask_off && d_mask_pitch && d_mask_hw && d_dets && d_out_masks, "kg_paste_masks: null argument");
  KG_REQUIRE(n <= 65535 && (long long)image_h * image_w < (1ll << 31), "kg_paste_masks: problem too large");
  dim3 grid((unsigned)(((long long)image_h * image_w + 255) / 256), (unsigned)n);
  paste_masks_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_masks, d_mask_off, d_mask_pitch, d_mask_hw, d_dets, input_h, input_w,
                                                            image_h, image_w, seg_thresh, d_out_masks, d_out_dets);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}
