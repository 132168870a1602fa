// Ground-truth encoder of the reference on the device (preprocessing.py:45-118, dataset_base.py:99-102): instance keypoints
// [n, 5, (x, y)] of one image -> the 55-channel target tensor gt = concat(kp heat [5], short offsets [10], mid offsets [40]).
// The reference builds num_insts x num_kps full-image distance maps in NumPy (4.4 s per 512x512 image); here one
// thread owns one (pixel, keypoint type) and scans the image's instances once.
//   disc mask (:62-77)      nearest instance (first minimum of the fp64 Euclidean distance, np.argmin) and distance <= KP_RADIUS
//   kp heat (:79-85)        1 inside any disc
//   mid offsets (:88-103)   inside the disc of keypoint `from`: (x, y) of keypoint `to` of the SAME instance minus the pixel position
//   short offsets (:45-60)  copy_with_border_check (:12-43) pastes the WHOLE (2R+1)^2 window around int(centre), instance after
//                           instance: inside the radius-R circle the integer offset int(centre) - pixel, zero in the window's
//                           corners; the disc-mask line of the reference (`temp_map[np.where(mask)==0,:] = 0.`) compares a tuple
//                           with 0 and is a no-op, so the LAST instance whose window covers a pixel wins -- reproduced as is.
#include "common.cuh"

namespace kg {

constexpr int GT_R = 5;      // config.KP_RADIUS
__constant__ int c_gt_mid_index[5][5] = {{-1, 0, 1, 2, 3}, {10, -1, 4, 5, 6}, {11, 14, -1, 7, 8}, {12, 15, 17, -1, 9}, {13, 16, 18, 19, -1}};

__global__ void __launch_bounds__(256) gt_encode_kernel(const float* __restrict__ boxes, const int* __restrict__ box_off, int H, int W,
                                                        float* __restrict__ gt) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int hw = H * W;
  if (p >= hw) return;
  const int y = p / W, x = p - y * W;
  const int j0 = box_off[b], j1 = box_off[b + 1];
  double best = 0.;
  int owner = -1;
  float sx = 0.f, sy = 0.f;
  for (int j = j0; j < j1; ++j) {
    const float cxf = __ldg(boxes + ((size_t)j * 5 + i) * 2), cyf = __ldg(boxes + ((size_t)j * 5 + i) * 2 + 1);
    const double dx = (double)cxf - (double)x, dy = (double)cyf - (double)y;           // float32 - int64 -> float64 (:70)
    const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    if (owner < 0 || d < best) { best = d; owner = j; }                                // np.argmin: first minimum
    const int cxi = (int)cxf, cyi = (int)cyf;                                          // int(center[k]) (:22)
    const int wy1 = max(cyi - GT_R, 0), wy2 = min(cyi + GT_R, H - 1) + 1;
    const int wx1 = max(cxi - GT_R, 0), wx2 = min(cxi + GT_R, W - 1) + 1;
    if (y >= wy1 && y < wy2 && x >= wx1 && x < wx2) {
      const int ox = cxi - x, oy = cyi - y;
      const bool in = ox * ox + oy * oy <= GT_R * GT_R;                                // sqrt(x*x + y*y) <= KP_RADIUS on integers (:53)
      sx = in ? (float)ox : 0.f; sy = in ? (float)oy : 0.f;
    }
  }
  const bool inside = owner >= 0 && best <= (double)GT_R;
  float* g = gt + (size_t)b * 55 * hw + p;
  g[(size_t)i * hw] = inside ? 1.f : 0.f;
  g[(size_t)(5 + 2 * i) * hw] = sx;
  g[(size_t)(5 + 2 * i + 1) * hw] = sy;
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    if (t == i) continue;
    const int m = c_gt_mid_index[i][t];
    float mx = 0.f, my = 0.f;
    if (inside) {
      mx = (float)((double)__ldg(boxes + ((size_t)owner * 5 + t) * 2) - (double)x);
      my = (float)((double)__ldg(boxes + ((size_t)owner * 5 + t) * 2 + 1) - (double)y);
    }
    g[(size_t)(15 + 2 * m) * hw] = mx;
    g[(size_t)(15 + 2 * m + 1) * hw] = my;
  }
}

}  // namespace kg

using namespace kg;

extern "C" int kg_encode_ground_truth(const float* d_boxes, const int* d_box_offsets, int B, int H, int W, float* d_gt, void* stream) {
  KG_REQUIRE(d_box_offsets && d_gt && B > 0 && H > 0 && W > 0, "kg_encode_ground_truth: bad arguments");
  KG_REQUIRE(B <= 65535 && (long long)H * W < (1ll << 31) / 55, "kg_encode_ground_truth: problem too large");
  dim3 grid((unsigned)((H * W + 255) / 256), 5, (unsigned)B);
  gt_encode_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_boxes, d_box_offsets, H, W, d_gt);
  KG_CUDA_CHECK(cudaGetLastError());
  return KG_OK;
}
