// Device-side decode path: Hough vote -> Gaussian blur -> peaks -> sort -> greedy grouping ->
// refine/box assembly -> NMS.  Replaces postprocessing.py:8-261 and nms.py:4-53 of the reference.
#pragma once
#include "common.cuh"

namespace kg {

constexpr int KG_NUM_KPS = 5;
constexpr int KG_MAX_SCALES = 4;

struct DecodeScale {
  const float* kp;    // [N,5,H,W]  fp32 NCHW (sigmoid applied)
  const float* sh;    // [N,10,H,W] short offsets, channel 2i = dx, 2i+1 = dy
  const float* mid;   // [N,40,H,W] mid offsets, edge m -> channels (2m, 2m+1) = (dx, dy)
  int H, W;
  int box_scale;      // 1,2,4,8 (postprocessing.py:256-259)
};

struct DecodeConfig {
  int N;
  int n_scales;
  int max_peaks;      // per (image, scale); <= 4096
  int max_boxes;      // per image, pre-NMS; <= 2048
  double nms_thresh;
};

// Caller-visible outputs (device pointers).  Any of the optional ones may be null.
struct DecodeOutputs {
  double* dets;        // [N, max_boxes, 5]  post-NMS (y1,x1,y2,x2,conf) in keep order
  int* det_count;      // [N]
  double* boxes;       // optional [N, max_boxes, 5]  pre-NMS boxes (gather_skeleton order)
  int* box_count;      // optional [N]
  double* skeletons;   // optional [N, n_scales, max_peaks, 5, 3] (x,y,conf), ALL skeletons (pre-refine)
  int* skel_count;     // optional [N, n_scales]
  double* peak_conf;   // optional [N, n_scales, max_peaks]  sorted (conf desc, generation order asc)
  int* peak_key;       // optional [N, n_scales, max_peaks]  id*H*W + y*W + x
  int* peak_count;     // optional [N, n_scales]
  double* heat;        // optional per-scale blurred-heatmap dump is not kept; this is the VOTED heat
                       // (pre-blur) of scale 0.. concatenated [sum_s N*5*H_s*W_s]; null to skip export
  int* status;         // [1] bit0 peaks overflow, bit1 boxes overflow
};

size_t decode_workspace_bytes(const DecodeConfig& cfg, const DecodeScale* scales);
int decode_launch(const DecodeConfig& cfg, const DecodeScale* scales, const DecodeOutputs& out,
                  void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace kg
