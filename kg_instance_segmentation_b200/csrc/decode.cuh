// Device-side decode path: Hough vote -> Gaussian blur -> peaks -> sort -> greedy grouping ->
// refine/box assembly -> NMS.  Replaces postprocessing.py:8-261 and nms.py:4-53 of the reference.
#pragma once
#include "common.cuh"
#include "../../include/kgnet_b200.h"

namespace kg {

size_t decode_workspace_bytes(const kg_decode_config* cfg, const kg_decode_scale* scales);
int decode_launch(const kg_decode_config* cfg, const kg_decode_scale* scales, const kg_decode_outputs* out,
                  void* workspace, size_t workspace_bytes, cudaStream_t stream, int* n_launches);
int skeletons_to_boxes_host(const double* h_skel, int n, int box_scale, int apply_refine, uint8_t* h_keep,
                            double* h_boxes, int* n_boxes);
int nms_host(const double* h_boxes, int n, double nms_thresh, double* h_out, int* n_out);

}  // namespace kg
