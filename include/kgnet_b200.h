/*
 * kgnet_b200 — C-ABI of the B200-native KGnet inference hot path.
 *
 * The reference (yijingru/KG_Instance_Segmentation) has no FFI layer: its boundary is a Python call
 * surface.  Every entry point below names the reference function(s) (file:line, relative to the
 * reference checkout) whose arithmetic it replaces; the Python package
 * `kg_instance_segmentation_b200` binds them with ctypes and re-exposes the reference names
 * (see INTEGRATION.md).
 *
 * Conventions
 *   - plain C types only; `void* stream` is a cudaStream_t (NULL = legacy default stream);
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers;
 *   - every function returns 0 on success or a negative kg_status; kg_last_error() returns the
 *     message of the last failure on the calling thread;
 *   - device entry points never allocate and never synchronise: the caller passes a workspace of
 *     kg_*_workspace_bytes() bytes and owns all buffers.  The *_host convenience entry points
 *     allocate, copy H2D/D2H and synchronise internally.
 *   - all tensors are dense; activations at this boundary are fp32 NCHW exactly like the reference's
 *     torch tensors, detections are fp64 rows [y1, x1, y2, x2, conf] like the reference's NumPy arrays.
 */
#ifndef KGNET_B200_H_
#define KGNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KG_ABI_VERSION 2
#define KG_NUM_KPS 5          /* config.py:3  */
#define KG_NUM_EDGES 10       /* config.py:2  */
#define KG_MAX_SCALES 4       /* test.py:105-108: c0..c3 */

typedef enum kg_status {
  KG_OK = 0,
  KG_ERR_INVALID = -1,   /* bad argument / unsupported shape                               */
  KG_ERR_CUDA = -2,      /* CUDA runtime / driver failure                                   */
  KG_ERR_CAPACITY = -3,  /* a bounded device list (peaks / boxes) overflowed: raise the cap */
  KG_ERR_STATE = -4,     /* call order violated                                             */
  KG_ERR_WORKSPACE = -5  /* caller-provided workspace too small                             */
} kg_status;

const char* kg_last_error(void);
int kg_abi_version(void);
/* Compute capability of the current device as major*10+minor (100 on B200); <0 on error. */
int kg_device_arch(void);

/* Optional per-stage device timing (CUDA events recorded on the launching stream around each kernel
 * class).  Stage ids: 0 vote, 1 blur+peak, 2 sort+group+boxes, 3 nms, >=8 network stages.
 * kg_timing_collect synchronises the device and returns the accumulated milliseconds and launch counts
 * since the last collect.  No reference equivalent (the reference only prints time.time(), test.py:96-98). */
int kg_timing_enable(int on);
int kg_timing_collect(float* ms_per_stage, int* launches_per_stage, int n_stages);

/* ------------------------------------------------------------------------------------------------
 * Decode path: Hough vote -> Gaussian blur -> peaks -> conf-sorted greedy keypoint-graph grouping
 * -> refine -> boxes -> NMS.
 * Replaces postprocessing.py:8-64 (compute_heatmaps / accumulate_votes / get_keypoints),
 * :80-147 (group_skeletons / get_skeletons_and_masks), :150-261 (refine_skeleton / skeleton_to_box /
 * gather_skeleton) and nms.py:4-53, batched over N images (the reference handles batch element 0 only).
 * ---------------------------------------------------------------------------------------------- */

typedef struct kg_decode_scale {
  const float* d_kp;     /* [N, 5, H, W]  keypoint heatmaps (sigmoid applied), KGnet.py:300-314           */
  const float* d_short;  /* [N,10, H, W]  short offsets, channel 2i = dx, 2i+1 = dy (postprocessing.py:49) */
  const float* d_mid;    /* [N,40, H, W]  mid offsets, directed edge m -> channels (2m, 2m+1) (:108-112)   */
  int H, W;
  int box_scale;         /* 1, 2, 4, 8 (postprocessing.py:256-259)                                         */
} kg_decode_scale;

typedef struct kg_decode_config {
  int N;                 /* images in the batch                                      */
  int n_scales;          /* 1..KG_MAX_SCALES                                         */
  int max_peaks;         /* cap per (image, scale): power of two, 64..8192           */
  int max_boxes;         /* cap per image over all scales (pre-NMS): power of two, 64..8192 */
  double nms_thresh;     /* nms.py:4 (default 0.5)                                   */
  double peak_thresh;    /* postprocessing.py:145 (0.004)                            */
} kg_decode_config;

/* Device outputs; any pointer other than d_status may be NULL to skip that export. */
typedef struct kg_decode_outputs {
  double* d_dets;        /* [N, max_boxes, 5]  post-NMS rows in keep order (nms.py:51)                      */
  int* d_det_count;      /* [N]                                                                               */
  double* d_boxes;       /* [N, max_boxes, 5]  pre-NMS boxes in gather_skeleton order (postprocessing.py:255)  */
  int* d_box_count;      /* [N]                                                                               */
  double* d_skeletons;   /* [N, n_scales, max_peaks, 5, 3]  (x, y, conf); ALL skeletons of group_skeletons     */
  int* d_skel_count;     /* [N, n_scales]                                                                     */
  uint8_t* d_skel_keep;  /* [N, n_scales, max_peaks]  1 where refine_skeleton keeps the skeleton (:150-159)   */
  double* d_peak_conf;   /* [N, n_scales, max_peaks]  peaks in group_skeletons' sorted order (:87)            */
  int* d_peak_key;       /* [N, n_scales, max_peaks]  id*H*W + y*W + x                                        */
  int* d_peak_count;     /* [N, n_scales]                                                                     */
  double* d_heat[KG_MAX_SCALES];  /* per scale [N,5,H,W]: Hough heat AFTER the Gaussian blur (:143-144)       */
  double* d_vote[KG_MAX_SCALES];  /* per scale [N,5,H,W]: Hough heat BEFORE the blur (compute_heatmaps)        */
  int* d_status;         /* [1] bit0: a peak list overflowed, bit1: a box list overflowed, bit2: an image has
                            more detections than det_packed_k (d_det_packed truncated; d_dets is complete)      */
  double* d_det_packed;  /* optional [N, det_packed_k + 1, 5]: fixed-size per-image record for the data-parallel
                            all-gather (no reference equivalent): row 0 = (count, rows stored, 0, 0, 0), rows 1.. =
                            the first min(count, det_packed_k) detections in keep order                          */
  int det_packed_k;
} kg_decode_outputs;

size_t kg_decode_workspace_bytes(const kg_decode_config* cfg, const kg_decode_scale* scales);

/* Enqueue the whole decode on `stream`.  No sync, no allocation.  Number of kernel launches is
 * returned through *n_launches when non-NULL. */
int kg_decode(const kg_decode_config* cfg, const kg_decode_scale* scales, const kg_decode_outputs* out,
              void* d_workspace, size_t workspace_bytes, void* stream, int* n_launches);

/* Host-buffer end-to-end variant of the same path (the reference API's data flow,
 * postprocessing.py:134-136: head maps in host memory -> detections in host memory).
 * h_kp/h_short/h_mid[s] are [N,C,H_s,W_s] fp32 host arrays (pinned for full speed).
 * h_dets [N,max_boxes,5], h_det_count [N].  Copies, runs and synchronises on `stream`. */
int kg_decode_host(const kg_decode_config* cfg, const float* const* h_kp, const float* const* h_short,
                   const float* const* h_mid, const int* H, const int* W, const int* box_scale,
                   double* h_dets, int* h_det_count, void* stream);

/* refine_skeleton + skeleton_to_box for host lists (postprocessing.py:150-242): h_skeletons [n,5,3]
 * -> h_keep [n] (refine mask) and h_boxes [n,5] rows for the kept skeletons that yield a box, in
 * order; *n_boxes receives the count.  Does not mutate the input (the reference scales it in place). */
int kg_skeletons_to_boxes_host(const double* h_skeletons, int n, int box_scale, int apply_refine,
                               uint8_t* h_keep, double* h_boxes, int* n_boxes);

/* nms.py:4-53 on a host array [n,5]; h_out [n,5] receives the kept rows in keep order. */
int kg_nms_host(const double* h_boxes, int n, double nms_thresh, double* h_out, int* n_out);

/* ------------------------------------------------------------------------------------------------
 * Network: truncated ResNet backbone + top-down decoder + 12 two-layer 7x7 heads (forward_dec) and the per-box
 * mask branch (forward_seg).  Replaces KGnet.py:123-350 (ResNet.__init__ / forward_dec / forward_seg /
 * get_patches / mask_forward / CombinationModule) of the reference.
 * ---------------------------------------------------------------------------------------------- */
typedef struct kg_net kg_net;

/* blocks = {3,4,6} (resnet50, KGnet.py:377-386), {3,4,23} (resnet101) or {3,8,36} (resnet152); NULL = resnet50. */
int kg_net_create(kg_net** out, const int* blocks);
void kg_net_destroy(kg_net* net);

/* Register one nn.Conv2d by its state-dict prefix (e.g. "layer1.0.conv1", "c0_conv.0", "kp_head_c2.2").
 * h_w: HOST fp32 [Cout,Cin,R,S] (torch layout); h_bias: [Cout] or NULL; h_bn_*: the eval-mode BatchNorm2d that
 * follows it (weight, bias, running_mean, running_var) or NULL — folded into the conv at load time. */
int kg_net_set_conv(kg_net* net, const char* name, const float* h_w, int Cout, int Cin, int R, int S, const float* h_bias,
                    const float* h_bn_weight, const float* h_bn_bias, const float* h_bn_mean, const float* h_bn_var,
                    double bn_eps);
/* Checks that every layer of the architecture was set, fuses the first-layer head convs, repacks and uploads. */
int kg_net_finalize(kg_net* net);

/* precision: 0 = CUDA-core fp32 FFMA everywhere (on-device reference), 1 = "fast" (tcgen05; split-fp16 3-pass in
 * backbone/decoder, single-pass fp16 in the heads; meets 1e-3 on the keypoint heatmaps), 2 = "exact" (tcgen05,
 * split-fp16 3-pass everywhere). */
size_t kg_net_workspace_bytes(kg_net* net, int N, int H, int W, int precision);

/* ResNet.forward_dec (KGnet.py:275-318).  d_x: [N,3,H,W] fp32 NCHW.  d_heads[12]: kp0, short0, mid0, kp1, ... mid3,
 * each [N,{5,10,40},H/2^s,W/2^s] fp32 NCHW.  d_feats[5] (or NULL): c0..c4 fp32 NCHW.  H, W multiples of 16.
 * The workspace keeps c0..c4 in the internal layout for a following kg_net_forward_seg. */
int kg_net_forward_dec(kg_net* net, const float* d_x, int N, int H, int W, float* const* d_heads, float* const* d_feats,
                       int precision, void* d_workspace, size_t workspace_bytes, void* stream, int* n_launches);

/* The same pass fed by the camera image: d_img = uint8 NHWC [N,H,W,3] (cv2 BGR order), i.e. BEFORE the `x / 255 - 0.5` of test.py:92.
 * The normalisation is folded into the two stem convs (exactly: k - 128 is an fp16 integer, w / 255 and the constant go into weights
 * and bias), so neither kg_preprocess_u8 nor an fp32 copy of the input is needed.  Tensor-core precisions (1, 2) only. */
int kg_net_forward_dec_u8(kg_net* net, const unsigned char* d_img, int N, int H, int W, float* const* d_heads, float* const* d_feats,
                          int precision, void* d_workspace, size_t workspace_bytes, void* stream, int* n_launches);

/* Loads caller-provided fp32 NCHW features c0..c4 into the workspace (forward_seg on features that did not come
 * from kg_net_forward_dec, KGnet.py:321). */
int kg_net_import_feats(kg_net* net, const float* const* d_feats, int N, int H, int W, int precision, void* d_workspace,
                        size_t workspace_bytes, void* stream);

/* ResNet.forward_seg (KGnet.py:321-350), two calls.  prepare: host boxes ([sum(box_counts),5] fp64 rows
 * y1,x1,y2,x2,score, image-major) -> crop rectangles (get_patches, :246-256), grouped problem lists, required
 * scratch bytes and the mask layout: h_mask_index[box] = mask slot or -1 when the box is skipped (:341-342),
 * h_mask_hw[slot] = (h, w), h_mask_off[slot] = float offset of the patch's first element in d_masks,
 * h_mask_pitch[slot] = its row stride in floats (the tensor-core path packs all patches into one atlas image).
 * prepare is HOST planning: it may (re)allocate the pinned staging buffer of the problem lists and waits for the
 * staging copy of a previous kg_net_forward_seg; it launches nothing.
 * run: all boxes together — dense tcgen05 convs over the per-level atlases (precision 1, 2) or grouped CUDA-core
 * launches (precision 0).  Enqueue only: one cudaMemcpyAsync of the problem lists from pinned memory into the caller's
 * seg workspace, then kernels; no allocation, no synchronisation. */
int kg_net_seg_prepare(kg_net* net, int N, int H, int W, const int* box_counts, const double* h_boxes,
                       size_t* seg_workspace_bytes, long long* mask_floats, int* n_masks, int* h_mask_index,
                       int* h_mask_hw, long long* h_mask_off, int* h_mask_pitch);
int kg_net_forward_seg(kg_net* net, void* d_dec_workspace, void* d_seg_workspace, size_t seg_workspace_bytes,
                       float* d_masks, void* stream, int* n_launches);

/* One nn.Conv2d (+bias, +residual, +ReLU) on fp32 NCHW device tensors: operator-level entry for unit tests.
 * mode: 0 CUDA cores; 1 / 2 / 3 = tcgen05 implicit GEMM with 1 / 2 / 3 split-fp16 passes (2 = split activations x
 * single-plane weights); 11 / 12 / 13 = the row-GEMM + shift-add kernel with 1 / 2 / 3 passes; 21 = single pass with
 * hi-plane output (CTA-pair kernel where eligible).  Allocates and synchronises internally. */
int kg_conv2d_nchw(const float* d_x, int N, int Cin, int H, int W, const float* h_w, const float* h_bias, int Cout, int R,
                   int S, int stride, int pad, int relu, const float* d_res, int mode, float* d_y, void* stream);

/* The three second-layer head convs of one scale (KGnet.py:161-209: 7x7, Cin -> 5 (sigmoid) / 10 / 40) through the
 * row-GEMM + shift-add tcgen05 kernel, on an fp32 NCHW device input with 3*Cin channels (head h reads channels
 * [h*Cin, (h+1)*Cin)); h_w[h] = host [Cout_h, Cin, 7, 7], d_y[h] = device [N, Cout_h, H, W].  Unit-test entry. */
int kg_heads_l2_nchw(const float* d_x, int N, int Cin, int H, int W, const float* const* h_w, const float* const* h_bias,
                     float* const* d_y, void* stream);

/* Work census of the current forward_dec plan: out[0]/out[1] = algorithmic FLOPs (2*MACs, real channel counts, one
 * pass) of the tensor-core / CUDA-core convs, out[2]/out[3] = their launch counts, out[4..7] = tensor-core FLOPs
 * of the backbone, decoder, first-layer heads, second-layer heads.  n >= 8. */
int kg_net_plan_info(kg_net* net, double* out, int n);

/* ------------------------------------------------------------------------------------------------
 * Callers either side of the network.
 * ---------------------------------------------------------------------------------------------- */

/* test.py:92 for a batch of already-resized images: uint8 NHWC (cv2 BGR order) -> fp32 NCHW, x / 255 - 0.5.
 * Lets the input cross PCIe as 3 bytes per pixel instead of 12. */
int kg_preprocess_u8(const uint8_t* d_img, int N, int H, int W, float* d_x, void* stream);

/* InstanceHeat.post_processing (test.py:127-157): for each of the n mask patches (fp32, patch k at
 * d_masks + d_mask_off[k], row stride d_mask_pitch[k], size d_mask_hw[2k] x d_mask_hw[2k+1]) and its detection row
 * d_dets[5k..] = (y1, x1, y2, x2, conf) fp32: resize the patch to the rounded box (cv2.resize INTER_LINEAR), paste it
 * into an input_h x input_w canvas, resize the canvas to image_h x image_w and threshold at seg_thresh.
 * d_out_masks: [n, image_h, image_w] uint8 in {0, 1}; d_out_dets (or NULL): [n, 5] fp32 boxes in image coordinates. */
int kg_paste_masks(const float* d_masks, const long long* d_mask_off, const int* d_mask_pitch, const int* d_mask_hw,
                   const float* d_dets, int n, int input_h, int input_w, int image_h, int image_w, float seg_thresh,
                   uint8_t* d_out_masks, float* d_out_dets, void* stream);

/* preprocessing.get_ground_truth (preprocessing.py:105-118) + the channel concat of dataset_base.py:99-102 for a batch:
 * d_boxes [sum n_b, 5, 2] fp32 keypoints (x, y) in the order tl, tr, bl, br, centre (dataset_base.masks_to_bboxes),
 * image b owns rows d_box_offsets[b] .. d_box_offsets[b+1] (B + 1 ints).  d_gt: [B, 55, H, W] fp32 =
 * concat(kp heat [5], short offsets [10], mid offsets [40]) -- the gt_c* tensors DetectionLossAll consumes. */
int kg_encode_ground_truth(const float* d_boxes, const int* d_box_offsets, int B, int H, int W, float* d_gt, void* stream);

/* Loss forward passes (the reference's validation loop, train.py:165-177; no backward pass in this library).
 * DetectionLossAll.forward (loss.py:40-49) of one scale: predictions [N,5|10|40,H,W] fp32, target d_gt [N,55,H,W] fp32.
 * d_scratch5: 5 doubles of device scratch (cleared here); d_out4 = (kp BCE, short, mid, kp + short + 0.25 * mid) fp32. */
int kg_detection_loss(const float* d_pr_kp, const float* d_pr_short, const float* d_pr_mid, const float* d_gt, int N, int H, int W,
                      float kp_radius, double* d_scratch5, float* d_out4, void* stream);

/* SEG_loss.forward's per-object term (seg_loss.py:62-86) for n_pairs matched (prediction, ground-truth object) pairs:
 * mean BCE between mask patch and the ground-truth mask cropped to the rounded box and resized (INTER_NEAREST) to the patch.
 * d_pairs: n_pairs records { int64 patch_off; int32 pitch, h, w, gt_index, y1, x1, y2, x2; } (40 bytes, 8-byte aligned);
 * d_gt_masks: [n_gt, H, W] fp32; d_pair_loss: [n_pairs] fp32. */
int kg_seg_loss_pairs(const float* d_masks, const void* d_pairs, int n_pairs, const float* d_gt_masks, int H, int W,
                      float* d_pair_loss, void* stream);

/* Gradient of kg_detection_loss's total with respect to the three prediction tensors (what autograd computes for loss.py:12-49:
 * PyTorch's binary_cross_entropy backward on the keypoint maps, sign(p - t) * mask / radius / (sum(mask) + 1e-10) on the offsets).
 * d_scratch5: the 5 sums left by kg_detection_loss on the SAME inputs; d_grad_out: upstream gradient (one device float) or NULL = 1;
 * d_grad_*: fp32 tensors of the predictions' shapes (overwritten). */
int kg_detection_loss_backward(const float* d_pr_kp, const float* d_pr_short, const float* d_pr_mid, const float* d_gt, int N, int H, int W,
                               float kp_radius, const double* d_scratch5, const float* d_grad_out, float* d_grad_kp, float* d_grad_short,
                               float* d_grad_mid, void* stream);

/* Gradient of sum_k d_pair_coeff[k] * pair_loss[k] (kg_seg_loss_pairs) with respect to the mask buffer: ADDED into d_grad_masks
 * (same indexing as d_masks; the caller clears it), autograd of seg_loss.py:84-90. */
int kg_seg_loss_pairs_backward(const float* d_masks, const void* d_pairs, int n_pairs, const float* d_gt_masks, int H, int W,
                               const float* d_pair_coeff, float* d_grad_masks, void* stream);

/* torch.optim.Adam.step() of the reference's training loop (train.py:71,154; PyTorch defaults: no weight decay, no amsgrad) for ALL
 * parameter tensors in one launch.  d_tensors: records { float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
 * int64 numel; } (40 bytes); d_chunks: n_chunks records { int32 tensor; int32 pad; int64 start; } covering every tensor in pieces of
 * at most 65536 elements; step = the 1-based step count of these tensors (bias correction). */
int kg_adam_step(const void* d_tensors, const void* d_chunks, int n_chunks, double lr, double beta1, double beta2, double eps, int step,
                 void* stream);

/* Test hook (host only, no GPU needed): the first-fit placement by liveness that lays out the activation workspace of forward_dec.
 * Buffer b (bytes[b]) is first written by op def[b] and last read by op last[b] (last[b] >= n_ops: never released); a buffer may reuse
 * memory released by ops < def[b] only.  Writes the byte offsets and the arena size. */
int kg_debug_place_by_liveness(int n_ops, int n_buffers, const int* def, const int* last, const unsigned long long* bytes,
                               unsigned long long* offsets, unsigned long long* total);

/* 1 when the tcgen05/TMA path initialised on the current device; kg_tc_status() says why not otherwise. */
int kg_tc_available(void);
const char* kg_tc_status(void);

#ifdef __cplusplus
}
#endif
#endif /* KGNET_B200_H_ */
