/*
 * pf_b200.h -- C ABI of the B200-native bg-forecast hot path (libpf_b200.so).
 *
 * The reference (nianticlabs/panoptic-forecasting) has no native layer: the path below is
 * pure PyTorch + torch_scatter.  These entry points are what a binding on the reference
 * side calls instead (INTEGRATION.md shows the ctypes stubs).  Conventions:
 *   - every function returns 0 on success, a positive cudaError_t on a CUDA failure, or a
 *     negative PF_E* code on a bad argument; pf_last_error() gives a message.
 *   - `*_dev` pointers are device memory owned by the caller; the library never frees or
 *     reallocates them.  Work space is caller-provided (query `*_workspace_bytes`).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All work is
 *     enqueued asynchronously on it; nothing synchronises unless stated.
 *   - there is NO CPU fallback: without a CUDA device the compute entry points fail.
 */
#ifndef PF_B200_H_
#define PF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PF_EINVAL (-1)   /* bad argument (null pointer, non-positive size, unsupported shape) */
#define PF_ENOMEM (-2)   /* work space too small */
#define PF_ESTATE (-3)   /* handle not ready (weights missing) */

int pf_version(void);
const char* pf_last_error(void);

/* ---------------------------------------------------------------------------------------
 * Stage A: unproject -> rigid chain -> reproject -> 4-way splat -> nearest-depth select.
 * Replaces PCTransformModel.predict,
 *   reference panoptic_forecasting/models/pc_transform/pc_transform_model.py:26-150
 * including its torch_scatter.scatter_min call (:118-119) and the gather/index_put tail
 * (:120-139).
 *
 *   depth_dev   f32 [b,t,H,W]    metric depth of each source pixel
 *   mask_dev    u8  [b,t,H,W]    depth_mask (0/1)
 *   seg_dev     u8  [b,t,H,W,payload]  payload = 1 (label) or 3 (is_img RGB)
 *   K, Kinv     f32 [b,3,3]      intrinsics and torch.inverse(intrinsics)      (DEVICE memory)
 *   E, Einv     f32 [b,4,4]      extrinsics and torch.inverse(extrinsics)      (DEVICE memory)
 *   T           f32 [b,t,4,4]    target_T                                      (DEVICE memory)
 *   lut_dev     u8  [256] or NULL  optional label remap applied at gather time (payload 1 only)
 *   out_seg_dev u8  [b,H,W,payload]   0 where no valid point won the cell
 *   out_depth_dev f32 [b,H,W]    winning z'; -1 for untouched cells; max(z')+1 for cells won by
 *                                an invalid point (reference semantics, :105,:136-139)
 *   out_coords_dev i64 [b,t,H,W,2] or NULL   clamped floor(u'),floor(v') (reference 'result2d', :147)
 * All t frames of one batch item compete in ONE z-buffer (the reference's only_this_ind=None
 * mode); the per-frame mode (only_this_ind=i) is the same call with t=1 on frame i's slices.
 * Ties on depth go to the lowest flattened source index e = replica*t*N + frame*N + v*W + u.
 * ------------------------------------------------------------------------------------- */
/* Work space: one z-buffer (8 bytes per target cell) per plane of the call plus one bit per output cell (default
 * scheme; with PF_ZSPLAT_MODE=slab only the z-buffers of ONE L2-sized group, PF_ZSPLAT_L2_MB MiB, reused by every
 * group).  Sized for the per-frame mode.  The environment is read per call: query with the one you run with. */
size_t pf_zsplat_workspace_bytes(int b, int t, int H, int W);

int pf_zsplat_forward(const float* depth_dev, const uint8_t* mask_dev, const uint8_t* seg_dev,
                      const float* K_dev, const float* Kinv_dev,
                      const float* E_dev, const float* Einv_dev, const float* T_dev,
                      int b, int t, int H, int W, int payload,
                      const uint8_t* lut_dev,
                      uint8_t* out_seg_dev, float* out_depth_dev, int64_t* out_coords_dev,
                      void* workspace_dev, size_t workspace_bytes, void* stream);

/* Per-frame mode: the t frames do NOT compete; frame i is splatted into its own z-buffer, which
 * is what the bg pipeline consumes (reference configs/bg/bg_val_mid.yaml:12-14 reads the
 * `..._ind0_all/_ind1_all/_ind2_all` exports made with model.only_this_ind = 0,1,2).  One call
 * == t reference predict() calls, including their per-call sentinel max(z')+1 (taken over the
 * whole batch of that frame, pc_transform_model.py:105).  Same arguments as pf_zsplat_forward
 * but out_seg_dev is u8 [b,t,H,W,payload] and out_depth_dev f32 [b,t,H,W]. */
int pf_zsplat_forward_frames(const float* depth_dev, const uint8_t* mask_dev, const uint8_t* seg_dev,
                             const float* K_dev, const float* Kinv_dev,
                             const float* E_dev, const float* Einv_dev, const float* T_dev,
                             int b, int t, int H, int W, int payload,
                             const uint8_t* lut_dev,
                             uint8_t* out_seg_dev, float* out_depth_dev, int64_t* out_coords_dev,
                             void* workspace_dev, size_t workspace_bytes, void* stream);

/* Per-frame mode with the disk hop fused into the resolve kernel: out_depth_dev receives the depth as
 * BGDataset would decode it from the exporter's uint16 PNG (see pf_depth_disk_hop) and out_mask_dev
 * (u8 [b,t,H,W]) its validity mask.  payload is 1. */
int pf_zsplat_forward_frames_hop(const float* depth_dev, const uint8_t* mask_dev, const uint8_t* seg_dev,
                                 const float* K_dev, const float* Kinv_dev,
                                 const float* E_dev, const float* Einv_dev, const float* T_dev,
                                 int b, int t, int H, int W, const uint8_t* lut_dev,
                                 uint8_t* out_seg_dev, float* out_depth_dev, uint8_t* out_mask_dev,
                                 float min_depth, float max_depth,
                                 void* workspace_dev, size_t workspace_bytes, void* stream);

/* Per-frame mode + fused disk hop with PACKED inputs (half the bytes of the reference's formats on the wire):
 *   depth_code_dev u16 [b,t,H,W]     depth = depth_lut_dev[code]; the caller builds the 65536-entry f32 table with
 *   depth_lut_dev  f32 [65536]       whatever disparity -> depth formula it uses (the Cityscapes disparity PNGs the
 *                                    reference's dataset decodes, pc_transform_dataset.py:274, are uint16), so the
 *                                    result is bit-identical to pf_zsplat_forward_frames_hop on depth = lut[code]
 *   mask_bits_dev  u8  [b,t,H*W/8]   depth_mask, bit i%8 (LSB first) of byte i/8 of each plane; 16-byte aligned
 * H*W must be a multiple of 8.  Everything else as pf_zsplat_forward_frames_hop. */
int pf_zsplat_forward_frames_hop_packed(const uint16_t* depth_code_dev, const float* depth_lut_dev,
                                        const uint8_t* mask_bits_dev, const uint8_t* seg_dev,
                                        const float* K_dev, const float* Kinv_dev,
                                        const float* E_dev, const float* Einv_dev, const float* T_dev,
                                        int b, int t, int H, int W, const uint8_t* lut_dev,
                                        uint8_t* out_seg_dev, float* out_depth_dev, uint8_t* out_mask_dev,
                                        float min_depth, float max_depth,
                                        void* workspace_dev, size_t workspace_bytes, void* stream);

/* pf_zsplat_forward with HOST buffers (convenience / smoke call): allocates device staging, copies the inputs,
 * runs, copies seg/depth back and synchronises.  Not a throughput path: the pipelined host-buffer front end that
 * bench.py times as `e2e` is panoptic_forecasting_b200.pipeline.PipelinedForecaster (pinned staging, overlapped
 * copies) on top of the device-pointer entry points above. */
int pf_zsplat_forward_host(const float* depth, const uint8_t* mask, const uint8_t* seg,
                           const float* K, const float* Kinv, const float* E, const float* Einv,
                           const float* T, int b, int t, int H, int W, int payload,
                           const uint8_t* lut, uint8_t* out_seg, float* out_depth);

/* ---------------------------------------------------------------------------------------
 * Disk-hop emulation between stage A and stage B (fused on device):
 *   exporter: u16 = round(clamp(d + 1, 0, 255) * 256)
 *             reference experiments/export_cityscapes_segmentation_results.py:119-122
 *   bg dataset: d = u16/256 - 1; mask = d > 0; d[~mask] = -1; d[mask] clamped to [min,max]
 *             reference data/datasets/bg_dataset.py:223-230
 * in/out f32 [n]; out_mask u8 [n].
 * ------------------------------------------------------------------------------------- */
int pf_depth_disk_hop(const float* depth_dev, float* out_depth_dev, uint8_t* out_mask_dev,
                      size_t n, float min_depth, float max_depth, void* stream);

/* ---------------------------------------------------------------------------------------
 * Stage B: BGModel forward / predict on the HarDNet-70 segmentation net.
 * Replaces reference panoptic_forecasting/models/bg/bg_model.py:53-71,91-102 and
 * panoptic_forecasting/models/bg/hardnet.py:16-25,176-240,243-258,262-327,353-387.
 * The topology (stem 16/24/32/48, growth 10/16/18/24/32, layers 4/4/8/8/8, 1x1 widths
 * 64/96/160/224/320) is the reference's; conv index order below is the order in which the
 * reference's forward executes them (base.0 .. base.3, block layers, 1x1, ..., decoder).
 * ------------------------------------------------------------------------------------- */
typedef struct pf_bgnet pf_bgnet_t;

typedef struct {
  int cin, cout, ksize, stride;
  char name[64];      /* state_dict prefix, e.g. "model.base.4.layers.2" */
} pf_conv_info_t;

/* precision: 0 = fp32 SIMT kernels (bit-faithful to 1e-6),
 *            1 = tensor-core (tcgen05) split-bf16 3-pass (fp32-faithful to ~3e-5). */
int pf_bgnet_create(pf_bgnet_t** out, int num_classes, int num_inputs, int use_depth, int precision);
void pf_bgnet_destroy(pf_bgnet_t* net);
int pf_bgnet_num_convs(const pf_bgnet_t* net);          /* ConvLayers (conv+BN+ReLU), excludes finalConv */
int pf_bgnet_conv_info(const pf_bgnet_t* net, int i, pf_conv_info_t* info);
/* HOST pointers; weight is [cout,cin,k,k] fp32 as in the state_dict; BN folded here. */
int pf_bgnet_load_conv(pf_bgnet_t* net, int i, const float* weight, const float* bn_weight,
                       const float* bn_bias, const float* bn_mean, const float* bn_var, float eps);
int pf_bgnet_load_final(pf_bgnet_t* net, const float* weight /*[classes,48]*/, const float* bias);
int pf_bgnet_set_depth_norm(pf_bgnet_t* net, float mean, float std);

/* 0 for an unsupported size: H must be a multiple of 4, W a multiple of 16, both >= 64 (the two stride-2 convs halve
 * exactly, the four AvgPool2d(2,2) floor like the reference's, hardnet.py:300). */
size_t pf_bgnet_workspace_bytes(const pf_bgnet_t* net, int b, int H, int W);

/*   labels_dev  u8  [b,t,H,W]   class ids; ids >= num_classes contribute an all-zero one-hot
 *   depth_dev   f32 [b,t,H,W]   mask_dev u8 [b,t,H,W]
 *   out_seg_*   argmax of the bilinearly (align_corners) upsampled logits at final_h x final_w:
 *               u8 [b,final_h,final_w] and/or i64 (either may be NULL)
 *   out_quarter_dev f32 [b,classes,H/4,W/4] or NULL   ('orig_size_logits')
 *   out_full_dev    f32 [b,classes,final_h,final_w] or NULL ('logits')
 */
int pf_bgnet_forward(pf_bgnet_t* net, const uint8_t* labels_dev, const float* depth_dev,
                     const uint8_t* mask_dev, int b, int H, int W, int final_h, int final_w,
                     uint8_t* out_seg_u8_dev, int64_t* out_seg_i64_dev,
                     float* out_quarter_dev, float* out_full_dev,
                     void* workspace_dev, size_t workspace_bytes, void* stream);

/* The same forward for the reference's `convert2onehot: False` input mode (bg_model.py:61-65: `inps` already is a
 * float tensor of per-class planes): scores_dev f32 [b, num_inputs, classes, H, W] replaces labels_dev; everything
 * else as pf_bgnet_forward.  Only the first ConvLayer differs (dense planes instead of the label look-up). */
int pf_bgnet_forward_dense(pf_bgnet_t* net, const float* scores_dev, const float* depth_dev,
                           const uint8_t* mask_dev, int b, int H, int W, int final_h, int final_w,
                           uint8_t* out_seg_u8_dev, int64_t* out_seg_i64_dev,
                           float* out_quarter_dev, float* out_full_dev,
                           void* workspace_dev, size_t workspace_bytes, void* stream);

/* number of kernel launches one pf_bgnet_forward enqueues (for bench.py's gpu_launches) */
int pf_bgnet_launches_per_forward(const pf_bgnet_t* net);
int pf_zsplat_launches_per_forward(void);
int pf_zsplat_launches_for(int b, int t, int H, int W);   /* per-frame mode: (points + resolve) per L2-sized group + patch */

/* Per-step device timing (CUDA events recorded on the caller's stream around every step of the
 * next `max_iters` forwards); pf_bgnet_read_profile synchronises on the last event and returns
 * the number of profiled forwards, with the mean milliseconds of each step in ms_per_step.
 * step types: 0 first conv (labels), 1 ConvLayer, 2 avg-pool, 3 bilinear upsample, 4 head. */
int pf_bgnet_set_profiling(pf_bgnet_t* net, int max_iters);
int pf_bgnet_num_steps(const pf_bgnet_t* net);
int pf_bgnet_step_info(const pf_bgnet_t* net, int k, int* type, int* conv_index);
int pf_bgnet_read_profile(pf_bgnet_t* net, float* ms_per_step, int cap);

/* Layer-level debug/test hooks: run ONE ConvLayer i on an NCHW fp32 device tensor. */
int pf_bgnet_debug_conv(pf_bgnet_t* net, int i, const float* x_nchw_dev, int b, int H, int W,
                        float* y_nchw_dev, void* stream);

/* Stand-alone fused bilinear(align_corners) upsample + argmax over NCHW fp32 logits
 * (reference hardnet.py:373-377 + bg_model.py:98). */
int pf_upsample_argmax(const float* logits_nchw_dev, int b, int classes, int h, int w,
                       int final_h, int final_w, uint8_t* out_seg_u8_dev, int64_t* out_seg_i64_dev,
                       float* out_full_dev, void* stream);

/* ---- "next" row: fg -> bg panoptic merge ---------------------------------------
 * Replaces the per-instance paste + z-test loop of FGModel.predict_panoptic
 * (panoptic_forecasting/models/fg/fg_model.py:515-518, 557-588) and model_utils.paste_mask
 * (panoptic_forecasting/models/fg/model_utils.py:30-57; bilinear grid_sample, align_corners=False, zeros).
 * Instances of all batch items are concatenated; item i owns [inst_begin[i], inst_begin[i+1]).
 *
 * pf_panoptic_paint_order (fg_model.py:560-577): per item, order[k] = index (into the concatenated arrays) of
 * the k-th painted instance -- depth descending (far to near, stable, NaN first) when depths is given, index
 * order when it is NULL -- and seg_vals[k] = the panoptic id that instance paints,
 * (class + 11) * 1000 + number of earlier painted instances of the same class (:576-577).
 *   classes int64 [n]; depths f32 [n] or NULL; inst_begin int32 [b+1]; outputs int32 [n]. */
int pf_panoptic_paint_order(const int64_t* classes_dev, const float* depths_dev, const int32_t* inst_begin_dev,
                            int b, int32_t* order_dev, int32_t* seg_vals_dev, void* stream);

/* pf_panoptic_merge (:515-518, :557-588): one pass over the frame.
 *   background      int64 [b,H,W] (uint8 if background_is_u8), or NULL (all 255); ids >= 11 become 255 (:517)
 *   bg_depth        f32 [b,H,W] or NULL; bg_depth_mask u8 [b,H,W] or NULL (0 => depth 1e9, :567)
 *   masks           f32 [n, mh, mw] mask probabilities (after the sigmoid, :541)
 *   boxes           f32 [n,4] (16-byte aligned): (x0,y0,x1,y1) if use_bbox_ulbr else (cx,cy,w,h)
 *   depths          f32 [n] or NULL.  The z-test (:583-586) runs iff depths and bg_depth are both given.
 *   seg_vals        int32 [n] indexed by PAINT POSITION; order int32 [n] paint position -> instance index,
 *                   or NULL when masks/boxes/depths are already stored in paint order.
 * Output int64 [b,H,W].  Bit-identical to the reference's CPU result (same float32 operation order). */
int pf_panoptic_merge(const void* background_dev, int background_is_u8, const float* bg_depth_dev,
                      const uint8_t* bg_depth_mask_dev, const float* masks_dev, const float* boxes_dev,
                      const float* depths_dev, const int32_t* seg_vals_dev, const int32_t* order_dev,
                      const int32_t* inst_begin_dev, int b, int H, int W, int mh, int mw, int use_bbox_ulbr,
                      int64_t* out_seg_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* PF_B200_H_ */
