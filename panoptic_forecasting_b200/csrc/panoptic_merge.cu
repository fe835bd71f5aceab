// fg -> bg panoptic merge: paste every forecast instance mask into the background label map with a z-test.
// Replaces the per-instance full-frame loop of FGModel.predict_panoptic
// (panoptic_forecasting/models/fg/fg_model.py:515-518, 557-588) and model_utils.paste_mask (:30-57):
// the reference materialises, per instance, a 1024x2048 grid, a grid_sample output, a thresholded id map and
// two masked full-frame blends (O(instances x 2M px) of HBM traffic); here one thread owns one pixel, walks the
// item's instances in paint order with the running (label, depth) in registers and writes the label once.
//
// HBM-bound byte work: 8 B (int64 background) + 4 B depth + 1 B mask in, 8 B out per pixel.
// Arithmetic is the reference's float32 operation order, INCLUDING the fused multiply-adds of ATen's CPU
// grid_sampler (oracle/panoptic_merge_oracle.py documents how that order was pinned), so the >= 0.5 threshold
// and the z-test fall on the same side as in the reference for every pixel.
#include "pf_common.cuh"

namespace pf {

constexpr int kMergeThreads = 256;
constexpr int kMergeChunk = 64;          // instances staged in shared memory per pass

struct MergeParams {
  const long long* background;   // [b,H,W] or null
  const float* bg_depth;         // [b,H,W] or null
  const uint8_t* bg_mask;        // [b,H,W] or null
  const float* masks;            // [n, mh, mw]
  const float* boxes;            // [n, 4]
  const float* depths;           // [n] or null
  const int* seg_vals;           // [n]
  const int* inst_begin;         // [b + 1]
  long long* out;                // [b,H,W]
  int H, W, mh, mw, ulbr;
};

struct Inst {
  float x0, y0, dx, dy;          // box origin and extent (x1 - x0, y1 - y0), reference rounding
  float lox, hix, loy, hiy;      // conservative pixel-centre window outside of which every bilinear tap is out of bounds
  float depth;
  int val, id;
};

// one axis of model_utils.paste_mask + ATen's unnormalize: pixel centre -> source coordinate
__device__ __forceinline__ float src_coord(float p, float lo, float ext, float half_size) {
  const float g = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(p, lo), ext), 2.0f), 1.0f);   // (p - lo) / (hi - lo) * 2 - 1
  return __fmaf_rn(__fadd_rn(g, 1.0f), half_size, -0.5f);                               // fma(g + 1, size / 2, -0.5)
}

__global__ void __launch_bounds__(kMergeThreads) panoptic_merge_kernel(const MergeParams p) {
  __shared__ Inst inst[kMergeChunk];
  const int bi = blockIdx.y;
  const int N = p.H * p.W;
  const int pix = blockIdx.x * kMergeThreads + threadIdx.x;
  const bool live = pix < N;
  const int y = live ? pix / p.W : 0, x = live ? pix - y * p.W : 0;
  const size_t gp = (size_t)bi * N + (live ? pix : 0);
  const bool zmode = p.depths != nullptr && p.bg_depth != nullptr;               // fg_model.py:582
  long long label = 255;                                                         // :520 (no background given)
  if (p.background && live) {
    label = p.background[gp];
    if (label >= 11) label = 255;                                                // :517
  }
  float cur = 0.f;
  if (zmode && live) {
    cur = p.bg_depth[gp];
    if (p.bg_mask && !p.bg_mask[gp]) cur = 1000000000.0f;                        // :567
  }
  const float px = __fadd_rn((float)x, 0.5f), py = __fadd_rn((float)y, 0.5f);    // model_utils.py:42-43
  const float hmw = (float)p.mw * 0.5f, hmh = (float)p.mh * 0.5f;
  const int k0 = p.inst_begin[bi], k1 = p.inst_begin[bi + 1];
  for (int base = k0; base < k1; base += kMergeChunk) {
    const int cnt = min(kMergeChunk, k1 - base);
    __syncthreads();
    if (threadIdx.x < cnt) {
      const int k = base + threadIdx.x;
      const float b0 = p.boxes[4 * k], b1 = p.boxes[4 * k + 1], b2 = p.boxes[4 * k + 2], b3 = p.boxes[4 * k + 3];
      float x0, y0, x1, y1;
      if (p.ulbr) { x0 = b0; y0 = b1; x1 = b2; y1 = b3; }
      else {                                                                     // model_utils.py:36-40
        const float hw = __fdiv_rn(b2, 2.0f), hh = __fdiv_rn(b3, 2.0f);
        x0 = __fsub_rn(b0, hw); x1 = __fadd_rn(b0, hw); y0 = __fsub_rn(b1, hh); y1 = __fadd_rn(b1, hh);
      }
      Inst t;
      t.x0 = x0; t.y0 = y0; t.dx = __fsub_rn(x1, x0); t.dy = __fsub_rn(y1, y0);
      // a tap is in bounds only if the source coordinate lies in (-1, size): |offset from the box| < 0.5 mask
      // pixel; the window below is 2 mask pixels wide on each side, far beyond any rounding of the exact chain
      const float padx = fabsf(t.dx) / (float)p.mw * 2.0f + 1.0f, pady = fabsf(t.dy) / (float)p.mh * 2.0f + 1.0f;
      t.lox = fminf(x0, x1) - padx; t.hix = fmaxf(x0, x1) + padx;
      t.loy = fminf(y0, y1) - pady; t.hiy = fmaxf(y0, y1) + pady;
      t.depth = p.depths ? p.depths[k] : 0.f;
      t.val = p.seg_vals[k];
      t.id = k;
      inst[threadIdx.x] = t;
    }
    __syncthreads();
    if (!live) continue;
    for (int j = 0; j < cnt; ++j) {
      const Inst& t = inst[j];
      if (!(px > t.lox && px < t.hix && py > t.loy && py < t.hiy)) continue;
      if (zmode && !(t.depth < cur)) continue;                                   // :583 (checked first: cheaper)
      const float ix = src_coord(px, t.x0, t.dx, hmw), iy = src_coord(py, t.y0, t.dy, hmh);
      const float fx = floorf(ix), fy = floorf(iy);
      if (!(fx >= -1.0f && fx < (float)p.mw && fy >= -1.0f && fy < (float)p.mh)) continue;   // all four taps out of bounds (or NaN)
      const int xw = (int)fx, yn = (int)fy;
      const float w = __fsub_rn(ix, fx), e = __fsub_rn(1.0f, w), n = __fsub_rn(iy, fy), s = __fsub_rn(1.0f, n);
      const float* m = p.masks + (size_t)t.id * p.mh * p.mw;
      const bool xin0 = xw >= 0, xin1 = xw + 1 < p.mw, yin0 = yn >= 0, yin1 = yn + 1 < p.mh;
      const float vnw = (xin0 && yin0) ? __ldg(m + yn * p.mw + xw) : 0.f;
      const float vne = (xin1 && yin0) ? __ldg(m + yn * p.mw + xw + 1) : 0.f;
      const float vsw = (xin0 && yin1) ? __ldg(m + (yn + 1) * p.mw + xw) : 0.f;
      const float vse = (xin1 && yin1) ? __ldg(m + (yn + 1) * p.mw + xw + 1) : 0.f;
      float v = __fmul_rn(vnw, __fmul_rn(s, e));
      v = __fmaf_rn(vne, __fmul_rn(s, w), v);
      v = __fmaf_rn(vsw, __fmul_rn(n, e), v);
      v = __fmaf_rn(vse, __fmul_rn(n, w), v);
      if (v >= 0.5f) {                                                           // :579
        label = t.val;                                                           // :584-585 / :588-589
        cur = t.depth;                                                           // :586
      }
    }
  }
  if (live) p.out[gp] = label;
}

}  // namespace pf

using namespace pf;

extern "C" int pf_panoptic_merge(const int64_t* background_dev, const float* bg_depth_dev, const uint8_t* bg_depth_mask_dev,
                                 const float* masks_dev, const float* boxes_dev, const float* depths_dev,
                                 const int32_t* seg_vals_dev, const int32_t* inst_begin_dev, int b, int H, int W, int mh,
                                 int mw, int use_bbox_ulbr, int64_t* out_seg_dev, void* stream) {
  PF_REQUIRE(out_seg_dev && inst_begin_dev, PF_EINVAL, "pf_panoptic_merge: null pointer argument");
  PF_REQUIRE(b > 0 && H > 0 && W > 0 && mh > 0 && mw > 0, PF_EINVAL, "pf_panoptic_merge: non-positive size");
  PF_REQUIRE(b <= 65535 && (double)H * W < 2147483647.0, PF_EINVAL, "pf_panoptic_merge: size out of range");
  MergeParams p;
  p.background = reinterpret_cast<const long long*>(background_dev);
  p.bg_depth = bg_depth_dev; p.bg_mask = bg_depth_mask_dev;
  p.masks = masks_dev; p.boxes = boxes_dev; p.depths = depths_dev; p.seg_vals = seg_vals_dev;
  p.inst_begin = inst_begin_dev;
  p.out = reinterpret_cast<long long*>(out_seg_dev);
  p.H = H; p.W = W; p.mh = mh; p.mw = mw; p.ulbr = use_bbox_ulbr ? 1 : 0;
  const int N = H * W;
  panoptic_merge_kernel<<<dim3(cdiv(N, kMergeThreads), b), kMergeThreads, 0, (cudaStream_t)stream>>>(p);
  PF_CHECK_CUDA(cudaGetLastError());
  return 0;
}
