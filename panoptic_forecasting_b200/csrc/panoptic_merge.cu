// fg -> bg panoptic merge: paste every forecast instance mask into the background label map with a z-test.
// Replaces the per-instance full-frame loop of FGModel.predict_panoptic
// (panoptic_forecasting/models/fg/fg_model.py:515-518, 557-588) and model_utils.paste_mask (:30-57):
// the reference materialises, per instance, a 1024x2048 grid, a grid_sample output, a thresholded id map and
// two masked full-frame blends (O(instances x 2M px) of HBM traffic); here every pixel is read once and written
// once, with the running (label, depth) in registers while the item's instances are walked in paint order.
//
// HBM-bound byte work: 8 B (int64 background; 1 B for the uint8 variant) + 4 B depth + 1 B mask in, 8 B out per
// pixel = 21 B (14 B).  Work decomposition: one WARP owns a 128-pixel segment of one image row, one lane 4
// consecutive pixels (256-bit loads/stores of the int64 maps, 128-bit of the depth).  The warp culls the instance list against its segment with one
// ballot per 32 instances (lane j tests instance j: exact row test, conservative column window) and broadcasts
// the survivors' parameters by shuffle, so pixels outside every box cost no per-instance work and nothing is
// staged in shared memory.  The row coordinate (iy, floor, weights) is computed once per (warp, instance).
//
// Arithmetic is the reference's float32 operation order, INCLUDING the fused multiply-adds of ATen's CPU
// grid_sampler (oracle/panoptic_merge_oracle.py documents how that order was pinned), so the >= 0.5 threshold
// and the z-test fall on the same side as in the reference for every pixel.
#include "pf_common.cuh"

namespace pf {

constexpr int kMergeThreads = 256;
constexpr int kMergeWarps = kMergeThreads / 32;
constexpr int kSegPx = 128;              // pixels per warp (4 per lane)
constexpr unsigned kFull = 0xffffffffu;

struct MergeParams {
  const void* background;        // [b,H,W] int64 or uint8, or null
  const float* bg_depth;         // [b,H,W] or null
  const uint8_t* bg_mask;        // [b,H,W] or null
  const float* masks;            // [n, mh, mw]
  const float* boxes;            // [n, 4]
  const float* depths;           // [n] or null
  const int* seg_vals;           // [n], by paint position
  const int* order;              // [n] paint position -> instance index, or null (identity)
  const int* inst_begin;         // [b + 1]
  long long* out;                // [b,H,W]
  int H, W, mh, mw, ulbr, bg_u8, segs_per_row;
  int vec;                       // W % 4 == 0 and every per-pixel pointer aligned for the 4-pixel vector accesses
};

// one axis of model_utils.paste_mask + ATen's unnormalize: pixel centre -> source coordinate
__device__ __forceinline__ float src_coord(float p, float lo, float ext, float half_size) {
  const float g = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(p, lo), ext), 2.0f), 1.0f);   // (p - lo) / (hi - lo) * 2 - 1
  return __fmaf_rn(__fadd_rn(g, 1.0f), half_size, -0.5f);                               // fma(g + 1, size / 2, -0.5)
}

__global__ void __launch_bounds__(kMergeThreads) panoptic_merge_kernel(const MergeParams p) {
  const int lane = threadIdx.x & 31;
  const int bi = blockIdx.y;
  const long long seg = (long long)blockIdx.x * kMergeWarps + (threadIdx.x >> 5);
  const int y = (int)(seg / p.segs_per_row);
  if (y >= p.H) return;                                                          // warp-uniform
  const int xs = (int)(seg - (long long)y * p.segs_per_row) * kSegPx;
  const int x = xs + lane * 4;
  const size_t g0 = ((size_t)bi * p.H + y) * p.W + x;
  const bool vec = p.vec != 0;                                                   // W % 4 == 0: x < W implies x + 3 < W
  const bool zmode = p.depths != nullptr && p.bg_depth != nullptr;               // fg_model.py:582

  long long label[4];
  float cur[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { label[q] = 255; cur[q] = 0.f; }                  // :520 (no background given)
  if (x < p.W) {
    if (p.background) {
      if (p.bg_u8) {
        const uint8_t* s = reinterpret_cast<const uint8_t*>(p.background) + g0;
        if (vec) {
          const uchar4 v = __ldcs(reinterpret_cast<const uchar4*>(s));
          label[0] = v.x; label[1] = v.y; label[2] = v.z; label[3] = v.w;
        } else {
          for (int q = 0; q < 4; ++q) if (x + q < p.W) label[q] = s[q];
        }
      } else {
        const long long* s = reinterpret_cast<const long long*>(p.background) + g0;
        if (vec) {                                                              // 4 x int64 = one 256-bit load (sm_100)
          asm volatile("ld.global.cs.v4.b64 {%0, %1, %2, %3}, [%4];"
                       : "=l"(label[0]), "=l"(label[1]), "=l"(label[2]), "=l"(label[3]) : "l"(s));
        } else {
          for (int q = 0; q < 4; ++q) if (x + q < p.W) label[q] = s[q];
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) if (label[q] >= 11) label[q] = 255;            // :517
    }
    if (zmode) {
      if (vec) {
        const float4 d = __ldcs(reinterpret_cast<const float4*>(p.bg_depth + g0));
        cur[0] = d.x; cur[1] = d.y; cur[2] = d.z; cur[3] = d.w;
        if (p.bg_mask) {
          const uchar4 m = __ldcs(reinterpret_cast<const uchar4*>(p.bg_mask + g0));
          if (!m.x) cur[0] = 1000000000.0f;                                      // :567
          if (!m.y) cur[1] = 1000000000.0f;
          if (!m.z) cur[2] = 1000000000.0f;
          if (!m.w) cur[3] = 1000000000.0f;
        }
      } else {
        for (int q = 0; q < 4; ++q)
          if (x + q < p.W) {
            cur[q] = p.bg_depth[g0 + q];
            if (p.bg_mask && !p.bg_mask[g0 + q]) cur[q] = 1000000000.0f;
          }
      }
    }
  }

  const float py = __fadd_rn((float)y, 0.5f);                                    // model_utils.py:42-43
  const float hmw = (float)p.mw * 0.5f, hmh = (float)p.mh * 0.5f;
  const float seg_lo = (float)xs + 0.5f, seg_hi = (float)min(xs + kSegPx - 1, p.W - 1) + 0.5f;
  const int k0 = p.inst_begin[bi], k1 = p.inst_begin[bi + 1];
  for (int base = k0; base < k1; base += 32) {
    // lane j examines the instance at paint position base + j
    const int k = base + lane;
    bool want = false;
    float i_x0 = 0.f, i_dx = 1.f, i_lox = 0.f, i_hix = 0.f, i_depth = 0.f, i_n = 0.f;
    int i_val = 0, i_id = 0, i_yn = 0;
    if (k < k1) {
      i_id = p.order ? p.order[k] : k;
      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.boxes) + i_id);
      float x0, y0, x1, y1;
      if (p.ulbr) { x0 = bb.x; y0 = bb.y; x1 = bb.z; y1 = bb.w; }
      else {                                                                     // model_utils.py:36-40
        const float hw = __fdiv_rn(bb.z, 2.0f), hh = __fdiv_rn(bb.w, 2.0f);
        x0 = __fsub_rn(bb.x, hw); x1 = __fadd_rn(bb.x, hw); y0 = __fsub_rn(bb.y, hh); y1 = __fadd_rn(bb.y, hh);
      }
      i_x0 = x0; i_dx = __fsub_rn(x1, x0);
      // a tap is in bounds only if the source coordinate lies in (-1, size): |offset from the box| < 0.5 mask
      // pixel; the column window below is 2 mask pixels wide on each side, far beyond any rounding of the chain
      const float padx = fabsf(i_dx) / (float)p.mw * 2.0f + 1.0f;
      i_lox = fminf(x0, x1) - padx; i_hix = fmaxf(x0, x1) + padx;
      const float iy = src_coord(py, y0, __fsub_rn(y1, y0), hmh), fy = floorf(iy);
      want = (fy >= -1.0f && fy < (float)p.mh) && (seg_hi > i_lox && seg_lo < i_hix);   // row test exact, false on NaN
      i_yn = want ? (int)fy : 0;
      i_n = __fsub_rn(iy, fy);
      i_depth = p.depths ? p.depths[i_id] : 0.f;
      i_val = p.seg_vals[k];
    }
    unsigned todo = __ballot_sync(kFull, want);
    while (todo) {                                                               // ascending j == paint order
      const int j = __ffs(todo) - 1;
      todo &= todo - 1;
      const float x0 = __shfl_sync(kFull, i_x0, j), dx = __shfl_sync(kFull, i_dx, j);
      const float lox = __shfl_sync(kFull, i_lox, j), hix = __shfl_sync(kFull, i_hix, j);
      const float depth = __shfl_sync(kFull, i_depth, j), n = __shfl_sync(kFull, i_n, j);
      const int val = __shfl_sync(kFull, i_val, j), id = __shfl_sync(kFull, i_id, j), yn = __shfl_sync(kFull, i_yn, j);
      const float s = __fsub_rn(1.0f, n);
      const bool yin0 = yn >= 0, yin1 = yn + 1 < p.mh;
      const float* m0 = p.masks + (size_t)id * p.mh * p.mw + yn * p.mw;          // row yn (dereferenced only if yin0)
      const float* m1 = m0 + p.mw;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float px = __fadd_rn((float)(x + q), 0.5f);
        if (!(px > lox && px < hix)) continue;
        if (zmode && !(depth < cur[q])) continue;                                // :583 (checked first: cheaper)
        const float ix = src_coord(px, x0, dx, hmw), fx = floorf(ix);
        if (!(fx >= -1.0f && fx < (float)p.mw)) continue;                        // both columns out of bounds (or NaN)
        const int xw = (int)fx;
        const float w = __fsub_rn(ix, fx), e = __fsub_rn(1.0f, w);
        const bool xin0 = xw >= 0, xin1 = xw + 1 < p.mw;
        const float vnw = (xin0 && yin0) ? __ldg(m0 + xw) : 0.f;
        const float vne = (xin1 && yin0) ? __ldg(m0 + xw + 1) : 0.f;
        const float vsw = (xin0 && yin1) ? __ldg(m1 + xw) : 0.f;
        const float vse = (xin1 && yin1) ? __ldg(m1 + xw + 1) : 0.f;
        float v = __fmul_rn(vnw, __fmul_rn(s, e));
        v = __fmaf_rn(vne, __fmul_rn(s, w), v);
        v = __fmaf_rn(vsw, __fmul_rn(n, e), v);
        v = __fmaf_rn(vse, __fmul_rn(n, w), v);
        if (v >= 0.5f) {                                                         // :579
          label[q] = val;                                                        // :584-585 / :588-589
          cur[q] = depth;                                                        // :586
        }
      }
    }
  }
  if (x < p.W) {
    long long* d = p.out + g0;
    if (vec) {                                                                  // one 256-bit store: a full sector per lane
      asm volatile("st.global.cs.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(d), "l"(label[0]), "l"(label[1]), "l"(label[2]), "l"(label[3])
                   : "memory");
    } else {
      for (int q = 0; q < 4; ++q) if (x + q < p.W) d[q] = label[q];
    }
  }
}

// fg_model.py:560-577 for every batch item: paint order (depth descending, stable; index order when depths is null)
// and the id each painted instance writes, (class + 11) * 1000 + number of earlier painted instances of that class.
// One block per item, rank sort (instance counts are tens, at most a few hundred).
__device__ __forceinline__ bool painted_before(float da, int a, float db, int b) {
  const bool na = da != da, nb = db != db;                                       // NaN sorts first (torch.sort descending)
  if (na || nb) return (na && !nb) || (na == nb && a < b);
  return da > db || (da == db && a < b);
}

__global__ void paint_order_kernel(const long long* classes, const float* depths, const int* inst_begin, int* order,
                                   int* seg_vals) {
  const int k0 = inst_begin[blockIdx.x], n = inst_begin[blockIdx.x + 1] - k0;
  for (int a = threadIdx.x; a < n; a += blockDim.x) {
    int rank = a;
    if (depths) {
      const float da = depths[k0 + a];
      rank = 0;
      for (int b = 0; b < n; ++b) rank += (b != a && painted_before(depths[k0 + b], b, da, a)) ? 1 : 0;
    }
    order[k0 + rank] = k0 + a;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const long long c = classes[order[k0 + r]];
    int cnt = 0;
    for (int q = 0; q < r; ++q) cnt += classes[order[k0 + q]] == c ? 1 : 0;
    seg_vals[k0 + r] = (int)((c + 11) * 1000 + cnt);
  }
}

}  // namespace pf

using namespace pf;

extern "C" int pf_panoptic_paint_order(const int64_t* classes_dev, const float* depths_dev, const int32_t* inst_begin_dev,
                                       int b, int32_t* order_dev, int32_t* seg_vals_dev, void* stream) {
  PF_REQUIRE(classes_dev && inst_begin_dev && order_dev && seg_vals_dev, PF_EINVAL, "pf_panoptic_paint_order: null pointer argument");
  PF_REQUIRE(b > 0, PF_EINVAL, "pf_panoptic_paint_order: non-positive batch");
  paint_order_kernel<<<b, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(classes_dev), depths_dev,
                                                          inst_begin_dev, order_dev, seg_vals_dev);
  PF_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int pf_panoptic_merge(const void* background_dev, int background_is_u8, const float* bg_depth_dev,
                                 const uint8_t* bg_depth_mask_dev, const float* masks_dev, const float* boxes_dev,
                                 const float* depths_dev, const int32_t* seg_vals_dev, const int32_t* order_dev,
                                 const int32_t* inst_begin_dev, int b, int H, int W, int mh, int mw, int use_bbox_ulbr,
                                 int64_t* out_seg_dev, void* stream) {
  PF_REQUIRE(out_seg_dev && inst_begin_dev, PF_EINVAL, "pf_panoptic_merge: null pointer argument");
  PF_REQUIRE(b > 0 && H > 0 && W > 0 && mh > 0 && mw > 0, PF_EINVAL, "pf_panoptic_merge: non-positive size");
  PF_REQUIRE(b <= 65535 && (double)H * W < 2147483647.0, PF_EINVAL, "pf_panoptic_merge: size out of range");
  MergeParams p;
  p.background = background_dev; p.bg_u8 = background_is_u8 ? 1 : 0;
  p.bg_depth = bg_depth_dev; p.bg_mask = bg_depth_mask_dev;
  p.masks = masks_dev; p.boxes = boxes_dev; p.depths = depths_dev; p.seg_vals = seg_vals_dev; p.order = order_dev;
  p.inst_begin = inst_begin_dev;
  p.out = reinterpret_cast<long long*>(out_seg_dev);
  p.H = H; p.W = W; p.mh = mh; p.mw = mw; p.ulbr = use_bbox_ulbr ? 1 : 0;
  p.segs_per_row = cdiv(W, kSegPx);
  auto aligned = [](const void* q, size_t a) { return q == nullptr || (reinterpret_cast<size_t>(q) & (a - 1)) == 0; };
  p.vec = ((W & 3) == 0 && aligned(background_dev, background_is_u8 ? 4 : 32) && aligned(bg_depth_dev, 16) &&
           aligned(bg_depth_mask_dev, 4) && aligned(out_seg_dev, 32)) ? 1 : 0;
  const long long segs = (long long)p.segs_per_row * H;
  panoptic_merge_kernel<<<dim3((unsigned)((segs + kMergeWarps - 1) / kMergeWarps), b), kMergeThreads, 0, (cudaStream_t)stream>>>(p);
  PF_CHECK_CUDA(cudaGetLastError());
  return 0;
}
