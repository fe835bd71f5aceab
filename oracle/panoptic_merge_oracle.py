"""CPU restatement (test infrastructure only) of the reference's fg -> bg panoptic merge.

Follows panoptic_forecasting/models/fg/fg_model.py:515-518 (background ids >= 11 -> 255),
:557-588 (depth-sorted paste + z-test loop of `predict_panoptic`) and
panoptic_forecasting/models/fg/model_utils.py:30-57 (`paste_mask`: bilinear `grid_sample`,
align_corners=False, zero padding).  Pinned against the UNMODIFIED reference by
tests/golden/make_golden_merge.py, which runs `FGModel.predict_panoptic` itself (with a stand-in
for the network forward) and commits its outputs; tests/test_oracle.py replays them through this file.

All arithmetic is float32 in the reference's operation order, including the fused multiply-adds of ATen's
vectorised CPU `grid_sampler_2d` (found by bit-comparison with the reference, 0 of 131 072 samples differ):
  grid:   g = ((p + 0.5) - lo) / (hi - lo) * 2 - 1                       (model_utils.py:42-45, torch elementwise)
  sample: i = fma(g + 1, size / 2, -0.5);  i0 = floor(i);  w = i - i0;  e = 1 - w
          v = fma(se, n*w, fma(sw, n*e, fma(ne, s*w, nw * (s*e))))          (bilinear, zeros padding)
(fma emulated in float64: the float32 product is exact there.)
Only `oracle/` consumers: tests/ and __graft_entry__.smoke().
"""
import numpy as np

f32 = np.float32


def bbox_corners(bbox, use_bbox_ulbr):
    """model_utils.py:33-40: (x0, y0, x1, y1) as float32 scalars."""
    b = np.asarray(bbox, dtype=f32)
    if use_bbox_ulbr:
        return b[0], b[1], b[2], b[3]
    cx, cy, w, h = b[0], b[1], b[2], b[3]
    two = f32(2)
    return f32(cx - f32(w / two)), f32(cy - f32(h / two)), f32(cx + f32(w / two)), f32(cy + f32(h / two))


def _fma(a, b, c):
    return (np.asarray(a, dtype=np.float64) * np.asarray(b, dtype=np.float64) + np.asarray(c, dtype=np.float64)).astype(f32)


def _axis(n, lo, hi, size):
    """pixel centres -> (floor index, weight toward +1 neighbour) along one axis; all float32."""
    p = np.arange(n, dtype=f32) + f32(0.5)
    g = (p - lo) / f32(hi - lo) * f32(2) - f32(1)
    i = _fma((g + f32(1)).astype(f32), f32(size / 2), f32(-0.5))
    i0 = np.floor(i)
    w = (i - i0).astype(f32)
    return i0.astype(np.int64), w, (f32(1) - w).astype(f32)


def paste_mask(mask, bbox, img_h, img_w, use_bbox_ulbr):
    """mask [mh, mw] float32 -> [img_h, img_w] float32 (model_utils.paste_mask for one instance)."""
    mask = np.asarray(mask, dtype=f32)
    mh, mw = mask.shape
    x0, y0, x1, y1 = bbox_corners(bbox, use_bbox_ulbr)
    ix0, wx, ex = _axis(img_w, x0, x1, mw)
    iy0, wy, ey = _axis(img_h, y0, y1, mh)

    def tap(iy, ix):
        ok = ((iy >= 0) & (iy < mh))[:, None] & ((ix >= 0) & (ix < mw))[None, :]
        v = mask[np.clip(iy, 0, mh - 1)[:, None], np.clip(ix, 0, mw - 1)[None, :]]
        return np.where(ok, v, f32(0)).astype(f32)

    n, s = wy[:, None], ey[:, None]          # weight of the lower row (n) / upper row (s)
    w, e = wx[None, :], ex[None, :]
    out = (tap(iy0, ix0) * (s * e).astype(f32)).astype(f32)
    out = _fma(tap(iy0, ix0 + 1), (s * w).astype(f32), out)
    out = _fma(tap(iy0 + 1, ix0), (n * e).astype(f32), out)
    out = _fma(tap(iy0 + 1, ix0 + 1), (n * w).astype(f32), out)
    return out


def paint_order(classes, depths, use_depth_sorting):
    """fg_model.py:560-577: processing order and the panoptic id of every instance.
    Returns (order, seg_vals) with seg_vals[k] the id painted by the k-th processed instance."""
    n = len(classes)
    if use_depth_sorting and depths is not None:
        order = list(np.argsort(-np.asarray(depths, dtype=f32), kind="stable"))
    else:
        order = list(range(n))
    counts = {}
    seg_vals = []
    for k in order:
        c = int(classes[k])
        seg_vals.append((c + 11) * 1000 + counts.get(c, 0))
        counts[c] = counts.get(c, 0) + 1
    return order, seg_vals


def merge(background, masks, bboxes, classes, depths=None, bg_depth=None, bg_depth_mask=None,
          use_depth_sorting=True, use_bbox_ulbr=True, order=None):
    """One batch item.  background [H, W] int64 (or None -> all 255); masks [n, mh, mw] float32 in [0, 1];
    bboxes [n, 4]; classes [n]; depths [n] or None; bg_depth [H, W] float32 / bg_depth_mask [H, W] bool or None.
    Returns int64 [H, W] (fg_model.py:557-588)."""
    if background is None:
        raise ValueError("background required")
    out = np.array(background, dtype=np.int64, copy=True)
    out[out >= 11] = 255                                               # :517
    H, W = out.shape
    zsort = bool(use_depth_sorting) and depths is not None
    cur = None
    if zsort and bg_depth is not None:
        cur = np.array(bg_depth, dtype=f32, copy=True)
        if bg_depth_mask is not None:
            cur[~np.asarray(bg_depth_mask, dtype=bool)] = f32(1000000000)   # :567
    if order is None:
        order, seg_vals = paint_order(classes, depths, zsort)
    else:
        _, seg_vals = paint_order([classes[k] for k in order], None, False)
    for k, val in zip(order, seg_vals):
        hit = paste_mask(masks[k], bboxes[k], H, W, use_bbox_ulbr) >= f32(0.5)   # :579
        if cur is not None:
            d = f32(depths[k])
            hit = hit & (d < cur)                                        # :583
            cur[hit] = d                                                 # :586
        out[hit] = val                                                   # :584-585 / :588-589
    return out
