"""fg -> bg panoptic merge on the B200.

Host-side mirror of the tail of the reference's `FGModel.predict_panoptic`
(panoptic_forecasting/models/fg/fg_model.py:515-588): same argument meaning (per-item lists of instance mask
probabilities, boxes, classes, depths; stacked background tensors), same result (`int64 [b, H, W]` panoptic ids).
The paint order and the per-class running instance index are computed on the device by
`pf_panoptic_paint_order` (no `.item()` syncs); the paste + z-test itself is `pf_panoptic_merge`
(csrc/panoptic_merge.cu).  No CPU fallback.
"""
import torch

from . import _lib


def prepare_instances(mask_preds, pred_bboxes, orig_classes, pred_depths=None, use_depth_sorting=True, stream=None):
    """Concatenates the per-item instance lists (no reordering) and runs `pf_panoptic_paint_order` on the device
    (fg_model.py:560-577; no `.item()` syncs): returns (masks [n,mh,mw], boxes [n,4], depths [n] | None,
    seg_vals int32 [n], order int32 [n], inst_begin int32 [b+1]) on the device of `mask_preds[0]`."""
    b = len(mask_preds)
    dev = mask_preds[0].device
    zsort = bool(use_depth_sorting) and pred_depths is not None
    begin = [0]
    for i in range(b):
        begin.append(begin[-1] + int(mask_preds[i].shape[0]))
    n = begin[-1]
    mh, mw = (mask_preds[0].shape[-2], mask_preds[0].shape[-1])
    begin_t = torch.tensor(begin, dtype=torch.int32, device=dev)
    if n == 0:
        z = torch.zeros((1,), dtype=torch.int32, device=dev)
        return torch.zeros((1, mh, mw), device=dev), torch.zeros((1, 4), device=dev), None, z, z, begin_t
    masks_t = torch.cat([m.reshape(-1, mh, mw) for m in mask_preds]).to(torch.float32).contiguous()
    boxes_t = torch.cat([x.reshape(-1, 4) for x in pred_bboxes]).to(torch.float32).contiguous()
    classes_t = torch.cat([c.reshape(-1) for c in orig_classes]).to(device=dev, dtype=torch.int64).contiguous()
    depths_t = torch.cat([d.reshape(-1) for d in pred_depths]).to(torch.float32).contiguous() if zsort else None
    order_t = torch.empty((n,), dtype=torch.int32, device=dev)
    vals_t = torch.empty((n,), dtype=torch.int32, device=dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    rc = _lib.lib().pf_panoptic_paint_order(classes_t.data_ptr(), depths_t.data_ptr() if zsort else None, begin_t.data_ptr(), b,
                                            order_t.data_ptr(), vals_t.data_ptr(), st.cuda_stream)
    _lib.check(rc, "pf_panoptic_paint_order")
    return masks_t, boxes_t, depths_t, vals_t, order_t, begin_t


def merge_prepared(prepared, b, background=None, background_depths=None, background_depth_masks=None, use_bbox_ulbr=True,
                   height=1024, width=2048, stream=None):
    """One `pf_panoptic_merge` launch over instances from `prepare_instances`.  `background` may be int64 (the
    reference's dtype) or uint8 (the bg exporter's label map, read as is)."""
    masks_t, boxes_t, depths_t, vals_t, order_t, begin_t = prepared
    dev = masks_t.device
    L = _lib.lib()
    bg_u8 = 0
    if background is not None:
        if background.dtype == torch.uint8:
            bg_u8 = 1
        elif background.dtype != torch.int64:
            background = background.to(torch.int64)
        background = background.contiguous()
        height, width = background.shape[-2], background.shape[-1]
    bgd = background_depths.to(torch.float32).contiguous() if (background_depths is not None and depths_t is not None) else None
    bgm = None
    if bgd is not None and background_depth_masks is not None:
        bgm = background_depth_masks.reshape(b, height, width).contiguous()
        bgm = bgm.view(torch.uint8) if bgm.dtype == torch.bool else bgm.to(torch.uint8)
    out = torch.empty((b, height, width), dtype=torch.int64, device=dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    rc = L.pf_panoptic_merge(ptr(background), bg_u8, ptr(bgd), ptr(bgm), ptr(masks_t), ptr(boxes_t), ptr(depths_t), ptr(vals_t),
                             ptr(order_t), ptr(begin_t), b, height, width, masks_t.shape[-2], masks_t.shape[-1],
                             1 if use_bbox_ulbr else 0, ptr(out), st.cuda_stream)
    _lib.check(rc, "pf_panoptic_merge")
    return out


def merge_instances(mask_preds, pred_bboxes, orig_classes, pred_depths=None, background=None, background_depths=None,
                    background_depth_masks=None, use_depth_sorting=True, use_bbox_ulbr=True, height=1024, width=2048,
                    stream=None):
    """mask_preds: list (per batch item) of [n_i, mh, mw] float32 CUDA tensors (sigmoid already applied);
    pred_bboxes: list of [n_i, 4]; orig_classes: list of [n_i] integer tensors; pred_depths: list of [n_i] or None;
    background: [b, H, W] int64 (or uint8) or None; background_depths: [b, H, W] float32 or None;
    background_depth_masks: [b, H, W] (or [b, 1, H, W]) bool or None.  Returns {'seg': int64 [b, H, W]}."""
    b = len(mask_preds)
    if b == 0:
        raise _lib.PFError("merge_instances: empty batch")
    if mask_preds[0].device.type != "cuda":
        raise _lib.PFError("merge_instances: tensors must be CUDA tensors (there is no CPU fallback)")
    prepared = prepare_instances(mask_preds, pred_bboxes, orig_classes, pred_depths, use_depth_sorting, stream)
    out = merge_prepared(prepared, b, background, background_depths, background_depth_masks, use_bbox_ulbr, height, width,
                         stream)
    return {"seg": out}
