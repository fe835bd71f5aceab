"""fg -> bg panoptic merge on the B200.

Host-side mirror of the tail of the reference's `FGModel.predict_panoptic`
(panoptic_forecasting/models/fg/fg_model.py:515-588): same argument meaning (per-item lists of instance mask
probabilities, boxes, classes, depths; stacked background tensors), same result (`int64 [b, H, W]` panoptic ids).
The paint order and the per-class running instance index are computed on the device (no `.item()` syncs); the
paste + z-test itself is `pf_panoptic_merge` (csrc/panoptic_merge.cu).  No CPU fallback.
"""
import torch

from . import _lib


def paint_order(classes, depths, use_depth_sorting):
    """fg_model.py:560-577 for one item: (order, seg_vals) as device tensors.  `order` is the processing order
    (depth descending when sorting), `seg_vals[k]` the id painted by the k-th processed instance."""
    n = classes.numel()
    if use_depth_sorting and depths is not None:
        order = torch.sort(depths, descending=True).indices          # reference: seq_depths.sort(descending=True)
    else:
        order = torch.arange(n, device=classes.device)
    c = classes[order].to(torch.int64)
    same = c[:, None] == c[None, :]
    inst_id = torch.tril(same, diagonal=-1).sum(1)                  # earlier processed instances of the same class
    return order, ((c + 11) * 1000 + inst_id).to(torch.int32)


def prepare_instances(mask_preds, pred_bboxes, orig_classes, pred_depths=None, use_depth_sorting=True):
    """Concatenates the per-item instance lists in paint order: (masks [n,mh,mw], boxes [n,4], depths [n] | None,
    seg_vals int32 [n], inst_begin int32 [b+1]) on the device of `mask_preds[0]`."""
    b = len(mask_preds)
    dev = mask_preds[0].device
    zsort = bool(use_depth_sorting) and pred_depths is not None
    masks, boxes, depths, vals, begin = [], [], [], [], [0]
    for i in range(b):
        n = mask_preds[i].shape[0]
        order, sv = paint_order(orig_classes[i].to(dev), pred_depths[i] if zsort else None, zsort)
        masks.append(mask_preds[i][order].to(torch.float32))
        boxes.append(pred_bboxes[i][order].to(torch.float32))
        if zsort:
            depths.append(pred_depths[i][order].to(torch.float32))
        vals.append(sv)
        begin.append(begin[-1] + n)
    mh, mw = (mask_preds[0].shape[-2], mask_preds[0].shape[-1])
    masks_t = torch.cat(masks).contiguous() if begin[-1] else torch.zeros((1, mh, mw), device=dev)
    boxes_t = torch.cat(boxes).contiguous() if begin[-1] else torch.zeros((1, 4), device=dev)
    vals_t = torch.cat(vals).contiguous() if begin[-1] else torch.zeros((1,), dtype=torch.int32, device=dev)
    depths_t = torch.cat(depths).contiguous() if (zsort and begin[-1]) else None
    begin_t = torch.tensor(begin, dtype=torch.int32, device=dev)
    return masks_t, boxes_t, depths_t, vals_t, begin_t


def merge_prepared(prepared, b, background=None, background_depths=None, background_depth_masks=None, use_bbox_ulbr=True,
                   height=1024, width=2048, stream=None):
    """One `pf_panoptic_merge` launch over instances from `prepare_instances`."""
    masks_t, boxes_t, depths_t, vals_t, begin_t = prepared
    dev = masks_t.device
    L = _lib.lib()
    if background is not None:
        background = background.to(torch.int64).contiguous()
        height, width = background.shape[-2], background.shape[-1]
    bgd = background_depths.to(torch.float32).contiguous() if (background_depths is not None and depths_t is not None) else None
    bgm = None
    if bgd is not None and background_depth_masks is not None:
        bgm = background_depth_masks.reshape(b, height, width).to(torch.uint8).contiguous()
    out = torch.empty((b, height, width), dtype=torch.int64, device=dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    rc = L.pf_panoptic_merge(ptr(background), ptr(bgd), ptr(bgm), ptr(masks_t), ptr(boxes_t), ptr(depths_t), ptr(vals_t),
                             ptr(begin_t), b, height, width, masks_t.shape[-2], masks_t.shape[-1], 1 if use_bbox_ulbr else 0,
                             ptr(out), st.cuda_stream)
    _lib.check(rc, "pf_panoptic_merge")
    return out


def merge_instances(mask_preds, pred_bboxes, orig_classes, pred_depths=None, background=None, background_depths=None,
                    background_depth_masks=None, use_depth_sorting=True, use_bbox_ulbr=True, height=1024, width=2048,
                    stream=None):
    """mask_preds: list (per batch item) of [n_i, mh, mw] float32 CUDA tensors (sigmoid already applied);
    pred_bboxes: list of [n_i, 4]; orig_classes: list of [n_i] integer tensors; pred_depths: list of [n_i] or None;
    background: [b, H, W] int64 or None; background_depths: [b, H, W] float32 or None;
    background_depth_masks: [b, H, W] (or [b, 1, H, W]) bool or None.  Returns {'seg': int64 [b, H, W]}."""
    b = len(mask_preds)
    if b == 0:
        raise _lib.PFError("merge_instances: empty batch")
    if mask_preds[0].device.type != "cuda":
        raise _lib.PFError("merge_instances: tensors must be CUDA tensors (there is no CPU fallback)")
    prepared = prepare_instances(mask_preds, pred_bboxes, orig_classes, pred_depths, use_depth_sorting)
    out = merge_prepared(prepared, b, background, background_depths, background_depth_masks, use_bbox_ulbr, height, width,
                         stream)
    return {"seg": out}
