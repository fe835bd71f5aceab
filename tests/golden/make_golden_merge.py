"""Generates tests/golden/merge_*.npz by running the UNMODIFIED reference (build container only):
`FGModel.predict_panoptic` (panoptic_forecasting/models/fg/fg_model.py:489-596) is called as a plain function
on a stand-in object whose forward returns prescribed mask logits / trajectories, so the paste + z-test merge
loop, `paste_mask` and `grid_sample` that run are the reference's own.
    python tests/golden/make_golden_merge.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from panoptic_forecasting_b200 import synthetic  # noqa: E402

H, W = 1024, 2048          # hard-coded in the reference's paste call (fg_model.py:574)


class StandIn:
    """Carries the attributes predict_panoptic reads and replays a prescribed network output."""

    def __init__(self, pred, use_depth_sorting, use_bbox_ulbr):
        self.pred = pred
        self.use_depth_inp = True
        self.only_loc_feats = False
        self.use_depth_sorting = use_depth_sorting
        self.use_bbox_ulbr = use_bbox_ulbr

    def __call__(self, *a, **k):
        return self.pred


def make_case(seed, n_per_item, use_depth_sorting, use_bbox_ulbr, with_bg_depth, with_bg_mask):
    case = synthetic.make_merge_inputs(len(n_per_item), n_per_item, H, W, seed=seed, use_bbox_ulbr=use_bbox_ulbr)
    ref_loader.load_reference()
    from panoptic_forecasting.models.fg.fg_model import FGModel
    t_in, t_out, D = 3, 3, 9
    ntot = sum(n_per_item)
    # network output: trajectories [ntot, t_in + t_out, D] with bbox in [:4] and depth in [8] at the output index
    traj = torch.zeros(ntot, t_in + t_out, D)
    out_inds = torch.full((ntot,), t_out - 1, dtype=torch.long)
    bb = torch.from_numpy(np.concatenate(case["bboxes"]))
    dp = torch.from_numpy(np.concatenate(case["depths"]))
    traj[:, -1, :4] = bb
    traj[:, -1, 8] = dp
    logits = torch.from_numpy(np.concatenate(case["mask_logits"]))
    pred = {"unnormalized_trajectory": traj, "masks": logits}
    split = lambda x: list(torch.split(x, list(n_per_item)))  # noqa: E731
    inputs = {
        "trajectories": split(torch.zeros(ntot, t_in, D)),
        "bbox_masks": split(torch.ones(ntot, t_in + t_out)),
        "bbox_vel_masks": split(torch.ones(ntot, t_in + t_out)),
        "feats": split(torch.zeros(ntot, 1)),
        "classes": [torch.from_numpy(c).long() for c in case["classes"]],
        "background": [torch.from_numpy(b).long() for b in case["background"]],
    }
    if with_bg_depth:
        inputs["background_depth"] = [torch.from_numpy(d) for d in case["bg_depth"]]
        if with_bg_mask:
            # the reference indexes a [1, H, W] view with this mask (fg_model.py:564-567): items must be [1, H, W]
            inputs["background_depth_mask"] = [torch.from_numpy(m)[None] for m in case["bg_depth_mask"]]
    labels = {"trajectories": split(torch.zeros(ntot, t_out, D)), "output_inds": split(out_inds)}
    with torch.no_grad():
        res = FGModel.predict_panoptic(StandIn(pred, use_depth_sorting, use_bbox_ulbr), inputs, labels)
    return case, res["seg"].numpy().astype(np.int64)


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    cases = [
        ("merge_zsort_ulbr", dict(seed=0, n_per_item=(6, 9), use_depth_sorting=True, use_bbox_ulbr=True, with_bg_depth=True, with_bg_mask=True)),
        ("merge_zsort_cxcywh_nomask", dict(seed=1, n_per_item=(5,), use_depth_sorting=True, use_bbox_ulbr=False, with_bg_depth=True, with_bg_mask=False)),
        ("merge_order_ulbr", dict(seed=2, n_per_item=(7,), use_depth_sorting=False, use_bbox_ulbr=True, with_bg_depth=False, with_bg_mask=False)),
    ]
    for name, kw in cases:
        case, seg = make_case(**kw)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), seg=seg.astype(np.int32),
                            **{k: np.asarray(v) for k, v in kw.items()})
        print(name, seg.shape, "painted px:", int((seg >= 1000).sum()), "ids:", np.unique(seg[seg >= 1000])[:12])


if __name__ == "__main__":
    main()
