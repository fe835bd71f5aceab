"""GPU parity of the composite path (Stage A per frame -> disk hop -> Stage B) vs the composite
oracle of SURVEY.md section 8c."""
import numpy as np
import pytest
import torch

from conftest import bg_params
from oracle import bg_oracle, pc_transform_oracle
from panoptic_forecasting_b200 import synthetic
from panoptic_forecasting_b200.models import build_model
from panoptic_forecasting_b200.pipeline import BGForecastPipeline
from test_bgnet_gpu import check_against

pytestmark = pytest.mark.gpu


def disk_hop_np(d, mn=0.1, mx=200.0):
    """export_cityscapes_segmentation_results.py:119-122 then bg_dataset.py:223-230,166-170."""
    q = ((torch.from_numpy(d) + 1).clamp(0, 255) * 256).round().numpy().astype(np.uint16)
    r = torch.from_numpy(q.astype(np.float32)) / 256.0 - 1
    m = r > 0
    r[~m] = -1
    r[m & (r > mx)] = mx
    r[m & (r < mn)] = mn
    return r, m


@pytest.mark.parametrize("b,h,w", [(1, 128, 256), (2, 64, 128)])
def test_composite_matches_composite_oracle(pf_lib, bg_shapes, b, h, w):
    pc_in = synthetic.make_pc_inputs(b=b, t=3, h=h, w=w, dist="R", seed=21)
    npin = {k: v.numpy() for k, v in pc_in.items()}
    npin["intrinsics_inv"] = torch.inverse(pc_in["intrinsics"]).numpy()
    npin["extrinsics_inv"] = torch.inverse(pc_in["extrinsics"]).numpy()
    segs, deps, masks = [], [], []
    for ind in range(3):
        r = pc_transform_oracle.predict(npin, only_this_ind=ind)
        d, m = disk_hop_np(r["depth"])
        segs.append(torch.from_numpy(r["seg"])); deps.append(d); masks.append(m)
    bg_in = {"seg": torch.stack(segs, 1).long(), "depth": torch.stack(deps, 1), "depth_mask": torch.stack(masks, 1)}
    sd = synthetic.make_bg_state_dict(bg_shapes, seed=21)
    q = bg_oracle.predict(sd, bg_in, None)["orig_size_logits"]
    sd["model.finalConv.bias"] = sd["model.finalConv.bias"] - q.mean((0, 2, 3))
    ref = bg_oracle.predict(sd, bg_in, None)

    bg = build_model(dict(bg_params(precision="fp32"), no_gpu=False)).eval()
    bg.load_state_dict(sd)
    pipe = BGForecastPipeline(bg)
    cu = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in npin.items()}
    out = pipe.forecast(cu)
    assert torch.equal(out["warped_seg"].cpu().long(), bg_in["seg"])              # bit-exact stage A
    assert torch.equal(out["warped_depth"].cpu(), bg_in["depth"])
    assert torch.equal(out["warped_mask"].cpu().bool(), bg_in["depth_mask"])
    check_against(out, ref, rel_tol=1e-4)


def test_pipelined_forecaster_equals_direct(pf_lib, bg_shapes):
    """Host-buffer pipelined front end (bench.py's e2e path) returns the same label maps as forecast()."""
    from panoptic_forecasting_b200.pipeline import PipelinedForecaster
    sd = synthetic.make_bg_state_dict(bg_shapes, seed=3)
    bg = build_model(dict(bg_params(return_logits=False, seg_dtype="uint8"), no_gpu=False)).eval()
    bg.load_state_dict(sd)
    pipe = BGForecastPipeline(bg)
    sets = []
    for seed in range(3):
        d = synthetic.make_pc_inputs(b=2, t=3, h=64, w=128, dist="R", seed=seed)
        d["intrinsics_inv"] = torch.inverse(d["intrinsics"]).contiguous()
        d["extrinsics_inv"] = torch.inverse(d["extrinsics"]).contiguous()
        sets.append({k: v.pin_memory() for k, v in d.items()})
    direct = [pipe.forecast({k: v.cuda() for k, v in s.items()})["seg"].cpu() for s in sets]
    pf = PipelinedForecaster(pipe, depth=2)
    got = []
    for i in range(5):
        pf.submit(sets[i % 3])
        if i >= 1:
            got.append(pf.collect().clone())
    got.append(pf.collect().clone())
    for i in range(5):
        assert torch.equal(got[i], direct[i % 3])


def test_forecast_panoptic_matches_merge_oracle(pf_lib, bg_shapes):
    """bg forecast -> fg merge on the device (BASELINE config 5's per-item work): the panoptic map equals the merge
    oracle applied to the bg path's own label map."""
    from oracle import panoptic_merge_oracle as merge_oracle
    h, w = 128, 256
    sd = synthetic.make_bg_state_dict(bg_shapes, seed=6)
    p = bg_params(return_logits=False, seg_dtype="uint8")
    p["model"]["final_h"], p["model"]["final_w"] = h, w
    bg = build_model(dict(p, no_gpu=False)).eval()
    bg.load_state_dict(sd)
    pipe = BGForecastPipeline(bg)
    d = synthetic.make_pc_inputs(b=2, t=3, h=h, w=w, dist="R", seed=2)
    case = synthetic.make_merge_inputs(2, (5, 3), h, w, seed=2)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    probs = [torch.sigmoid(torch.from_numpy(l)) for l in case["mask_logits"]]
    out = pipe.forecast_panoptic({k: v.cuda() for k, v in d.items()}, [m.cuda() for m in probs],
                                 [t(x) for x in case["bboxes"]], [t(c) for c in case["classes"]],
                                 [t(x) for x in case["depths"]])
    assert out["seg"].dtype == torch.uint8 and out["panoptic"].dtype == torch.int64
    seg = out["seg"].cpu().numpy()
    for i in range(2):
        ref = merge_oracle.merge(seg[i].astype(np.int64), probs[i].numpy(), case["bboxes"][i], case["classes"][i],
                                 case["depths"][i])
        assert np.array_equal(out["panoptic"][i].cpu().numpy(), ref)
    assert (out["panoptic"] >= 11000).any() and (out["panoptic"] < 11).any()
