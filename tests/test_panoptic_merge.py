"""fg -> bg panoptic merge: the oracle against fixtures produced by the UNMODIFIED
reference (tests/golden/make_golden_merge.py), and the CUDA kernel against both."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import panoptic_merge_oracle as merge_oracle
from oracle import ref_loader
from panoptic_forecasting_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "merge_*.npz")))
H, W = 1024, 2048


def load_case(path):
    g = np.load(path)
    n_per = tuple(int(x) for x in np.atleast_1d(g["n_per_item"]))
    kw = dict(ulbr=bool(g["use_bbox_ulbr"]), zsort=bool(g["use_depth_sorting"]), with_depth=bool(g["with_bg_depth"]),
              with_mask=bool(g["with_bg_mask"]))
    case = synthetic.make_merge_inputs(len(n_per), n_per, H, W, seed=int(g["seed"]), use_bbox_ulbr=kw["ulbr"])
    case["mask_probs"] = [torch.sigmoid(torch.from_numpy(l)).numpy() for l in case["mask_logits"]]   # fg_model.py:541
    return case, kw, g["seg"].astype(np.int64)


def oracle_item(case, kw, i, h=None, w=None):
    return merge_oracle.merge(case["background"][i], case["mask_probs"][i], case["bboxes"][i], case["classes"][i],
                              case["depths"][i], bg_depth=case["bg_depth"][i] if kw["with_depth"] else None,
                              bg_depth_mask=case["bg_depth_mask"][i] if kw["with_mask"] else None,
                              use_depth_sorting=kw["zsort"], use_bbox_ulbr=kw["ulbr"])


def test_fixtures_present():
    assert len(CASES) == 3


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_oracle_matches_reference_fixture_exactly(path):
    case, kw, seg = load_case(path)
    for i in range(seg.shape[0]):
        out = oracle_item(case, kw, i)
        assert out.dtype == np.int64 and np.array_equal(out, seg[i])


def test_known_answers():
    """Hand-checkable cases: id rule, `>= 11 -> 255`, far-to-near painting, z-test against the background."""
    h, w = 8, 16
    bg = np.full((h, w), 3, dtype=np.int64)
    bg[:, :4] = 17
    one = np.ones((1, 4, 4), dtype=np.float32)
    masks = np.concatenate([one, one])
    boxes = np.array([[4, 0, 12, 8], [8, 0, 16, 8]], dtype=np.float32)
    out = merge_oracle.merge(bg, masks, boxes, np.array([2, 2]), np.array([10.0, 20.0], dtype=np.float32),
                             bg_depth=np.full((h, w), 15.0, dtype=np.float32), use_depth_sorting=True, use_bbox_ulbr=True)
    assert (out[:, :4] == 255).all()                       # background id 17 -> 255
    # instance 1 (depth 20, painted first, id 13000) is behind the background (15): never visible
    # instance 0 (depth 10, painted second, id 13001) wins its whole box
    assert (out[:, 4:12] == 13001).all() and (out[:, 12:] == 3).all()
    order, vals = merge_oracle.paint_order(np.array([2, 5, 2]), np.array([1.0, 3.0, 2.0], dtype=np.float32), True)
    assert list(order) == [1, 2, 0] and vals == [16000, 13000, 13001]
    # without depth sorting: index order, plain overwrite
    out2 = merge_oracle.merge(bg, masks, boxes, np.array([2, 2]), None, use_depth_sorting=False)
    assert (out2[:, 4:8] == 13000).all() and (out2[:, 8:] == 13001).all()


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree only exists in the build container")
def test_paste_mask_bit_exact_vs_live_reference():
    ref_loader.load_reference()
    from panoptic_forecasting.models.fg import model_utils
    rng = np.random.RandomState(11)
    for ulbr in (True, False):
        m = rng.rand(28, 28).astype(np.float32)
        bb = np.array([rng.uniform(-40, 300), rng.uniform(-30, 100), 0, 0], dtype=np.float32)
        bb[2], bb[3] = bb[0] + rng.uniform(30, 250), bb[1] + rng.uniform(20, 120)
        if not ulbr:
            bb = np.array([(bb[0] + bb[2]) / 2, (bb[1] + bb[3]) / 2, bb[2] - bb[0], bb[3] - bb[1]], dtype=np.float32)
        ref = model_utils.paste_mask(torch.from_numpy(m)[None, None], torch.from_numpy(bb)[None], 192, 384, ulbr)[0, 0].numpy()
        mine = merge_oracle.paste_mask(m, bb, 192, 384, ulbr)
        assert np.array_equal(ref.view(np.uint32), mine.view(np.uint32))


def test_abi_symbol_exported(pf_lib):
    assert hasattr(pf_lib, "pf_panoptic_merge") and hasattr(pf_lib, "pf_panoptic_paint_order")


# ---------------------------------------------------------------------------------------------- GPU
def run_cuda(case, kw, b):
    from panoptic_forecasting_b200 import panoptic
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    res = panoptic.merge_instances(
        [t(m) for m in case["mask_probs"]], [t(x) for x in case["bboxes"]], [t(c) for c in case["classes"]],
        [t(d) for d in case["depths"]], background=torch.stack([t(x) for x in case["background"]]),
        background_depths=torch.stack([t(x) for x in case["bg_depth"]]) if kw["with_depth"] else None,
        background_depth_masks=torch.stack([t(x) for x in case["bg_depth_mask"]]) if kw["with_mask"] else None,
        use_depth_sorting=kw["zsort"], use_bbox_ulbr=kw["ulbr"])
    return res["seg"].cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_cuda_merge_matches_reference_fixture_exactly(path):
    case, kw, seg = load_case(path)
    out = run_cuda(case, kw, seg.shape[0])
    assert out.dtype == np.int64 and out.shape == seg.shape
    assert np.array_equal(out, seg), int((out != seg).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,ulbr,zsort", [(7, (0, 3), True, True), (8, (70,), False, True), (9, (12, 1, 5), True, False)])
def test_cuda_merge_matches_oracle_ragged(seed, n, ulbr, zsort):
    """Empty items, more instances than one shared-memory chunk (64), cxcywh boxes, no depth sorting; 96 x 160 frames."""
    h, w = 96, 160
    case = synthetic.make_merge_inputs(len(n), n, h, w, seed=seed, use_bbox_ulbr=ulbr)
    case["mask_probs"] = [torch.sigmoid(torch.from_numpy(l)).numpy() for l in case["mask_logits"]]
    kw = dict(ulbr=ulbr, zsort=zsort, with_depth=zsort, with_mask=zsort)
    out = run_cuda(case, kw, len(n))
    for i in range(len(n)):
        ref = merge_oracle.merge(case["background"][i], case["mask_probs"][i], case["bboxes"][i], case["classes"][i],
                                 case["depths"][i], bg_depth=case["bg_depth"][i] if zsort else None,
                                 bg_depth_mask=case["bg_depth_mask"][i] if zsort else None, use_depth_sorting=zsort,
                                 use_bbox_ulbr=ulbr)
        assert np.array_equal(out[i], ref), (i, int((out[i] != ref).sum()))


@pytest.mark.gpu
def test_cuda_paint_order_matches_oracle():
    """Depth-descending stable order with ties, per-class running ids, ragged items incl. an empty one."""
    from panoptic_forecasting_b200 import panoptic
    dev = torch.device("cuda", torch.cuda.current_device())
    rng = np.random.RandomState(3)
    ns = (5, 0, 200, 1)
    classes = [rng.randint(0, 8, n).astype(np.int64) for n in ns]
    depths = [np.round(rng.uniform(5, 70, n), 0).astype(np.float32) for n in ns]       # rounding forces ties
    masks = [torch.zeros(n, 28, 28, device=dev) for n in ns]
    boxes = [torch.zeros(n, 4, device=dev) for n in ns]
    for zsort in (True, False):
        prep = panoptic.prepare_instances(masks, boxes, [torch.from_numpy(c).to(dev) for c in classes],
                                          [torch.from_numpy(d).to(dev) for d in depths] if zsort else None, zsort)
        vals, order, begin = prep[3].cpu().numpy(), prep[4].cpu().numpy(), prep[5].cpu().numpy()
        assert list(begin) == [0, 5, 5, 205, 206]
        for i, n in enumerate(ns):
            ref_order, ref_vals = merge_oracle.paint_order(classes[i], depths[i] if zsort else None, zsort)
            assert list(order[begin[i]:begin[i + 1]] - begin[i]) == [int(k) for k in ref_order]
            assert list(vals[begin[i]:begin[i + 1]]) == ref_vals


@pytest.mark.gpu
def test_cuda_merge_uint8_background_and_odd_width():
    """uint8 label map straight from the bg exporter; W not a multiple of 4 (scalar load/store path)."""
    from panoptic_forecasting_b200 import panoptic
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    for h, w in ((64, 256), (50, 131)):
        n = (6, 2)
        case = synthetic.make_merge_inputs(len(n), n, h, w, seed=21)
        probs = [torch.sigmoid(torch.from_numpy(l)).numpy() for l in case["mask_logits"]]
        bg = torch.stack([t(x) for x in case["background"]])
        kwargs = dict(pred_bboxes=[t(x) for x in case["bboxes"]], orig_classes=[t(c) for c in case["classes"]],
                      pred_depths=[t(d) for d in case["depths"]],
                      background_depths=torch.stack([t(x) for x in case["bg_depth"]]),
                      background_depth_masks=torch.stack([t(x) for x in case["bg_depth_mask"]]))
        a = panoptic.merge_instances([t(m) for m in probs], background=bg, **kwargs)["seg"]
        c = panoptic.merge_instances([t(m) for m in probs], background=bg.to(torch.uint8), **kwargs)["seg"]
        assert torch.equal(a, c)
        for i in range(len(n)):
            ref = merge_oracle.merge(case["background"][i], probs[i], case["bboxes"][i], case["classes"][i], case["depths"][i],
                                     bg_depth=case["bg_depth"][i], bg_depth_mask=case["bg_depth_mask"][i])
            assert np.array_equal(a[i].cpu().numpy(), ref), (h, w, i)


@pytest.mark.gpu
def test_cuda_merge_requires_cuda_tensors():
    from panoptic_forecasting_b200 import _lib, panoptic
    with pytest.raises(_lib.PFError):
        panoptic.merge_instances([torch.zeros(1, 28, 28)], [torch.zeros(1, 4)], [torch.zeros(1, dtype=torch.long)])
