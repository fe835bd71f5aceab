"""GPU parity: Stage B (pf_bgnet_* through BGModel / the C ABI) vs the oracle (fp32 torch CPU
restatement of the reference).  Tolerance (north_star): logits within 1e-3 relative
(max |diff| / max |ref|); label map exact except where the reference's own top-2 logit gap is
below the stated float tolerance (near-ties), which is checked pixel by pixel."""
import ctypes as C
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import bg_params
from oracle import bg_oracle
from panoptic_forecasting_b200 import _lib, synthetic
from panoptic_forecasting_b200.models import build_model
from test_oracle import golden_bg_inputs, golden_dense_inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REL_TOL = 1e-3


def gpu_model(sd, final=None, **b200):
    fh, fw = final if final is not None else (None, None)
    b200.setdefault("precision", "fp32")
    m = build_model(dict(bg_params(fh, fw, **b200), no_gpu=False)).eval()
    m.load_state_dict(sd)
    return m


def check_against(out, ref, rel_tol=REL_TOL):
    scale = ref["logits"].abs().max().item()
    err_full = (out["logits"].cpu() - ref["logits"]).abs().max().item() / scale
    err_q = (out["orig_size_logits"].cpu() - ref["orig_size_logits"]).abs().max().item() / scale
    assert err_full <= rel_tol and err_q <= rel_tol, (err_full, err_q)
    seg, rseg = out["seg"].cpu().long(), ref["seg"]
    mism = seg != rseg
    if mism.any():
        top2 = ref["logits"].topk(2, dim=1).values
        gap = (top2[:, 0] - top2[:, 1])[mism]
        assert (gap <= 2 * rel_tol * scale).all(), "argmax differs where the reference is not a near-tie"
    return err_full, int(mism.sum())


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "bg_*.npz"))), ids=os.path.basename)
def test_golden_vectors(pf_lib, bg_shapes, path):
    z = np.load(path)
    sd, inp, final = golden_bg_inputs(z, bg_shapes)
    m = gpu_model(sd, final)
    out = m.predict({k: v.cuda() for k, v in inp.items()}, {})
    scale = np.abs(z["out_quarter"]).max()
    assert np.abs(out["orig_size_logits"].cpu().numpy() - z["out_quarter"]).max() <= 1e-4 * scale
    assert np.abs(out["logits"].cpu().numpy()[:, :, ::7, ::5] - z["out_logits_sample"]).max() <= 1e-4 * scale
    assert (out["seg"].cpu().numpy() != z["out_seg"]).mean() <= 1e-3


@pytest.mark.parametrize("shape,final", [((1, 3, 64, 128), None), ((2, 3, 128, 192), (256, 384)), ((1, 3, 256, 512), (256, 512))])
def test_whole_net_fp32_path(pf_lib, bg_shapes, shape, final):
    b, t, h, w = shape
    sd = synthetic.make_bg_state_dict(bg_shapes, seed=h)
    pc = synthetic.make_pc_inputs(b, 3, h, w, "R", seed=h)
    inp = {"seg": pc["seg"].long(), "depth": pc["depth"].clamp(0.1, 200), "depth_mask": pc["depth_mask"]}
    q = bg_oracle.predict(sd, inp, final)["orig_size_logits"]
    sd["model.finalConv.bias"] = sd["model.finalConv.bias"] - q.mean((0, 2, 3))     # every class wins somewhere
    ref = bg_oracle.predict(sd, inp, final)
    assert len(torch.unique(ref["seg"])) >= 8
    m = gpu_model(sd, final)
    out = m.predict({k: v.cuda() for k, v in inp.items()}, {})
    err, mism = check_against(out, ref, rel_tol=1e-4)
    assert mism <= 1e-4 * ref["seg"].numel()
    # forward() surface (bg_model.py:61-71)
    full = m(inp["seg"].cuda(), inp["depth"].cuda(), inp["depth_mask"].cuda())
    assert torch.equal(full, out["logits"])
    # uint8 labels + label-map-only export mode give the same label map
    m2 = gpu_model(sd, final, return_logits=False, seg_dtype="uint8")
    o2 = m2.predict({"seg": inp["seg"].to(torch.uint8).cuda(), "depth": inp["depth"].cuda(),
                     "depth_mask": inp["depth_mask"].cuda()}, {})
    assert set(o2) == {"seg"} and o2["seg"].dtype == torch.uint8
    assert torch.equal(o2["seg"].long(), out["seg"])


def test_labels_out_of_range_and_masked_depth(pf_lib, bg_shapes):
    """class ids >= 11 -> all-zero one-hot (bg_model.py:54-57); masked depth -> 0 (bg_model.py:68)."""
    sd = synthetic.make_bg_state_dict(bg_shapes, seed=1)
    inp = synthetic.make_bg_inputs(1, 3, 64, 64, seed=1)           # labels iid 0..18, mask = depth > 5
    inp["seg"][0, 0, :8] = 255
    ref = bg_oracle.predict(sd, {k: v.clone() for k, v in inp.items()}, None)
    out = gpu_model(sd).predict({k: v.cuda() for k, v in inp.items()}, {})
    check_against(out, ref, rel_tol=1e-4)


def test_loss_eval_mode(pf_lib, bg_shapes):
    """BGModel.loss (bg_model.py:73-89) in .eval() mode against the numbers the unmodified reference produced
    (tests/golden/loss_iid64.npz) and against the oracle; .train() refuses."""
    z = np.load(os.path.join(GOLD, "loss_iid64.npz"))
    sd, inp, target = golden_dense_inputs(z, bg_shapes, dense=False)
    want = bg_oracle.loss(sd, inp, {"seg": target})
    for precision in ("fp32", "tc"):
        m = gpu_model(sd, precision=precision)
        got = m.loss({k: v.cuda() for k, v in inp.items()}, {"seg": target.cuda()})
        for ref_loss, ref_acc in ((want["loss"].item(), want["accuracy"].item()), (float(z["loss"]), float(z["accuracy"]))):
            assert abs(got["loss"].item() - ref_loss) <= 1e-3 * abs(ref_loss), (precision, got["loss"].item(), ref_loss)
            assert abs(got["accuracy"].item() - ref_acc) <= 1e-3, (precision, got["accuracy"].item(), ref_acc)
    with pytest.raises(NotImplementedError):
        m.train().loss({k: v.cuda() for k, v in inp.items()}, {"seg": target.cuda()})


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("tc", 3e-4)])
def test_dense_planes_input_mode(pf_lib, bg_shapes, precision, tol):
    """`convert2onehot: False` (bg_model.py:61-65): float per-class planes [b,t,C,H,W] through pf_bgnet_forward_dense,
    against the oracle and the unmodified reference's outputs (tests/golden/dense_soft64.npz).  Soft class scores, so
    every weight of the first layer matters; one-hot planes must reproduce the label path."""
    z = np.load(os.path.join(GOLD, "dense_soft64.npz"))
    sd, dense_in, target = golden_dense_inputs(z, bg_shapes, dense=True)
    ref = bg_oracle.predict_dense(sd, dense_in, None)
    p = dict(bg_params(None, None, precision=precision), no_gpu=False)
    p["model"] = dict(p["model"], convert2onehot=False)
    md = build_model(p).eval()
    md.load_state_dict(sd)
    out = md.predict({k: v.cuda() for k, v in dense_in.items()}, {})
    check_against(out, ref, rel_tol=tol)
    scale = np.abs(z["out_quarter"]).max()
    assert np.abs(out["orig_size_logits"].cpu().numpy() - z["out_quarter"]).max() <= tol * scale
    assert (out["seg"].cpu().numpy() != z["out_seg"]).mean() <= 1e-3
    ls = md.loss({k: v.cuda() for k, v in dense_in.items()}, {"seg": target.cuda()})
    assert abs(ls["loss"].item() - float(z["loss"])) <= 1e-3 * float(z["loss"])
    # one-hot planes of a label map == the label path of the same net (both accumulate the first layer in fp32; the
    # summation orders differ)
    inp = synthetic.make_bg_inputs(2, 3, 64, 128, seed=5)
    lab = inp["seg"].clamp(max=10)
    onehot = F.one_hot(lab.long(), 11).permute(0, 1, 4, 2, 3).float()
    o1 = md.predict({"seg": onehot.cuda(), "depth": inp["depth"].cuda(), "depth_mask": inp["depth_mask"].cuda()}, {})
    m = gpu_model(sd, precision=precision)
    o2 = m.predict({"seg": lab.cuda(), "depth": inp["depth"].cuda(), "depth_mask": inp["depth_mask"].cuda()}, {})
    assert (o1["logits"] - o2["logits"]).abs().max().item() <= tol * o2["logits"].abs().max().item()
    with pytest.raises(ValueError):
        md.predict({"seg": lab.cuda(), "depth": inp["depth"].cuda(), "depth_mask": inp["depth_mask"].cuda()}, {})


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("tc", 3e-4)])
@pytest.mark.parametrize("h,w,final", [(68, 144, None), (100, 176, (200, 352)), (132, 80, None)])
def test_sizes_not_multiples_of_64(pf_lib, bg_shapes, precision, tol, h, w, final):
    """H % 4 == 0, W % 16 == 0: odd sizes at the pooled levels (AvgPool2d floors, hardnet.py:300; TransitionUp
    interpolates to the skip's size, hardnet.py:249-255), ragged tiles everywhere."""
    sd = synthetic.make_bg_state_dict(bg_shapes, seed=h)
    inp = synthetic.make_bg_inputs(2, 3, h, w, seed=h)
    q = bg_oracle.predict(sd, {k: v.clone() for k, v in inp.items()}, final)["orig_size_logits"]
    sd["model.finalConv.bias"] = sd["model.finalConv.bias"] - q.mean((0, 2, 3))
    ref = bg_oracle.predict(sd, {k: v.clone() for k, v in inp.items()}, final)
    out = gpu_model(sd, final, precision=precision).predict({k: v.cuda() for k, v in inp.items()}, {})
    check_against(out, ref, rel_tol=tol)
    with pytest.raises(ValueError):
        gpu_model(sd, precision=precision).predict({k: v[..., :66, :w].cuda() for k, v in inp.items()}, {})


def test_every_conv_layer_vs_torch(pf_lib, bg_shapes):
    """Each of the 69 ConvLayers + finalConv through pf_bgnet_debug_conv vs F.conv2d + BN + ReLU."""
    sd = synthetic.make_bg_state_dict(bg_shapes, seed=2)
    m = gpu_model(sd)
    m._upload(torch.device("cuda", torch.cuda.current_device()))
    n = pf_lib.pf_bgnet_num_convs(m._net)
    info = _lib.ConvInfo()
    g = torch.Generator().manual_seed(0)
    for i in range(1, n + 1):
        assert pf_lib.pf_bgnet_conv_info(m._net, i, C.byref(info)) == 0
        name = info.name.decode()
        H, W = (24, 40) if info.stride == 1 else (24, 48)
        x = torch.randn(2, info.cin, H, W, generator=g).relu()
        if i < n:
            ref = bg_oracle.conv_layer(sd, name, x, info.ksize, info.stride)
        else:
            ref = F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"])
        y = torch.empty(ref.shape, device="cuda")
        rc = pf_lib.pf_bgnet_debug_conv(m._net, i, x.cuda().data_ptr(), 2, H, W, y.data_ptr(), None)
        assert rc == 0, pf_lib.pf_last_error()
        err = (y.cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
        assert err <= 2e-5, (name, err)


def test_upsample_argmax_standalone(pf_lib):
    g = torch.Generator().manual_seed(3)
    q = torch.randn(2, 11, 16, 32, generator=g)
    for fh, fw in ((64, 128), (33, 70), (16, 32)):
        ref = F.interpolate(q, size=(fh, fw), mode="bilinear", align_corners=True)
        full = torch.empty(ref.shape, device="cuda")
        seg8 = torch.empty((2, fh, fw), dtype=torch.uint8, device="cuda")
        seg64 = torch.empty((2, fh, fw), dtype=torch.int64, device="cuda")
        rc = pf_lib.pf_upsample_argmax(q.cuda().data_ptr(), 2, 11, 16, 32, fh, fw, seg8.data_ptr(), seg64.data_ptr(),
                                       full.data_ptr(), None)
        assert rc == 0
        assert (full.cpu() - ref).abs().max() <= 1e-5
        assert torch.equal(seg64.cpu(), full.cpu().argmax(1))
        assert torch.equal(seg8.cpu().long(), seg64.cpu())
        assert (seg64.cpu() != ref.argmax(1)).float().mean() <= 1e-3


def test_full_size_properties(pf_lib, bg_shapes):
    """BASELINE size 1024x2048: the oracle takes ~1.5 s here, so compare directly once, plus
    batch-consistency (b=2 equals two b=1 calls) and determinism."""
    sd = synthetic.make_bg_state_dict(bg_shapes, seed=4)
    pc = synthetic.make_pc_inputs(2, 3, 1024, 2048, "R", seed=4)
    inp = {"seg": pc["seg"], "depth": pc["depth"].clamp(0.1, 200), "depth_mask": pc["depth_mask"]}
    m = gpu_model(sd, (1024, 2048), return_logits=True)
    cu = {k: v.cuda() for k, v in inp.items()}
    out2 = m.predict(cu, {})
    out_a = m.predict({k: v[:1] for k, v in cu.items()}, {})
    assert torch.equal(out2["seg"][:1], out_a["seg"]) and torch.equal(out2["logits"][:1], out_a["logits"])
    assert torch.equal(m.predict(cu, {})["seg"], out2["seg"])
    ref = bg_oracle.predict(sd, {k: v[:1].long() if k == "seg" else v[:1] for k, v in inp.items()}, (1024, 2048))
    check_against(out_a, ref, rel_tol=1e-4)


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_two_input_frames(pf_lib, precision):
    """num_inputs = 2 (a [16, 24, 3, 3] first conv): the first conv's runtime-frame-count instantiation and every
    downstream shape, against the oracle."""
    p = bg_params(precision=precision)
    p["model"]["num_inputs"] = 2
    shapes = {k: torch.zeros(v.shape) for k, v in build_model(dict(p)).state_dict().items()}
    assert tuple(shapes["model.base.0.conv.weight"].shape) == (16, 24, 3, 3)
    sd = synthetic.make_bg_state_dict(shapes, seed=5)
    pc = synthetic.make_pc_inputs(2, 2, 64, 192, "R", seed=5)
    inp = {"seg": pc["seg"].long(), "depth": pc["depth"].clamp(0.1, 200), "depth_mask": pc["depth_mask"]}
    ref = bg_oracle.predict(sd, inp, None)
    m = build_model(dict(p, no_gpu=False)).eval()
    m.load_state_dict(sd)
    out = m.predict({k: v.cuda() for k, v in inp.items()}, {})
    check_against(out, ref, rel_tol=REL_TOL if precision == "fp32" else 3e-4)
