"""CPU: the oracle restatements against the committed golden vectors (generated from the
unmodified reference by tests/golden/make_golden.py) and, when /root/reference is present
(build container only), against the live reference."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import bg_oracle, pc_transform_oracle, ref_loader
from panoptic_forecasting_b200 import synthetic

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_pc_case(path):
    z = np.load(path)
    inp = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    ind = int(z["only_this_ind"])
    return inp, (None if ind < 0 else ind), bool(int(z["is_img"])), z


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "pc_*.npz"))), ids=os.path.basename)
def test_pc_oracle_matches_golden_bit_exact(path):
    inp, ind, is_img, z = load_pc_case(path)
    out = pc_transform_oracle.predict(inp, only_this_ind=ind, is_img=is_img)
    assert np.array_equal(out["seg"], z["out_seg"])
    assert np.array_equal(out["depth"].view(np.uint32), z["out_depth"].view(np.uint32))
    assert np.array_equal(out["result2d"], z["out_result2d"].astype(np.int64))


def golden_bg_inputs(z, shapes):
    h, w, seed = int(z["h"]), int(z["w"]), int(z["seed"])
    sd = synthetic.make_bg_state_dict(shapes, seed=seed)
    sd["model.finalConv.bias"] = sd["model.finalConv.bias"] - torch.from_numpy(z["bias_shift"])
    if str(z["mode"]) == "pc":
        pc = synthetic.make_pc_inputs(1, 3, h, w, "R", seed=seed)
        inp = {"seg": pc["seg"].long(), "depth": pc["depth"].clamp(0.1, 200), "depth_mask": pc["depth_mask"]}
    else:
        inp = synthetic.make_bg_inputs(1, 3, h, w, seed=seed)
    return sd, inp, (int(z["fh"]), int(z["fw"]))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "bg_*.npz"))), ids=os.path.basename)
def test_bg_oracle_matches_golden(path, bg_shapes):
    z = np.load(path)
    sd, inp, final = golden_bg_inputs(z, bg_shapes)
    out = bg_oracle.predict(sd, inp, final)
    scale = np.abs(z["out_quarter"]).max()
    # same torch operators as the reference -> agreement to rounding noise of the CPU kernels
    assert np.abs(out["orig_size_logits"].numpy() - z["out_quarter"]).max() <= 1e-5 * scale
    assert np.abs(out["logits"].numpy()[:, :, ::7, ::5] - z["out_logits_sample"]).max() <= 1e-5 * scale
    assert (out["seg"].numpy() != z["out_seg"]).mean() <= 1e-3


def golden_dense_inputs(z, shapes, dense):
    """inputs of tests/golden/make_golden_dense.py's two cases, regenerated from the stored seed"""
    h, w, seed = int(z["h"]), int(z["w"]), int(z["seed"])
    sd = synthetic.make_bg_state_dict(shapes, seed=seed)
    sd["model.finalConv.bias"] = sd["model.finalConv.bias"] - torch.from_numpy(z["bias_shift"])
    inp = synthetic.make_bg_dense_inputs(2, 3, h, w, seed=seed) if dense else synthetic.make_bg_inputs(2, 3, h, w, seed=seed)
    target = synthetic.make_loss_target(torch.from_numpy(z["out_seg"].astype(np.int64)), seed=seed)
    return sd, inp, target


def test_bg_oracle_dense_planes_and_loss_match_golden(bg_shapes):
    """`convert2onehot` off and BGModel.loss (bg_model.py:61-69,73-89) against the unmodified reference's outputs."""
    z = np.load(os.path.join(GOLD, "dense_soft64.npz"))
    sd, inp, target = golden_dense_inputs(z, bg_shapes, dense=True)
    out = bg_oracle.predict_dense(sd, inp, None)
    scale = np.abs(z["out_quarter"]).max()
    assert np.abs(out["orig_size_logits"].numpy() - z["out_quarter"]).max() <= 1e-5 * scale
    assert np.abs(out["logits"].numpy()[:, :, ::7, ::5] - z["out_logits_sample"]).max() <= 1e-5 * scale
    assert (out["seg"].numpy() != z["out_seg"]).mean() <= 1e-3
    ls = bg_oracle.loss(sd, inp, {"seg": target}, dense=True)
    assert abs(ls["loss"].item() - float(z["loss"])) <= 1e-5 * float(z["loss"])
    assert abs(ls["accuracy"].item() - float(z["accuracy"])) <= 1e-4
    z = np.load(os.path.join(GOLD, "loss_iid64.npz"))
    sd, inp, target = golden_dense_inputs(z, bg_shapes, dense=False)
    ls = bg_oracle.loss(sd, inp, {"seg": target})
    assert abs(ls["loss"].item() - float(z["loss"])) <= 1e-5 * float(z["loss"])
    assert abs(ls["accuracy"].item() - float(z["accuracy"])) <= 1e-4


def test_scatter_tie_rule_known_answer():
    """Two sources land in the same cell with equal depth: the lower flattened source index wins
    (torch_scatter CPU rule); invalid-only cells get label 0 and depth max+1; untouched cells -1."""
    H, W = 4, 8
    K = np.eye(3, dtype=np.float32)[None]
    E = np.eye(4, dtype=np.float32)[None]
    T = np.eye(4, dtype=np.float32)[None, None].repeat(2, 1)
    depth = np.full((1, 2, H, W), 5.0, np.float32)
    mask = np.ones((1, 2, H, W), bool)
    seg = np.zeros((1, 2, H, W), np.uint8)
    seg[0, 0] = 3
    seg[0, 1] = 7          # identical geometry in frame 1: ties everywhere -> frame 0 must win
    mask[0, :, 1, 2] = False
    out = pc_transform_oracle.predict({"intrinsics": K, "extrinsics": E, "depth": depth, "depth_mask": mask,
                                       "target_T": T, "seg": seg})
    # identity camera: pixel (u,v) with depth 5 projects to (u/5*... ) -> K = I means u' = u*d/d = u
    assert out["seg"][0, 0, 0] == 3 and out["seg"][0, 3, 7] == 3
    assert out["seg"][0, 1, 2] == 0 and out["depth"][0, 1, 2] == np.float32(6.0)
    assert (out["depth"][0][out["seg"][0] == 3] == 5.0).all()


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree only exists in the build container")
def test_oracles_match_live_reference():
    import warnings
    warnings.filterwarnings("ignore")
    m = ref_loader.load_reference()
    inp = synthetic.make_pc_inputs(b=2, t=3, h=96, w=160, dist="U", seed=11)
    for ind in (1, None):
        ref = m.build_model(ref_loader.ref_pc_params(ind)).predict({k: v.clone() for k, v in inp.items()}, {})
        npin = {k: v.numpy() for k, v in inp.items()}
        npin["intrinsics_inv"] = torch.inverse(inp["intrinsics"]).numpy()
        npin["extrinsics_inv"] = torch.inverse(inp["extrinsics"]).numpy()
        out = pc_transform_oracle.predict(npin, only_this_ind=ind)
        assert np.array_equal(out["seg"], ref["seg"].numpy())
        assert np.array_equal(out["depth"].view(np.uint32), ref["depth"].numpy().view(np.uint32))
        assert np.array_equal(out["result2d"], ref["result2d"].numpy())
    bg = m.build_model(ref_loader.ref_bg_params(128, 256)).eval()
    sd = synthetic.make_bg_state_dict(bg.state_dict(), seed=5)
    bg.load_state_dict(sd)
    x = synthetic.make_bg_inputs(1, 3, 64, 128, seed=5)
    with torch.no_grad():
        ref = bg.predict({k: v.clone() for k, v in x.items()}, {})
    out = bg_oracle.predict(sd, x, (128, 256))
    assert torch.equal(out["seg"], ref["seg"])
    assert (out["logits"] - ref["logits"]).abs().max() <= 1e-5 * ref["logits"].abs().max()


def test_cpu_port_matches_numpy_oracle():
    """The torch-CPU baseline port (bench.py's cpu_baseline / --impl reference) agrees bit for bit
    with the numpy oracle on labels and depths."""
    from oracle import cpu_port
    inp = synthetic.make_pc_inputs(b=2, t=3, h=64, w=96, dist="R", seed=8)
    npin = {k: v.numpy() for k, v in inp.items()}
    npin["intrinsics_inv"] = torch.inverse(inp["intrinsics"]).numpy()
    npin["extrinsics_inv"] = torch.inverse(inp["extrinsics"]).numpy()
    for ind in (0, None):
        a = cpu_port.pc_predict(inp, only_this_ind=ind)
        b = pc_transform_oracle.predict(npin, only_this_ind=ind)
        assert np.array_equal(a["seg"].numpy(), b["seg"])
        assert np.array_equal(a["depth"].numpy().view(np.uint32), b["depth"].view(np.uint32))
