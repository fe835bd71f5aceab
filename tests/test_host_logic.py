"""CPU: host-side mirror of the reference interface (builder dispatch, state_dict keys, error
behaviour) and the synthetic input generators."""
import os

import numpy as np
import pytest
import torch

from conftest import bg_params, pc_params
from oracle import ref_loader
from panoptic_forecasting_b200 import _lib, synthetic
from panoptic_forecasting_b200.models import BGModel, PCTransformModel, build_model


def test_build_model_dispatch_and_errors(pf_lib):
    assert isinstance(build_model(bg_params()), BGModel)
    assert isinstance(build_model(pc_params(0)), PCTransformModel)
    with pytest.raises(ValueError):
        build_model({"task": "nope", "no_gpu": True, "load_best_model": False, "load_model": None})
    with pytest.raises(ValueError):
        build_model({"task": "fg", "no_gpu": True, "load_best_model": False, "load_model": None})


def test_state_dict_surface(pf_lib, tmp_path):
    m = build_model(bg_params(64, 128)).eval()
    sd = m.state_dict()
    assert len(sd) == 418
    assert sd["model.base.0.conv.weight"].shape == (16, 36, 3, 3)
    assert sd["model.finalConv.weight"].shape == (11, 48, 1, 1)
    assert sd["model.conv1x1_up.0.conv.weight"].shape == (267, 534, 1, 1)
    assert sd["model.base.16.layers.7.conv.weight"].shape == (158, 402, 3, 3)
    # save / load round trip through the reference's BaseModel surface (base_model.py:19-23)
    new = synthetic.make_bg_state_dict(sd, seed=3)
    m.load_state_dict(new)
    p = str(tmp_path / "ckpt.pt")
    m.save(p)
    m2 = build_model(dict(bg_params(64, 128), load_model=p))
    for k, v in m2.state_dict().items():
        assert torch.equal(v, new[k]), k
    with pytest.raises(RuntimeError):
        m.load_state_dict({"depth_mean": torch.zeros(1)})       # strict, like the reference


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree only exists in the build container")
def test_state_dict_keys_equal_reference(pf_lib):
    import warnings
    warnings.filterwarnings("ignore")
    ref = ref_loader.load_reference().build_model(ref_loader.ref_bg_params(64, 128))
    mine = build_model(bg_params(64, 128))
    a, b = ref.state_dict(), mine.state_dict()
    assert set(a) == set(b)
    for k in a:
        assert a[k].shape == b[k].shape, k
    mine.load_state_dict(a)          # a reference checkpoint loads strictly


def test_cpu_tensors_fail_loudly(pf_lib):
    m = build_model(bg_params()).eval()
    x = synthetic.make_bg_inputs(1, 3, 64, 64)
    with pytest.raises(_lib.PFError):
        m.predict(x, {})
    pc = build_model(pc_params(0))
    with pytest.raises(_lib.PFError):
        pc.predict(synthetic.make_pc_inputs(1, 3, 8, 8), {})


def test_dense_mode_and_loss_host_logic(pf_lib):
    """`convert2onehot: False` builds the same parameter tree; CPU tensors, .train() and a label tensor in dense mode are
    refused before any CUDA call (no CPU fallback, no silent reinterpretation)."""
    p = bg_params()
    p["model"] = dict(p["model"], convert2onehot=False)
    md = build_model(p).eval()
    assert set(md.state_dict()) == set(build_model(bg_params()).state_dict())
    x = synthetic.make_bg_dense_inputs(1, 3, 64, 64)
    assert x["seg"].shape == (1, 3, 11, 64, 64) and abs(x["seg"].sum(2) - 1).max() < 1e-5
    with pytest.raises(_lib.PFError):
        md.predict(x, {})
    target = synthetic.make_loss_target(x["seg"].argmax(2)[:, 0])
    assert (target[:, :5] == 255).all() and target.dtype.is_floating_point is False
    with pytest.raises(_lib.PFError):
        md.loss(x, {"seg": target})
    with pytest.raises(NotImplementedError):
        md.train().loss(x, {"seg": target})


def test_synthetic_inputs_are_seeded_and_shaped():
    a = synthetic.make_pc_inputs(b=2, t=3, h=32, w=64, dist="R", seed=7)
    b = synthetic.make_pc_inputs(b=2, t=3, h=32, w=64, dist="R", seed=7)
    for k in a:
        assert torch.equal(a[k], b[k])
    assert a["depth"].shape == (2, 3, 32, 64) and a["seg"].dtype == torch.uint8
    assert a["depth_mask"].dtype == torch.bool and a["target_T"].shape == (2, 3, 4, 4)
    frac = 1.0 - a["depth_mask"].float().mean().item()
    assert 0.05 < frac < 0.4
    T = a["target_T"][0, 0].numpy()
    assert abs(np.linalg.det(T[:3, :3]) - 1) < 1e-4 and T[0, 3] < -5.0    # ~15 steps of ~0.6 m forward


def test_ego_transforms_match_reference_formulas():
    """ego.target_T_from_odometry == product of the reference's unicycle steps (data_utils.py:117-165,
    restated in synthetic.vehicle_now_T_prev with np.linalg.inv as the reference does)."""
    from panoptic_forecasting_b200 import ego
    rng = np.random.default_rng(3)
    speed = rng.uniform(0, 20, size=(4, 15))
    yaw = rng.normal(0, 0.05, size=(4, 15))
    yaw[0, :3] = 0.0001            # straight-line branch
    dt = rng.uniform(0.05, 0.07, size=(4, 15))
    got = ego.target_T_from_odometry(torch.from_numpy(speed), torch.from_numpy(yaw), torch.from_numpy(dt)).numpy()
    for b in range(4):
        T = np.eye(4)
        for k in range(15):
            T = synthetic.vehicle_now_T_prev(speed[b, k], yaw[b, k], dt[b, k]) @ T
        assert np.abs(got[b] - T.astype(np.float32)).max() <= 1e-5
