import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pf_lib():
    from panoptic_forecasting_b200 import build, _lib
    build.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def bg_shapes():
    """state_dict key -> shape of the bg model (taken from the product's parameter tree, which
    tests/test_host_logic.py checks against the reference's 418 keys)."""
    import torch
    from panoptic_forecasting_b200.models import build_model
    m = build_model(bg_params())
    return {k: torch.zeros(v.shape) for k, v in m.state_dict().items()}


def bg_params(final_h=None, final_w=None, **b200):
    return {"task": "bg", "no_gpu": True, "load_best_model": False, "load_model": None,
            "data": {"num_classes": 11, "min_depth": 0.1, "max_depth": 200},
            "model": {"num_inputs": 3, "use_depth_inps": True, "convert2onehot": True,
                      "final_w": final_w, "final_h": final_h, "b200": b200}}


def pc_params(only_this_ind=None, is_img=None, **extra):
    model = {"only_this_ind": only_this_ind, "is_img": is_img}
    model.update(extra)
    return {"task": "pc_transform", "no_gpu": True, "load_best_model": False, "load_model": None,
            "data": {}, "model": model}
