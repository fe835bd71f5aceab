"""CPU: the drop-in overlay `dropin/panoptic_forecasting` -- with it ahead of the reference on PYTHONPATH,
`from panoptic_forecasting.models import build_model` (experiments/export_cityscapes_segmentation_results.py:16,
scripts/bg/run_export_bg_val.sh) builds the B200 bg / pc_transform models while every other reference module (data,
utils.config, training.train_utils, the fg / odom models) is still the reference's own file.  Runs in a subprocess
(the overlay and oracle/ref_loader.py must not share a process).  Needs the reference tree (/root/reference in the
build container, baseline/_ref on the GPU box); skipped where neither exists."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBE = r'''
import os, sys, types
cs, h, l = (types.ModuleType(n) for n in ("cityscapesscripts", "cityscapesscripts.helpers", "cityscapesscripts.helpers.labels"))
l.labels, l.id2label, l.trainId2label = [], {}, {}
cs.helpers, h.labels = h, l
sys.modules.update({"cityscapesscripts": cs, "cityscapesscripts.helpers": h, "cityscapesscripts.helpers.labels": l,
                    "h5py": types.ModuleType("h5py")})
from panoptic_forecasting.models import build_model                       # the reference's import line, verbatim
import panoptic_forecasting.models as M
import panoptic_forecasting.utils.config as cfg
import panoptic_forecasting.training.train_utils as tu
from panoptic_forecasting.data import build_dataset
ref = os.path.realpath(sys.argv[1])
assert "dropin" in M.__file__, M.__file__
for mod in (cfg, tu, sys.modules["panoptic_forecasting.data"]):
    assert os.path.realpath(mod.__file__).startswith(ref), mod.__file__
base = {"no_gpu": True, "load_best_model": False, "load_model": None}
bg = build_model(dict(base, task="bg", data={"num_classes": 11, "min_depth": 0.1, "max_depth": 200},
                      model={"num_inputs": 3, "use_depth_inps": True, "convert2onehot": True}))
assert type(bg).__module__ == "panoptic_forecasting_b200.models.bg_model" and len(bg.state_dict()) == 418
pc = build_model(dict(base, task="pc_transform", data={}, model={"only_this_ind": 1}))
assert type(pc).__module__ == "panoptic_forecasting_b200.models.pc_transform_model" and pc.ind == 1
od = build_model(dict(base, task="odom", data={}, model={"predict_type": "direct", "rnn_hidden": 8, "loss_fn": "mse"}))
assert os.path.realpath(sys.modules[type(od).__module__].__file__).startswith(ref)       # reference's own OdomModel
print("DROPIN_OK")
'''


def reference_root():
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "panoptic_forecasting")):
            return cand
    return None


def test_overlay_resolves_builder_to_b200_and_everything_else_to_the_reference():
    ref = reference_root()
    if ref is None:
        pytest.skip("reference tree not available")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT, ref]), PF_REFERENCE_ROOT=ref)
    r = subprocess.run([sys.executable, "-c", PROBE, ref], env=env, cwd="/tmp", capture_output=True, text=True, timeout=300)
    assert "DROPIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
