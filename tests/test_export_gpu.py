"""GPU: the reference's OWN export loop (experiments/export_cityscapes_segmentation_results.py:53-127, imported
unmodified) driven through the drop-in overlay -- `from panoptic_forecasting.models import build_model` resolves to
the B200 BGModel -- on synthetic items, and this repo's sharded/pipelined `export_results` on the same items: the two
PNG sets are byte-identical and equal the model's direct predictions.  Runs in a subprocess (overlay on PYTHONPATH);
needs the reference tree (baseline/_ref on the GPU box, /root/reference in the build container)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import hashlib, os, sys, types
cs, h, l = (types.ModuleType(n) for n in ("cityscapesscripts", "cityscapesscripts.helpers", "cityscapesscripts.helpers.labels"))
l.labels, l.id2label, l.trainId2label = [], {}, {}
cs.helpers, h.labels = h, l
sys.modules.update({"cityscapesscripts": cs, "cityscapesscripts.helpers": h, "cityscapesscripts.helpers.labels": l,
                    "h5py": types.ModuleType("h5py")})
import numpy as np, torch
from PIL import Image
import panoptic_forecasting.experiments.export_cityscapes_segmentation_results as E      # the reference's script module
from panoptic_forecasting.models import build_model                                     # ... whose builder is the overlay's
from panoptic_forecasting_b200 import synthetic
from panoptic_forecasting_b200.export import export_results
assert E.build_model is build_model and "dropin" in sys.modules["panoptic_forecasting.models"].__file__
tmp, H, W, N = sys.argv[1], 128, 256, 5

class Items(torch.utils.data.Dataset):
    split = "val"
    def __len__(self): return N
    def __getitem__(self, i):
        d = synthetic.make_bg_inputs(1, 3, H, W, seed=i)
        return {"inputs": {k: v[0] for k, v in d.items()}, "labels": {},
                "meta": {"city": ["ulm", "bonn"][i % 2], "seq": "%06d" % (3 * i), "frame": i, "target_frame": 19 + i}}

params = {"task": "bg", "no_gpu": False, "load_best_model": False, "load_model": None,
          "data": {"num_classes": 11, "min_depth": 0.1, "max_depth": 200},
          "model": {"num_inputs": 3, "use_depth_inps": True, "convert2onehot": True, "final_h": H, "final_w": W},
          "training": {"batch_size": 2, "num_data_workers": 0}, "no_convert": True, "convert_to_trainid": False,
          "viz": False, "is_img": False, "save_depth": False, "save_disp_as_png": False, "save_depth_as_png": False,
          "disp_factor": None, "export_name": "exp"}
model = build_model(params).eval()
model.load_state_dict(synthetic.make_bg_state_dict({k: v.cpu() for k, v in model.state_dict().items()}, seed=4))
E.export_results(model, Items(), "val", dict(params, working_dir=os.path.join(tmp, "ref_loop")))     # reference loop
w, s, f = export_results(model, Items(), "val", dict(params, working_dir=os.path.join(tmp, "ours")))  # this repo's
assert (w, s, f) == (N, 0, 0)

def digest(base):
    return {os.path.relpath(os.path.join(r, x), base): hashlib.sha1(open(os.path.join(r, x), "rb").read()).hexdigest()
            for r, _, fs in os.walk(base) for x in fs}
a, b = digest(os.path.join(tmp, "ref_loop", "exp", "val")), digest(os.path.join(tmp, "ours", "exp", "val"))
assert len(a) == N and a == b, (sorted(a), sorted(b))
it = Items()[3]
direct = model.predict({k: v[None].cuda() for k, v in it["inputs"].items()}, {})["seg"][0].cpu().numpy().astype(np.uint8)
got = np.array(Image.open(os.path.join(tmp, "ours", "exp", "val", "bonn", "bonn_000009_000022_gtFine_labelIds.png")))
assert np.array_equal(got, direct) and len(np.unique(direct)) > 1
print("EXPORT_OK")
'''


def reference_root():
    for cand in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "panoptic_forecasting")):
            return cand
    return None


def test_reference_export_loop_runs_against_the_overlay_and_matches_our_driver(pf_lib, tmp_path):
    ref = reference_root()
    if ref is None:
        pytest.skip("reference tree not available")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT, ref]), PF_REFERENCE_ROOT=ref)
    r = subprocess.run([sys.executable, "-c", SCRIPT, str(tmp_path)], env=env, cwd=str(tmp_path), capture_output=True,
                       text=True, timeout=600)
    assert "EXPORT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
