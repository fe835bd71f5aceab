"""Loader for the UNMODIFIED reference package (test infrastructure only).

Imports the reference's ``panoptic_forecasting`` either from ``/root/reference`` (read-only, only
present in the build container) or from ``baseline/_ref`` (the offline ``pip install --target``
of the same tree: git-ignored, but it travels to the GPU box), so that the oracle restatement in
this directory can be pinned against the reference's own code, ``tests/golden/make_golden.py`` can
generate committed fixtures, and ``bench.py --impl reference`` can time the reference itself.

Two shims are needed:
  * empty stub modules for ``cityscapesscripts`` / ``h5py`` (import-time only
    dependencies of ``panoptic_forecasting.data``),
  * ``torch_scatter.scatter_min`` (pytorch_scatter 2.0.5, un-vendored): restated
    with its CPU tie rule (first strictly-smaller update wins => lowest source
    index; untouched cells: out = 0, arg = src.size(dim)).
    Call site: panoptic_forecasting/models/pc_transform/pc_transform_model.py:118-119.

Nothing in the product package imports this file.
"""
import os
import sys
import types

import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    for cand in (os.environ.get("PF_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "panoptic_forecasting")):
            return cand
    return os.environ.get("PF_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    B, E = src.shape
    mn = torch.full((B, dim_size), float("inf"), dtype=src.dtype, device=src.device)
    mn.scatter_reduce_(1, index, src, "amin", include_self=True)
    e = torch.arange(E, device=src.device).expand(B, E)
    cand = torch.where(src == mn.gather(1, index), e, torch.full_like(e, E))
    arg = torch.full((B, dim_size), E, dtype=torch.long, device=src.device)
    arg.scatter_reduce_(1, index, cand, "amin", include_self=True)
    mn[arg == E] = 0
    return mn, arg


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "panoptic_forecasting"))


def load_reference():
    """Returns the reference's ``panoptic_forecasting.models`` module."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    cs, h, l = (types.ModuleType(n) for n in (
        "cityscapesscripts", "cityscapesscripts.helpers",
        "cityscapesscripts.helpers.labels"))
    l.labels, l.id2label, l.trainId2label = [], {}, {}
    cs.helpers, h.labels = h, l
    for name, mod in (("cityscapesscripts", cs), ("cityscapesscripts.helpers", h),
                      ("cityscapesscripts.helpers.labels", l),
                      ("h5py", types.ModuleType("h5py"))):
        sys.modules.setdefault(name, mod)
    ts = types.ModuleType("torch_scatter")
    ts.scatter_min = scatter_min
    sys.modules["torch_scatter"] = ts
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import panoptic_forecasting.models as ref_models
    return ref_models


def ref_bg_params(final_h=None, final_w=None):
    return {"task": "bg", "no_gpu": True, "load_best_model": False, "load_model": None,
            "data": {"num_classes": 11, "min_depth": 0.1, "max_depth": 200},
            "model": {"num_inputs": 3, "use_depth_inps": True, "convert2onehot": True,
                      "final_w": final_w, "final_h": final_h}}


def ref_pc_params(only_this_ind=None, is_img=None):
    return {"task": "pc_transform", "no_gpu": True, "load_best_model": False,
            "load_model": None, "data": {},
            "model": {"only_this_ind": only_this_ind, "is_img": is_img}}
