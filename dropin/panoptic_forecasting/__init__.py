"""Drop-in overlay of the reference's top-level package.

Put THIS directory's parent on PYTHONPATH ahead of the reference:

    PYTHONPATH=<repo>/dropin:<repo>:$PYTHONPATH bash scripts/bg/run_export_bg_val.sh     # script unchanged

`panoptic_forecasting.models.build_model` then builds the bg / pc_transform models of the B200 path
(panoptic_forecasting_b200, libpf_b200.so), while every other module of the reference --
`panoptic_forecasting.data`, `.utils.config`, `.training.train_utils`, `.experiments.*`, the fg and odom
models -- is still the reference's own file: this package only extends its search path with the reference's
package directory.  Reference import sites served: experiments/export_cityscapes_segmentation_results.py:15-19,
experiments/export_cityscapes_panoptic_results.py (same builder import), scripts/bg/run_export_bg_val.sh:16-22,
scripts/fg/run_fg_eval_panoptic.sh.

The reference is located through (first hit): $PF_REFERENCE_ROOT, any other `panoptic_forecasting` package
directory on sys.path (a `pip install` of the reference, or its checkout root), <repo>/baseline/_ref.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))


def _reference_package_dir():
    cands = []
    if os.environ.get("PF_REFERENCE_ROOT"):
        cands.append(os.environ["PF_REFERENCE_ROOT"])
    cands += [p or os.getcwd() for p in sys.path]
    cands.append(os.path.join(_REPO, "baseline", "_ref"))
    for root in cands:
        d = os.path.join(root, "panoptic_forecasting")
        if os.path.isfile(os.path.join(d, "models", "base_model.py")) and os.path.realpath(d) != os.path.realpath(_HERE):
            return d
    return None


REFERENCE_PACKAGE_DIR = _reference_package_dir()
if REFERENCE_PACKAGE_DIR is not None:
    __path__.append(REFERENCE_PACKAGE_DIR)          # submodules not overlaid here resolve to the reference's files
if _REPO not in sys.path:
    sys.path.append(_REPO)                          # panoptic_forecasting_b200
