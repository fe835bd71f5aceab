"""Overlay of panoptic_forecasting/models/__init__.py:16-41: same `build_model(params)`, but tasks `bg` and
`pc_transform` (the bg-forecast hot path) build the B200 models; `fg` and `odom` build the reference's own
classes, imported from the reference tree through this package's extended search path."""
import os

from panoptic_forecasting import REFERENCE_PACKAGE_DIR

if REFERENCE_PACKAGE_DIR is not None:
    __path__.append(os.path.join(REFERENCE_PACKAGE_DIR, "models"))      # .base_model, .fg, .odom, .bg, .pc_transform

from panoptic_forecasting_b200.models import build_model as _build_b200  # noqa: E402
from panoptic_forecasting_b200.models.bg_model import BGModel  # noqa: E402,F401
from panoptic_forecasting_b200.models.pc_transform_model import PCTransformModel  # noqa: E402,F401


def build_model(params):
    task = params['task']
    if task in ('bg', 'pc_transform'):
        return _build_b200(params)
    if REFERENCE_PACKAGE_DIR is None:
        raise ImportError("task %r is built by the reference package, which was not found "
                          "(set PF_REFERENCE_ROOT or install nianticlabs/panoptic-forecasting)" % task)
    import torch
    print("Building model for task: ", task)
    if task == 'fg':
        from panoptic_forecasting.models.fg.fg_model import FGModel
        model = FGModel(params)
    elif task == 'odom':
        from panoptic_forecasting.models.odom.odom_model import OdomModel
        model = OdomModel(params)
    else:
        raise ValueError('task not recognized: ', task)
    if not params['no_gpu']:
        if 'gpu' in params:
            model.to(torch.device('cuda'))
        else:
            model.cuda()
    if params['load_best_model']:
        model.load(os.path.join(params['working_dir'], 'best_model'))
    elif params['load_model']:
        print("LOADING MODEL FROM SPECIFIED PATH")
        model.load(params['load_model'])
    return model
