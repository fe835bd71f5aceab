"""Drop-in for ``panoptic_forecasting.models`` on the bg-forecast hot path
(reference panoptic_forecasting/models/__init__.py:16-41): same ``build_model(params)``
dispatch, device placement and checkpoint loading; tasks outside the hot path are not built here.
"""
import os

import torch

from .bg_model import BGModel
from .pc_transform_model import PCTransformModel


def build_model(params):
    task = params['task']
    print("Building model for task: ", task)
    if task == 'bg':
        model = BGModel(params)
    elif task == 'pc_transform':
        model = PCTransformModel(params)
    elif task in ('fg', 'odom'):
        raise ValueError('task %s is outside the B200 hot path; build it with the reference package' % task)
    else:
        raise ValueError('task not recognized: ', task)
    if not params['no_gpu']:
        if 'gpu' in params:
            device = torch.device('cuda')
            model.to(device)
        else:
            model.cuda()
    if params['load_best_model']:
        path = os.path.join(params['working_dir'], 'best_model')
        model.load(path)
    elif params['load_model']:
        print("LOADING MODEL FROM SPECIFIED PATH")
        model.load(params['load_model'])
    return model
