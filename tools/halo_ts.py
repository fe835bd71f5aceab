"""Dev tool (GPU box): phase timestamps / wait accounting of the halo kernel on full-size layers.
Needs a library built with the instrumentation: PF_HALO_DBG=1 python -m panoptic_forecasting_b200.build --force
PF_HALO_TS=1 python tools/halo_ts.py [layer indices...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from conftest import bg_params
from panoptic_forecasting_b200 import _lib, synthetic
from panoptic_forecasting_b200.models import build_model

L = _lib.lib()
m = build_model(dict(bg_params(precision="tc"), no_gpu=False)).eval()
sd = synthetic.make_bg_state_dict({k: v.cpu() for k, v in m.state_dict().items()}, seed=2)
m.load_state_dict(sd)
m._upload(torch.device("cuda", 0))
info = _lib.ConvInfo()
shapes = {1: (512, 1024), 3: (256, 512), 7: (256, 512), 8: (256, 512), 13: (128, 256), 23: (64, 128), 43: (16, 32), 76: (256, 512), 72: (256, 512), 66: (128, 256), 56: (64, 128), 14: (128, 256), 70: (128, 256), 68: (128, 256)}
for i in [int(a) for a in sys.argv[1:]] or sorted(shapes):
    L.pf_bgnet_conv_info(m._net, i, C.byref(info))
    H, W = shapes.get(i, (256, 512))
    B = int(os.environ.get("PF_TS_BATCH", "2"))
    x = torch.randn(B, info.cin, H, W, device="cuda").relu()
    y = torch.empty((B, info.cout, H // info.stride, W // info.stride), device="cuda")
    rc = L.pf_bgnet_debug_conv(m._net, i, x.data_ptr(), B, H, W, y.data_ptr(), None)
    assert rc == 0, L.pf_last_error()
