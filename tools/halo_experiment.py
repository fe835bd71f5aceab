"""Dev experiment (GPU box): per-layer error of the halo tcgen05 kernel under descriptor variants.
env: PF_TC_HALO=1 PF_HALO_HX={10,16} PF_HALO_BASEOFF={0,1}"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.nn.functional as F

from conftest import bg_params
from oracle import bg_oracle
from panoptic_forecasting_b200 import _lib, synthetic
from panoptic_forecasting_b200.models import build_model

L = _lib.lib()
m = build_model(dict(bg_params(precision="tc"), no_gpu=False)).eval()
sd = synthetic.make_bg_state_dict({k: v.cpu() for k, v in m.state_dict().items()}, seed=2)
m.load_state_dict(sd)
m._upload(torch.device("cuda", 0))
n = L.pf_bgnet_num_convs(m._net)
info = _lib.ConvInfo()
g = torch.Generator().manual_seed(0)
bad = 0
worst = 0.0
for i in range(1, n + 1):
    L.pf_bgnet_conv_info(m._net, i, C.byref(info))
    name = info.name.decode()
    if info.ksize != 3 or info.stride != 1:
        continue
    H, W = 40, 24
    x = torch.randn(2, info.cin, H, W, generator=g).relu()
    ref = bg_oracle.conv_layer(sd, name, x, info.ksize, info.stride)
    y = torch.empty(ref.shape, device="cuda")
    rc = L.pf_bgnet_debug_conv(m._net, i, x.cuda().data_ptr(), 2, H, W, y.data_ptr(), None)
    if rc != 0:
        print("FAIL rc", rc, name, L.pf_last_error())
        sys.exit(1)
    err = (y.cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
    worst = max(worst, err)
    if err > 1e-4:
        bad += 1
        if bad <= 6:
            print("  BAD %-36s %4d->%-4d err %.3e" % (name, info.cin, info.cout, err))
print("env HX=%s BASEOFF=%s: %d bad layers, worst err %.3e" % (os.environ.get("PF_HALO_HX"), os.environ.get("PF_HALO_BASEOFF"), bad, worst))
