"""Dev tool (GPU box): per-step device time of pf_bgnet_forward (CUDA events around every step).
usage: python tools/layer_profile.py [precision] [batch] [H] [W]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from panoptic_forecasting_b200 import _lib, synthetic
from panoptic_forecasting_b200.models import build_model


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "tc"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    W = int(sys.argv[4]) if len(sys.argv) > 4 else 2048
    L = _lib.lib()
    p = bench.bg_params(precision)
    p["model"]["final_h"], p["model"]["final_w"] = H, W
    bg = build_model(p).eval()
    bg.load_state_dict(bench.make_state_dict(bg, 0))
    inp = synthetic.make_bg_inputs(batch, 3, H, W, seed=0, device="cuda", label_dtype=torch.uint8)
    for _ in range(3):
        bg.predict(inp, {})
    n = L.pf_bgnet_num_steps(bg._net)
    iters = 10
    _lib.check(L.pf_bgnet_set_profiling(bg._net, iters), "prof")
    for _ in range(iters):
        bg.predict(inp, {})
    ms = (C.c_float * n)()
    L.pf_bgnet_read_profile(bg._net, ms, n)
    info = _lib.ConvInfo()
    ty, ci = C.c_int(), C.c_int()
    names = {0: "first", 1: "conv", 2: "pool", 3: "upsample", 4: "head"}
    tot = sum(ms)
    print("precision %s batch %d %dx%d: total %.3f ms" % (precision, batch, H, W, tot))
    for k in range(n):
        L.pf_bgnet_step_info(bg._net, k, C.byref(ty), C.byref(ci))
        desc = ""
        if ci.value >= 0:
            L.pf_bgnet_conv_info(bg._net, ci.value, C.byref(info))
            desc = "%-34s %4d->%-4d k%d s%d" % (info.name.decode(), info.cin, info.cout, info.ksize, info.stride)
        print("%3d %-9s %-52s %8.1f us  %5.1f%%" % (k, names[ty.value], desc, ms[k] * 1e3, 100 * ms[k] / tot))


if __name__ == "__main__":
    main()
