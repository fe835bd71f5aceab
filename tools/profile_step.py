"""Dev tool (GPU box): ONE bench step (Stage A + Stage B, batch 16 by default) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics ... --csv --log-file gpurun_out/step.csv python tools/profile_step.py
and the launch-order -> layer-name table profiles/per_layer_table.py joins it with (gpurun_out/step_layers.json)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from panoptic_forecasting_b200 import _lib, synthetic
from panoptic_forecasting_b200.models import build_model
from panoptic_forecasting_b200.pipeline import BGForecastPipeline


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    dev = torch.device("cuda", 0)
    L = _lib.lib()
    bg = build_model(bench.bg_params("tc")).eval()
    bg.load_state_dict(bench.make_state_dict(bg, 0, synthetic))
    pipe = BGForecastPipeline(bg)
    sets = [{k: v.to(dev) for k, v in s[1].items()} for s in bench.host_input_sets(2, batch, 0, "R", synthetic, packed=True)]
    for i in range(3):
        pipe.forecast(sets[i % 2])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for i in range(steps):
        pipe.forecast(sets[(i + 1) % 2])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    info = _lib.ConvInfo()
    ty, ci = C.c_int(), C.c_int()
    rows = []
    for k in range(L.pf_bgnet_num_steps(bg._net)):
        L.pf_bgnet_step_info(bg._net, k, C.byref(ty), C.byref(ci))
        row = {"step": k, "type": {0: "first", 1: "conv", 2: "pool", 3: "upsample", 4: "head"}[ty.value]}
        if ci.value >= 0:
            L.pf_bgnet_conv_info(bg._net, ci.value, C.byref(info))
            row.update(name=info.name.decode(), cin=info.cin, cout=info.cout, k=info.ksize, stride=info.stride)
        rows.append(row)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "step_layers.json")
    json.dump({"batch": batch, "steps": steps, "layers": rows}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
