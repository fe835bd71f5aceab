"""Dev tool (GPU box): Stage A of step i+1 on a second stream under Stage B of step i, vs the sequential loop.
usage: python tools/overlap_ab.py [batch] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from panoptic_forecasting_b200 import synthetic
from panoptic_forecasting_b200.models import build_model
from panoptic_forecasting_b200.pipeline import BGForecastPipeline


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    dev = torch.device("cuda", 0)
    bg = build_model(bench.bg_params("tc")).eval()
    bg.load_state_dict(bench.make_state_dict(bg, 0, synthetic))
    pipe = BGForecastPipeline(bg)
    sets = [{k: v.to(dev) for k, v in s[1].items()} for s in bench.host_input_sets(3, batch, 0, "R", synthetic, packed=True)]

    def sequential():
        for i in range(steps):
            seg, d, m = pipe.warp(sets[i % 3], fuse_hop=True)
            bg.predict({"seg": seg, "depth": d, "depth_mask": m}, {})

    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

    def overlapped():
        keep = []
        evs = []
        for i in range(steps):
            with torch.cuda.stream(sa):
                if i >= 2:
                    sa.wait_event(evs[i - 2][1])            # at most two warps ahead of the net
                w = pipe.warp(sets[i % 3], fuse_hop=True)
                ea = torch.cuda.Event(); ea.record(sa)
            with torch.cuda.stream(sb):
                sb.wait_event(ea)
                out = bg.predict({"seg": w[0], "depth": w[1], "depth_mask": w[2]}, {})
                eb = torch.cuda.Event(); eb.record(sb)
            evs.append((ea, eb))
            keep.append((w, out))
        torch.cuda.current_stream().wait_stream(sa)
        torch.cuda.current_stream().wait_stream(sb)

    for name, fn in (("sequential", sequential), ("overlapped", overlapped), ("sequential", sequential), ("overlapped", overlapped)):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print("%-10s %.3f ms/step  %.1f frames/s" % (name, ms, batch / ms * 1e3))


if __name__ == "__main__":
    main()
