"""Dev tool (GPU box): device time of each stage of the composite path on several synthetic sets.
usage: python tools/time_stages.py [precision] [batch]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from panoptic_forecasting_b200.models import build_model
from panoptic_forecasting_b200.pipeline import BGForecastPipeline


def timed(name, fn, n=5):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    cpu = []
    ev0.record()
    for _ in range(n):
        t0 = time.perf_counter()
        r = fn()
        cpu.append((time.perf_counter() - t0) * 1e3)
    ev1.record()
    torch.cuda.synchronize()
    print("  %-14s gpu %.3f ms   cpu enqueue %.3f ms" % (name, ev0.elapsed_time(ev1) / n, min(cpu)))
    return r


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "tc"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda", 0)
    bg = build_model(bench.bg_params(precision)).eval()
    bg.load_state_dict(bench.make_state_dict(bg, 0))
    pipe = BGForecastPipeline(bg)
    for dist in ("R", "U"):
        for seed in (0, 1, 2):
            hs = bench.host_input_sets(1, batch, seed, dist)[0][0]
            inp = {k: v.to(dev) for k, v in hs.items()}
            print("dist %s seed %d batch %d precision %s" % (dist, seed, batch, precision))
            seg, depth = timed("warp", lambda: pipe.warp(inp))
            d, m = timed("decode_depth", lambda: pipe.decode_depth(depth))
            timed("bg.predict", lambda: bg.predict({"seg": seg, "depth": d, "depth_mask": m}, {}))
            timed("forecast", lambda: pipe.forecast(inp))
            # warp timed inside the full pipeline (L2 holds the previous step's activations)
            evs = []
            for _ in range(5):
                a, b_, c_ = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                a.record(); s2, d2 = pipe.warp(inp); b_.record()
                dd, mm = pipe.decode_depth(d2); c_.record()
                bg.predict({"seg": s2, "depth": dd, "depth_mask": mm}, {})
                evs.append((a, b_, c_))
            torch.cuda.synchronize()
            print("  interleaved: warp %.3f ms, disk hop %.3f ms" % (
                sum(a.elapsed_time(b_) for a, b_, _ in evs) / 5, sum(b_.elapsed_time(c_) for _, b_, c_ in evs) / 5))


if __name__ == "__main__":
    main()
