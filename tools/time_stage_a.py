"""Dev tool (GPU box): device time of Stage A (pf_zsplat_forward_frames_hop[_packed]) at the bench size.
usage: python tools/time_stage_a.py [batch] [dist] [packed 0/1] [reps]   (env: PF_ZSPLAT_FAST, PF_ZSPLAT_MODE, PF_ZSPLAT_L2_MB)"""
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
    dist = sys.argv[2] if len(sys.argv) > 2 else "R"
    packed = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    dev = torch.device("cuda", 0)
    bg = build_model(bench.bg_params("tc")).eval()
    pipe = BGForecastPipeline(bg)
    sets = []
    for s in range(2):
        hs = bench.host_input_sets(1, batch, s, dist, synthetic, packed=True)[0][1 if packed else 0]
        sets.append({k: v.to(dev) for k, v in hs.items()})
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(3):
        pipe.warp(sets[i % 2], fuse_hop=True)
    torch.cuda.synchronize()
    ts = []
    for i in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pipe.warp(sets[i % 2], fuse_hop=True); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print("stage A batch %d dist %s packed %d FAST=%s MODE=%s L2_MB=%s: median %.3f ms, min %.3f ms per step" % (
        batch, dist, packed, os.environ.get("PF_ZSPLAT_FAST"), os.environ.get("PF_ZSPLAT_MODE"), os.environ.get("PF_ZSPLAT_L2_MB"),
        ts[len(ts) // 2], ts[0]))


if __name__ == "__main__":
    main()
