"""Dev tool (GPU box): device time of pf_panoptic_merge vs its HBM roofline, and the oracle's CPU time beside it.
usage: python tools/time_merge.py [batch] [instances_per_item]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import panoptic_merge_oracle as merge_oracle
from panoptic_forecasting_b200 import panoptic, synthetic

H, W = 1024, 2048


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    case = synthetic.make_merge_inputs(b, (n,) * b, H, W, seed=0)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    probs = [torch.sigmoid(t(l)) for l in case["mask_logits"]]
    args = dict(pred_bboxes=[t(x) for x in case["bboxes"]], orig_classes=[t(c) for c in case["classes"]],
                pred_depths=[t(d) for d in case["depths"]], background=torch.stack([t(x) for x in case["background"]]),
                background_depths=torch.stack([t(x) for x in case["bg_depth"]]),
                background_depth_masks=torch.stack([t(x) for x in case["bg_depth_mask"]]))
    for _ in range(3):
        out = panoptic.merge_instances(probs, **args)["seg"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = panoptic.merge_instances(probs, **args)["seg"]
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = float(np.median(ms))
    prepared = panoptic.prepare_instances(probs, args["pred_bboxes"], args["orig_classes"], args["pred_depths"])
    bgm = args["background_depth_masks"].to(torch.uint8)
    kms = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out2 = panoptic.merge_prepared(prepared, b, args["background"], args["background_depths"], bgm)
        e1.record()
        torch.cuda.synchronize()
        kms.append(e0.elapsed_time(e1))
    kms = float(np.median(kms))
    assert torch.equal(out, out2)
    bg8 = args["background"].to(torch.uint8)
    k8 = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out3 = panoptic.merge_prepared(prepared, b, bg8, args["background_depths"], bgm)
        e1.record()
        torch.cuda.synchronize()
        k8.append(e0.elapsed_time(e1))
    k8 = float(np.median(k8))
    assert torch.equal(out, out3)
    print("pf_panoptic_merge kernel, uint8 background: %.3f ms -> %.0f GB/s of 14 B/px" % (k8, b * H * W * 14 / k8 / 1e6))
    algo = b * H * W * (8 + 4 + 1 + 8)
    print("pf_panoptic_merge kernel alone: %.3f ms -> %.0f frames/s, %.0f GB/s (HBM peak 6556)" % (kms, b / kms * 1e3, algo / kms / 1e6))
    print("pf_panoptic_merge (wrapper incl. order/ids): batch %d x %d instances: %.3f ms  -> %.0f frames/s, %.0f GB/s of %d MB algorithmic"
          % (b, n, ms, b / ms * 1e3, algo / ms / 1e6, algo >> 20))
    t0 = time.time()
    ref = merge_oracle.merge(case["background"][0], probs[0].cpu().numpy(), case["bboxes"][0], case["classes"][0], case["depths"][0],
                             bg_depth=case["bg_depth"][0], bg_depth_mask=case["bg_depth_mask"][0])
    cpu = time.time() - t0
    print("oracle (numpy, 1 core): %.2f s per frame; identical: %s" % (cpu, bool(np.array_equal(ref, out[0].cpu().numpy()))))


if __name__ == "__main__":
    main()
