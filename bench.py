#!/usr/bin/env python
"""bg-forecast frames/s @1024x2048 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic input: `batch` target frames,
each forecast from 3 input frames: per-frame reprojection + z-buffer splat (Stage A) -> disk-hop
depth quantisation -> HarDNet-70 encoder/decoder + fused x4 upsample + argmax (Stage B).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = same path through
the public Python API from pinned HOST buffers (H2D of the inputs and D2H of the label map inside
the timed region).  The oracle is executed only in the `cpu_baseline` / `--impl reference` legs.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "bg-forecast frames/sec @1024x2048 (3 input frames -> 1 target)"
H, W, T = 1024, 2048, 3
STAGE_A_BYTES_PER_FRAME = 11 * T * H * W          # (4+1+1 read, 1+4 write) B x 3 frames x H*W
STAGE_B_FLOP_PER_FRAME = 75.32e9                  # SURVEY.md 8a conv census (2*MAC of the 70 convs)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                    "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    "source": "MEASURED_PEAKS.json (of measured)"}
        except Exception:
            pass
    # MEASURED_PEAKS.json is driver-written and git-ignored; BASELINE.md section 2 records its values.
    return {"hbm_gbs": 6555.8, "bf16_tflops": 1618.0, "bf16_tflops_sustained": 1354.4,
            "source": "BASELINE.md section 2 copy of MEASURED_PEAKS.json (of measured)"}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def make_state_dict(model, seed=0):
    from panoptic_forecasting_b200 import synthetic
    return synthetic.make_bg_state_dict({k: v.cpu() for k, v in model.state_dict().items()}, seed=seed)


def bg_params(precision):
    return {"task": "bg", "no_gpu": False, "load_best_model": False, "load_model": None,
            "data": {"num_classes": 11, "min_depth": 0.1, "max_depth": 200},
            "model": {"num_inputs": T, "use_depth_inps": True, "convert2onehot": True,
                      "final_w": W, "final_h": H,
                      "b200": {"precision": precision, "return_logits": False, "seg_dtype": "uint8"}}}


def host_input_sets(nsets, batch, seed0, dist):
    from panoptic_forecasting_b200 import synthetic
    sets = []
    for s in range(nsets):
        d = synthetic.make_pc_inputs(b=batch, t=T, h=H, w=W, dist=dist, seed=seed0 + s)
        d["intrinsics_inv"] = torch.inverse(d["intrinsics"]).contiguous()
        d["extrinsics_inv"] = torch.inverse(d["extrinsics"]).contiguous()
        sets.append(d)
    return sets


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (oracle port, all host threads)."""
    if rank != 0:
        return
    from oracle import cpu_port
    from panoptic_forecasting_b200.models import build_model
    torch.set_num_threads(os.cpu_count())
    p = bg_params("fp32")
    p["no_gpu"] = True
    sd = make_state_dict(build_model(p), 0)
    inp = host_input_sets(1, 1, 1000, args.dist)[0]
    inp = {k: v for k, v in inp.items() if not k.endswith("_inv")}
    t_first = None
    for _ in range(max(1, min(args.warmup, 1))):
        t0 = time.perf_counter()
        cpu_port.composite_predict(sd, inp, (H, W))
        t_first = time.perf_counter() - t0
    steps = max(1, min(args.steps, int(150.0 / max(t_first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_port.composite_predict(sd, inp, (H, W))
    dt = time.perf_counter() - t0
    fps = steps / dt
    sample = "%d x (1 target frame from 3 input frames @%dx%d, dist %s) of the %d requested steps" % (
        steps, H, W, args.dist, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, batch):
    return {"workload": "config 3: full reproject + z-buffer splat + disk-hop + HarDNet-70 decode, "
                        "3 input frames -> 1 target @1024x2048",
            "batch_per_step": batch, "input_frames": T, "height": H, "width": W, "depth_distribution": args.dist,
            "precision": args.precision, "outputs": "uint8 label map (full-res logits not materialised)",
            "l2": "inputs rotate over %d distinct sets and the activation arena (> L2) is rewritten every step" % args.nsets,
            "weights": "seeded random HarDNet-70 (BN statistics randomised)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="target frames per step (the reference export uses batch_size 2)")
    ap.add_argument("--precision", default=os.environ.get("PF_PRECISION", "tc"), choices=["fp32", "tc"])
    ap.add_argument("--dist", default="R", choices=["R", "U"])
    ap.add_argument("--nsets", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from panoptic_forecasting_b200 import _lib
    from panoptic_forecasting_b200.models import build_model
    from panoptic_forecasting_b200.pipeline import BGForecastPipeline

    L = _lib.lib()
    bg = build_model(bg_params(args.precision)).eval()
    bg.load_state_dict(make_state_dict(bg, 0))
    pipe = BGForecastPipeline(bg)
    B = args.batch

    host_sets = host_input_sets(args.nsets, B, 100 * rank, args.dist)
    pinned = [{k: v.pin_memory() for k, v in s.items()} for s in host_sets]
    dev_sets = [{k: v.to(dev) for k, v in s.items()} for s in host_sets]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host_sets[0].values())
    d2h_bytes = B * H * W

    K = args.steps
    out_maps = torch.empty((K * B, H, W), dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(out_maps) for _ in range(world)] if (world > 1 and rank == 0) else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if (rank == 0 and not os.environ.get('PF_NO_SAMPLER')) else None   # runs through warm-up + both timed regions
    # ---- warm-up
    # Same statement sequence as the timed loop, with the previous step's outputs still referenced while the next
    # step allocates its own: otherwise the caching allocator holds ONE set of output buffers after warm-up and the
    # second timed step pays a 600 MB cudaMalloc with the GPU idle (seen as a 17-180 ms outlier in 1 run out of 4).
    seg = d = m = out = None
    for i in range(args.warmup):
        seg, d, m = pipe.warp(dev_sets[i % args.nsets], fuse_hop=True)
        out = bg.predict({"seg": seg, "depth": d, "depth_mask": m}, {})
        out_maps[(i % K) * B:(i % K + 1) * B].copy_(out["seg"])
    if world > 1:
        dist.gather(out_maps, gathered, dst=0)          # warm-up: NCCL connection set-up is not part of the job
    barrier()

    # ---- timed region 1: device-resident inputs (`value`)
    # CUDA events bracket Stage A (splat) and Stage B (net) of every step on the launching stream; no
    # events are recorded BETWEEN the net's kernels here because they would serialise the programmatic
    # dependent launches of consecutive conv layers (the per-kernel split is taken in a separate pass below).
    nsteps_net = L.pf_bgnet_num_steps(bg._net)
    ev_a = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    ev_b = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        inp = dev_sets[i % args.nsets]
        ev_a[i][0].record()
        seg, d, m = pipe.warp(inp, fuse_hop=True)          # disk hop fused into the resolve kernel
        ev_a[i][1].record()
        ev_b[i][0].record()
        out = bg.predict({"seg": seg, "depth": d, "depth_mask": m}, {})
        ev_b[i][1].record()
        out_maps[i * B:(i + 1) * B].copy_(out["seg"])
    if world > 1:
        # the path's only collective: ONE gather of the per-rank label maps
        dist.gather(out_maps, gathered, dst=0)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    fps = world * B * K / (ms_total_max / 1e3)
    warp_ms = statistics.mean(a.elapsed_time(b_) for a, b_ in ev_a)
    net_ms = statistics.mean(a.elapsed_time(b_) for a, b_ in ev_b)
    step_a = [a.elapsed_time(b_) for a, b_ in ev_a]
    step_b = [a.elapsed_time(b_) for a, b_ in ev_b]
    gaps = [ev_b[i][1].elapsed_time(ev_a[i + 1][0]) for i in range(K - 1)]
    if os.environ.get('PF_BENCH_DEBUG') or max(step_a) > 2 * statistics.median(step_a) or max(step_b) > 2 * statistics.median(step_b):
        print('stage A ms per step:', ['%.2f' % x for x in step_a], file=sys.stderr)
        print('stage B ms per step:', ['%.2f' % x for x in step_b], file=sys.stderr)
        print('gap ms between steps:', ['%.2f' % x for x in gaps], file=sys.stderr)

    # per-kernel split of Stage B (separate pass, events between all kernels; not part of `value`)
    Kp = min(K, 5)
    _lib.check(L.pf_bgnet_set_profiling(bg._net, Kp), "pf_bgnet_set_profiling")
    for i in range(Kp):
        pipe.forecast(dev_sets[i % args.nsets])
    ms_steps = (C.c_float * nsteps_net)()
    n_prof = L.pf_bgnet_read_profile(bg._net, ms_steps, nsteps_net)
    _lib.check(L.pf_bgnet_set_profiling(bg._net, 0), "pf_bgnet_set_profiling")
    conv_ms = first_ms = other_ms = 0.0
    ty, ci = C.c_int(), C.c_int()
    for k in range(nsteps_net):
        L.pf_bgnet_step_info(bg._net, k, C.byref(ty), C.byref(ci))
        if ty.value == 1:
            conv_ms += ms_steps[k]
        elif ty.value == 0:
            first_ms += ms_steps[k]
        else:
            other_ms += ms_steps[k]
    peaks = load_peaks()
    # dominant kernel family: the ConvLayer kernels of pf_bgnet_forward (68 tcgen05 launches per step)
    net_tflops = STAGE_B_FLOP_PER_FRAME * B / (max(net_ms, 1e-9) * 1e-3) / 1e12
    conv_share = (conv_ms + first_ms) / max(conv_ms + first_ms + other_ms, 1e-9)
    roof = {"bound": "tensor", "kernel": "pf_bgnet_forward: 70 ConvLayers (%s path; conv kernels = %.0f%% of its time)" % (
                args.precision, 100 * conv_share),
            "achieved": net_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": net_tflops / peaks["bf16_tflops_sustained"], "traffic": None,
            "ms_per_step": net_ms, "timed_steps": K, "peak_source": peaks["source"],
            "note": "algorithmic 75.32 GFLOP/frame; the split-bf16 scheme executes 2 tensor-core MMAs per algorithmic MAC"}
    # DRAM traffic per step from the committed ncu capture of this same command (profiles/, batch 16 only)
    traffic_b = traffic_a = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_v14_traffic_per_step.json")))
        if tr.get("batch") == B:
            ks = tr["per_step"]
            tot = lambda pred: sum(v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in ks.items() if pred(k))
            traffic_b = tot(lambda k: not k.startswith("zsplat") and "zsplat" not in k)
            traffic_a = tot(lambda k: "zsplat" in k)
    except Exception:
        pass
    roof["traffic"] = traffic_b
    roof["traffic_note"] = "bytes per step (all Stage B kernels), ncu dram__bytes_read+write, profiles/r1_v14_traffic_per_step.json"
    a_gbs = STAGE_A_BYTES_PER_FRAME * B / (warp_ms * 1e-3) / 1e9
    roof_a = {"bound": "hbm", "kernel": "pf_zsplat_forward_frames (points + resolve)", "achieved": a_gbs,
              "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": a_gbs / peaks["hbm_gbs"], "traffic": traffic_a,
              "ms_per_step": warp_ms}

    # ---- timed region 2: end to end through the public API from pinned host buffers (`e2e`):
    # every step uploads its own inputs (pinned host -> device) and downloads its own label map; the
    # PipelinedForecaster overlaps step i+1's upload with step i's kernels (2 slots in flight).
    from panoptic_forecasting_b200.pipeline import PipelinedForecaster
    Ke = max(6, min(K, 20))
    pf = PipelinedForecaster(pipe, depth=2)
    for i in range(3):
        pf.submit(pinned[i % args.nsets])
        pf.collect()
    barrier()
    checksum = 0
    e0.record()                                             # GPU idle here: timestamp = region start
    for i in range(Ke):
        pf.submit(pinned[i % args.nsets])
        if i >= 1:
            checksum += int(pf.collect()[0, 0, 0])            # the caller consumes every label map
    checksum += int(pf.collect()[0, 0, 0])
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    barrier()
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_fps = world * B * Ke / (float(t.item()) / 1e3)
    clocks = sampler.stop() if sampler else None

    launches_per_step = L.pf_zsplat_launches_for(B, T, H, W) + L.pf_bgnet_launches_per_forward(bg._net)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_port
        torch.set_num_threads(os.cpu_count())
        sd = {k: v.cpu() for k, v in bg.state_dict().items()}
        one = {k: v[:1].contiguous() for k, v in host_sets[0].items() if not k.endswith("_inv")}
        t0 = time.perf_counter()
        ref = cpu_port.composite_predict(sd, one, (H, W))
        t1 = time.perf_counter() - t0
        n = max(1, min(5, int(15.0 / t1)))
        t0 = time.perf_counter()
        for _ in range(n):
            cpu_port.composite_predict(sd, one, (H, W))
        dt = (time.perf_counter() - t0) / n
        # the same frame through the CUDA path: label-map agreement with the CPU port (reported, not timed)
        mine = pipe.forecast({k: v[:1].to(dev) for k, v in host_sets[0].items()})["seg"].cpu()
        agree = float((mine.long() == ref["seg"]).float().mean())
        cpu_base = {"value": 1.0 / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": "%d x 1 target frame @%dx%d after 1 warm-up (oracle/cpu_port.py, torch %s CPU ops)" % (
                        n, H, W, torch.__version__),
                    "label_agreement_with_cuda_path": agree}

    if rank == 0:
        line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K,
                "warmup": args.warmup, "ms_per_step": ms_total_max / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if args.precision == "fp32" else "bf16x3 (split bf16, fp32 accumulate)",
                "data": "synthetic", "config": workload_config(args, B),
                "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes, "steps": Ke},
                "gpu_launches": launches_per_step * K, "gpu_launches_per_step": launches_per_step,
                "roofline": roof, "roofline_stage_a": roof_a,
                "stage_ms_per_step": {"stage_a_warp": warp_ms, "stage_b_net": net_ms,
                                      "profiled_pass": {"convs": conv_ms, "first_conv": first_ms,
                                                        "pool_upsample_head": other_ms, "steps": n_prof}},
                "clocks": clocks, "cpu_baseline": cpu_base}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
