#!/usr/bin/env python
"""bg-forecast frames/s @1024x2048 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic input: `batch` target frames,
each forecast from 3 input frames: per-frame reprojection + z-buffer splat (Stage A) -> disk-hop
depth quantisation -> HarDNet-70 encoder/decoder + fused x4 upsample + argmax (Stage B).

  python bench.py --gpus N --steps K --warmup W              # our CUDA path (config 3)
  python bench.py --config 2 ...                             # bg net only, 512x1024 (BASELINE.json configs[1])
  python bench.py --impl reference --steps K --warmup W      # the reference's own CPU path (baseline/_ref, else the port)
  python bench.py --impl torch_cuda --steps K --warmup W     # the reference's torch ops on the B200 (cuDNN fp32): library bar

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = same path through
the public Python API from pinned HOST buffers (H2D of the inputs and D2H of the label map inside
the timed region).  The oracle is executed only in the `cpu_baseline` / `library_baseline` /
`parity` legs and in `--impl reference|torch_cuda`; never in a timed region of our arm.
"""
import argparse
import ctypes as C
import importlib.util
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
METRIC2 = "bg HarDNet-70 forward frames/sec @512x1024 (3 input frames)"
H, W, T = 1024, 2048, 3
STAGE_A_BYTES_PER_FRAME = 11 * T * H * W          # (4+1+1 read, 1+4 write) B x 3 frames x H*W
STAGE_B_FLOP_PER_FRAME = 75.32e9                  # SURVEY.md 8a conv census (2*MAC of the 70 convs)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2_traffic_per_step.json")


def load_synthetic():
    """panoptic_forecasting_b200/synthetic.py loaded by PATH: numpy/torch only, and the reference arm must not
    import the product package (nor map libpf_b200.so)."""
    spec = importlib.util.spec_from_file_location("pf_synthetic", os.path.join(ROOT, "panoptic_forecasting_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


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
    return {"hbm_gbs": 6551.0, "bf16_tflops": 1686.7, "bf16_tflops_sustained": 1402.0,
            "source": "copy of MEASURED_PEAKS.json kept in bench.py (of measured)"}


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


def bg_params(precision, h=H, w=W, return_logits=False):
    b200 = {"precision": precision, "return_logits": return_logits}
    if not return_logits:
        b200["seg_dtype"] = "uint8"
    return {"task": "bg", "no_gpu": False, "load_best_model": False, "load_model": None,
            "data": {"num_classes": 11, "min_depth": 0.1, "max_depth": 200},
            "model": {"num_inputs": T, "use_depth_inps": True, "convert2onehot": True,
                      "final_w": w, "final_h": h, "b200": b200}}


def make_state_dict(model_or_shapes, seed=0, synthetic=None):
    synthetic = synthetic or load_synthetic()
    if hasattr(model_or_shapes, "state_dict"):
        like = {k: v.cpu() for k, v in model_or_shapes.state_dict().items()}
    else:
        like = {k: torch.zeros(s) for k, s in model_or_shapes.items()}
    return synthetic.make_bg_state_dict(like, seed=seed)


def host_input_sets(nsets, batch, seed0, dist, synthetic=None, packed=False):
    """Seeded synthetic PCTransformModel inputs.  Depth is a value of the Cityscapes disparity table (as real
    Cityscapes depth is: uint16 disparity PNG -> depth), so the packed (uint16 code + table + 1-bit mask) and
    the reference-format (float32 depth + bool mask) dicts describe the SAME frames.
    Returns [(reference-format dict, packed dict or None)]."""
    synthetic = synthetic or load_synthetic()
    sets = []
    for s in range(nsets):
        d = synthetic.make_pc_inputs(b=batch, t=T, h=H, w=W, dist=dist, seed=seed0 + s)
        pk, d = synthetic.pack_pc_inputs(d)
        inv = {"intrinsics_inv": torch.inverse(d["intrinsics"]).contiguous(),
               "extrinsics_inv": torch.inverse(d["extrinsics"]).contiguous()}
        d = dict(d, **inv)
        pk = dict(pk, **inv) if packed else None
        sets.append((d, pk))
    return sets


def workload_config(args, batch):
    if args.config == 2:
        return {"workload": "config 2: bg HarDNet-70 encoder+decoder forward (one-hot/depth-norm first conv ... fused x4 "
                            "upsample + argmax), 3 input frames @512x1024",
                "batch_per_step": batch, "input_frames": T, "height": 512, "width": 1024,
                "precision": args.precision, "outputs": "uint8 label map @512x1024 (final_size = input size)",
                "l2": "inputs rotate over %d distinct sets and the activation arena (> L2) is rewritten every step" % args.nsets,
                "weights": "seeded random HarDNet-70 (BN statistics randomised)"}
    return {"workload": "config 3: full reproject + z-buffer splat + disk-hop + HarDNet-70 decode, "
                        "3 input frames -> 1 target @1024x2048",
            "batch_per_step": batch, "input_frames": T, "height": H, "width": W, "depth_distribution": args.dist,
            "precision": args.precision, "outputs": "uint8 label map (full-res logits not materialised)",
            "input_format": ("packed: uint16 disparity code + 65536-entry float32 depth table + 1-bit mask + uint8 labels "
                             "(9.4 B/px/target frame)" if args.input_format == "packed" else
                             "reference formats: float32 depth + bool mask + uint8 labels (18 B/px/target frame)"),
            "l2": "inputs rotate over %d distinct sets and the activation arena (> L2) is rewritten every step" % args.nsets,
            "weights": "seeded random HarDNet-70 (BN statistics randomised)"}


# ------------------------------------------------------------------------------------------------------------------
# reference arms: no product import, no libpf_b200.so
# ------------------------------------------------------------------------------------------------------------------
def reference_step_fn(device, frames, dist, seed):
    """Returns (step(), kind, description): one call = the reference's composite path over `frames` target frames on
    `device`: the UNMODIFIED reference from baseline/_ref (or /root/reference) when importable, else the oracle port."""
    from oracle import bg_oracle, cpu_port, ref_loader
    synthetic = load_synthetic()
    inp = host_input_sets(1, frames, seed, dist, synthetic)[0][0]
    inp = {k: v.to(device) for k, v in inp.items() if not k.endswith("_inv")}
    sd = {k: v.to(device) for k, v in make_state_dict(bg_oracle.state_dict_shapes(), 0, synthetic).items()}
    try:
        ref = ref_loader.load_reference()
        p = ref_loader.ref_bg_params(H, W)
        bg = ref.build_model(p).eval()
        bg.load_state_dict({k: v.cpu() for k, v in sd.items()})
        bg.to(device)
        pcs = [ref.build_model(ref_loader.ref_pc_params(i)).to(device) for i in range(T)]

        def step():
            with torch.no_grad():
                segs, deps, masks = [], [], []
                for i in range(T):
                    r = pcs[i].predict({k: v for k, v in inp.items()}, {})
                    d, m = cpu_port.disk_hop(r["depth"])          # disk hop between the two exports
                    segs.append(r["seg"]); deps.append(d); masks.append(m)
                out = bg.predict({"seg": torch.stack(segs, 1).long(), "depth": torch.stack(deps, 1),
                                  "depth_mask": torch.stack(masks, 1)}, {})
                return out["seg"]
        return step, "reference", "unmodified reference package (%s) + torch_scatter.scatter_min shim" % ref_loader.REFERENCE_ROOT
    except Exception as e:                                       # reference tree absent: the port (same torch ops)
        why = "%s: %s" % (type(e).__name__, e)

        def step():
            return cpu_port.composite_predict(sd, inp, (H, W))["seg"]
        return step, "port", "oracle/cpu_port.py (reference not importable: %s)" % why[:80]


def run_reference(args, rank):
    """The reference's own CPU implementation of the path, all host threads."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    frames = max(1, min(args.batch, args.ref_frames))            # bounded sample of the step's `batch` frames
    step, kind, desc = reference_step_fn(torch.device("cpu"), frames, args.dist, 1000)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = args.steps * frames / dt
    sample = "each of the %d steps (after %d warm-up) = %d of the step's %d target frames @%dx%d, dist %s; %s; torch %s" % (
        args.steps, args.warmup, frames, args.batch, H, W, args.dist, desc, torch.__version__)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps * (args.batch / frames),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.batch),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": kind,
                             "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def time_library_baseline(frames, dist, warmup, steps):
    """The reference's torch ops on the B200: cuDNN fp32 (TF32 off, cudnn.benchmark=True as the reference's export
    script sets, export_cityscapes_segmentation_results.py:171), shimmed scatter on the device."""
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", torch.cuda.current_device())
    step, kind, desc = reference_step_fn(dev, frames, dist, 1000)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": frames / (ms * 1e-3), "unit": "frames/s", "kind": kind, "ms_per_step": ms, "batch": frames,
            "sample": "%d steps of %d target frames after %d warm-up; %s; torch %s CUDA ops, cuDNN %s fp32, TF32 off, "
                      "cudnn.benchmark on, inputs resident on the device" % (
                          steps, frames, warmup, desc, torch.__version__, torch.backends.cudnn.version())}


def run_torch_cuda(args, rank):
    if rank != 0:
        return
    assert torch.cuda.is_available(), "--impl torch_cuda needs a CUDA device"
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    lb = time_library_baseline(max(1, min(args.batch, args.ref_frames)), args.dist, args.warmup, args.steps)
    line = {"impl": "torch_cuda", "metric": METRIC, "value": lb["value"], "unit": "frames/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": lb["ms_per_step"] * args.batch / lb["batch"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.batch), "library_baseline": lb}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def parity_block(pipe_fast, bg_fast, host_ref_inputs, dev, precision):
    """One frame of the benched workload through the CPU port (reference arithmetic) and through the CUDA path:
    label mismatches, whether every one is a near-tie of the reference's own logits, logits error.  Same definitions
    as tests/test_parity_full_gpu.py.  Returns (parity dict, seconds the CPU port took for the frame)."""
    from oracle import cpu_port
    from panoptic_forecasting_b200.models import build_model
    from panoptic_forecasting_b200.pipeline import BGForecastPipeline
    sd = {k: v.cpu() for k, v in bg_fast.state_dict().items()}
    one = {k: v[:1].contiguous() for k, v in host_ref_inputs.items()}
    t0 = time.perf_counter()
    ref = cpu_port.composite_predict(sd, {k: v for k, v in one.items() if not k.endswith("_inv")}, (H, W))
    t_ref = time.perf_counter() - t0
    cu = {k: v.to(dev) for k, v in one.items()}
    mine = pipe_fast.forecast(cu)
    full_model = build_model(bg_params(precision, return_logits=True)).eval()
    full_model.load_state_dict(sd)
    full = BGForecastPipeline(full_model).forecast(cu)
    scale = ref["logits"].abs().max().item()
    eps = (full["logits"].cpu() - ref["logits"]).abs().max().item()
    ours = mine["seg"].cpu().long()
    mism = ours != ref["seg"]
    n = int(mism.sum())
    gap = 0.0
    if n:
        top = ref["logits"].max(1).values
        got = ref["logits"].gather(1, ours.unsqueeze(1)).squeeze(1)
        gap = float((top - got)[mism].max())
    stage_a = bool(torch.equal(mine["warped_seg"].cpu().long(), ref["warped_seg"]) and
                   torch.equal(mine["warped_depth"].cpu(), ref["warped_depth"]))
    del full_model, full
    return {"frame": "1 target frame of the benched workload (set 0, item 0) vs oracle/cpu_port.py",
            "pixels": int(mism.numel()), "label_mismatch_px": n,
            "all_near_ties": bool(gap <= 2.0 * eps + 1e-6 * scale) and bool(eps <= 1e-3 * scale),
            "within_tolerance": bool(eps <= 1e-3 * scale) and stage_a,
            "worst_ref_gap_at_mismatch_rel": gap / scale, "logits_rel_err": eps / scale,
            "stage_a_bit_exact": stage_a,
            "near_tie_definition": "reference logit of our class within 2 x (max abs logit error) of the reference maximum"}, t_ref, ref


def run_config2(args, rank, world, local_rank, dist):
    """BASELINE.json configs[1]: the bg net alone on 512x1024 synthetic frames."""
    from panoptic_forecasting_b200 import _lib, synthetic
    from panoptic_forecasting_b200.models import build_model
    dev = torch.device("cuda", local_rank)
    L = _lib.lib()
    h2, w2 = 512, 1024
    bg = build_model(bg_params(args.precision, h2, w2)).eval()
    bg.load_state_dict(make_state_dict(bg, 0, synthetic))
    B, K = args.batch, args.steps
    host = [synthetic.make_bg_inputs(B, T, h2, w2, seed=100 * rank + s, label_dtype=torch.uint8) for s in range(args.nsets)]
    devs = [{k: v.to(dev) for k, v in s.items()} for s in host]
    pinned = [{k: v.pin_memory() for k, v in s.items()} for s in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    out = None
    for i in range(args.warmup):
        out = bg.predict(devs[i % args.nsets], {})
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        out = bg.predict(devs[i % args.nsets], {})
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    fps = world * B * K / (ms * 1e-3)
    # e2e: pinned host inputs in, label map out, copies inside the timed region (2 streams, 2 slots)
    copy_s = torch.cuda.Stream()
    slots = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items()} for _ in range(2)]
    outs = [torch.empty((B, h2, w2), dtype=torch.uint8).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    up = [torch.cuda.Event() for _ in range(2)]
    barrier()
    e0.record()
    for i in range(K):
        s = i % 2
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(done[s])
            for k, v in pinned[i % args.nsets].items():
                slots[s][k].copy_(v, non_blocking=True)
            up[s].record(copy_s)
        torch.cuda.current_stream().wait_event(up[s])
        o = bg.predict(slots[s], {})["seg"]
        outs[s].copy_(o, non_blocking=True)
        done[s].record()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_fps = world * B * K / (float(t.item()) * 1e-3)
    lat = []
    one = {k: v[:1].contiguous() for k, v in devs[0].items()}
    for _ in range(3):
        bg.predict(one, {})
    for _ in range(20):
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); bg.predict(one, {}); b_.record(); torch.cuda.synchronize()
        lat.append(a.elapsed_time(b_))
    clocks = sampler.stop() if sampler else None
    peaks = load_peaks()
    tfl = 18.83e9 * world * B * K / (ms * 1e-3) / 1e12 / world
    if rank == 0:
        line = {"metric": METRIC2, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if args.precision == "fp32" else "bf16x3 (split bf16, fp32 accumulate)",
                "data": "synthetic", "config": workload_config(args, B),
                "e2e": {"value": e2e_fps, "unit": "frames/s",
                        "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host[0].values()),
                        "d2h_bytes_per_step": B * h2 * w2},
                "gpu_launches": L.pf_bgnet_launches_per_forward(bg._net) * K,
                "latency_ms_batch1": statistics.median(lat),
                "roofline": {"bound": "tensor", "kernel": "pf_bgnet_forward (70 ConvLayers, 18.83 GFLOP/frame)",
                             "achieved": tfl, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                             "frac": tfl / peaks["bf16_tflops_sustained"], "traffic": None,
                             "peak_source": peaks["source"]},
                "clocks": clocks, "cpu_baseline": None}
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_cuda"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3], help="BASELINE.json configs index (3 = the metric's)")
    ap.add_argument("--batch", type=int, default=16, help="target frames per step (the reference export uses batch_size 2)")
    ap.add_argument("--ref-frames", type=int, default=2,
                    help="reference arms: target frames actually processed per step (bounded sample of --batch)")
    ap.add_argument("--precision", default=os.environ.get("PF_PRECISION", "tc"), choices=["fp32", "tc"])
    ap.add_argument("--dist", default="R", choices=["R", "U"])
    ap.add_argument("--input-format", default="packed", choices=["packed", "f32"])
    ap.add_argument("--nsets", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.impl == "torch_cuda":
        run_torch_cuda(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.config == 2:
        run_config2(args, rank, world, local_rank, dist)
        if world > 1:
            dist.destroy_process_group()
        return

    from panoptic_forecasting_b200 import _lib, synthetic
    from panoptic_forecasting_b200.models import build_model
    from panoptic_forecasting_b200.pipeline import BGForecastPipeline, PipelinedForecaster, bind_to_gpu_numa

    numa_cores = bind_to_gpu_numa(local_rank) if world > 1 else []   # before any pinned allocation
    L = _lib.lib()
    bg = build_model(bg_params(args.precision)).eval()
    bg.load_state_dict(make_state_dict(bg, 0, synthetic))
    pipe = BGForecastPipeline(bg)
    B = args.batch
    packed = args.input_format == "packed"

    sets = host_input_sets(args.nsets, B, 100 * rank, args.dist, synthetic, packed=packed)
    host_ref = [s[0] for s in sets]                              # reference formats (CPU baseline / parity)
    host_sets = [s[1] if packed else s[0] for s in sets]         # what our arm consumes
    pinned = [{k: v.pin_memory() for k, v in s.items()} for s in host_sets]
    dev_sets = [{k: v.to(dev) for k, v in s.items()} for s in host_sets]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host_sets[0].values())
    d2h_bytes = B * H * W

    K = args.steps
    out_maps = torch.empty((K, B, H, W), dtype=torch.uint8, device=dev)
    # the path's only collective: the per-rank label maps go to rank 0 -- issued per step as an
    # asynchronous NCCL gather so that step i's transfer runs under step i+1's kernels
    gathered = [[torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in range(world)] for _ in range(K)] \
        if (world > 1 and rank == 0) else None

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
        out_maps[i % K].copy_(out["seg"])
        if world > 1:
            dist.gather(out_maps[i % K], gathered[i % K] if rank == 0 else None, dst=0)   # NCCL connection set-up
    barrier()

    # ---- timed region 1: device-resident inputs (`value`)
    # CUDA events bracket Stage A (splat) and Stage B (net) of every step on the launching stream; no
    # events are recorded BETWEEN the net's kernels here because they would serialise the programmatic
    # dependent launches of consecutive conv layers (the per-kernel split is taken in a separate pass below).
    nsteps_net = L.pf_bgnet_num_steps(bg._net)
    ev_a = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    ev_b = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    handles = []
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
        out_maps[i].copy_(out["seg"])
        if world > 1:
            handles.append(dist.gather(out_maps[i], gathered[i] if rank == 0 else None, dst=0, async_op=True))
    for h in handles:
        h.wait()                                             # stream-level wait: the gathers are inside the timed region
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
        tr = json.load(open(TRAFFIC_FILE))
        ks = tr["per_step"]
        scale = B / float(tr["batch"])                      # the capture is one step of tr["batch"] frames
        tot = lambda pred: scale * sum(v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in ks.items() if pred(k))
        traffic_b = tot(lambda k: "zsplat" not in k)
        traffic_a = tot(lambda k: "zsplat" in k)
    except Exception:
        pass
    roof["traffic"] = traffic_b
    roof["traffic_note"] = "bytes per step (all Stage B kernels), ncu dram__bytes_read+write, " + os.path.relpath(TRAFFIC_FILE, ROOT)
    a_gbs = STAGE_A_BYTES_PER_FRAME * B / (warp_ms * 1e-3) / 1e9
    roof_a = {"bound": "hbm", "kernel": "pf_zsplat_forward_frames_hop%s (memset + point kernels per L2-sized group + one resolve)" % ("_packed" if packed else ""),
              "achieved": a_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": a_gbs / peaks["hbm_gbs"],
              "traffic": traffic_a, "ms_per_step": warp_ms}

    # ---- latency of one call at the reference export's batch sizes (device-resident inputs, not part of `value`)
    latency = {}
    for bl in (1, 2):
        one = {k: (v[:bl].contiguous() if v.dim() > 1 and v.shape[0] == B else v) for k, v in dev_sets[0].items()}
        for _ in range(3):
            pipe.forecast(one)
        ts = []
        for _ in range(15):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); pipe.forecast(one); b_.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b_))
        latency["batch%d_ms" % bl] = statistics.median(ts)

    # ---- timed region 2: end to end through the public API from pinned host buffers (`e2e`):
    # every step uploads its own inputs (pinned host -> device) and downloads its own label map; the
    # PipelinedForecaster overlaps step i+1's upload with step i's kernels (3 slots in flight).
    Ke = max(6, min(K, 20))
    pf = PipelinedForecaster(pipe, depth=3)
    for i in range(3):
        pf.submit(pinned[i % args.nsets])
    for i in range(3):
        pf.collect()
    barrier()
    checksum = 0
    e0.record()                                             # GPU idle here: timestamp = region start
    for i in range(Ke):
        pf.submit(pinned[i % args.nsets])
        if i >= 2:
            checksum += int(pf.collect()[0, 0, 0])            # the caller consumes every label map
    checksum += int(pf.collect()[0, 0, 0])
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

    cpu_base = parity = lib_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_port
        torch.set_num_threads(os.cpu_count())
        parity, t1, _ = parity_block(pipe, bg, host_ref[0], dev, args.precision)
        sd = {k: v.cpu() for k, v in bg.state_dict().items()}
        one = {k: v[:1].contiguous() for k, v in host_ref[0].items() if not k.endswith("_inv")}
        n = max(1, min(5, int(15.0 / t1)))
        t0 = time.perf_counter()
        for _ in range(n):
            cpu_port.composite_predict(sd, one, (H, W))
        dt = (time.perf_counter() - t0) / n
        cpu_base = {"value": 1.0 / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": "%d x 1 target frame @%dx%d after 1 warm-up (oracle/cpu_port.py, torch %s CPU ops); the "
                              "unmodified reference is timed by --impl reference" % (n, H, W, torch.__version__)}
    if rank == 0 and world == 1 and not args.no_library_baseline:
        try:
            lib_base = time_library_baseline(2, args.dist, 2, 5)
        except Exception as e:                               # never let the context number break the bench line
            lib_base = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:120])}

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
                "latency": latency, "parity": parity, "numa_cores_rank0": len(numa_cores),
                "clocks": clocks, "cpu_baseline": cpu_base, "library_baseline": lib_base}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
