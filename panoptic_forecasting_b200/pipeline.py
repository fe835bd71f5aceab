"""Composite bg-forecast path: per-frame reprojection (Stage A) -> disk-hop emulation -> BGModel
(Stage B), all on the device.

The reference runs these as two separate exports joined through PNG / HDF5 files
(SURVEY.md section 3.1/3.2): `task: pc_transform` with model.only_this_ind = 0,1,2 writes label
PNGs and uint16 depth PNGs (experiments/export_cityscapes_segmentation_results.py:107-124);
`task: bg` reads them back (data/datasets/bg_dataset.py:172-232).  This class keeps the same
arithmetic (including the 1/256 m depth quantisation of the disk format) without the disk.
"""
import torch

from . import _lib
from .models.bg_model import BGModel


class BGForecastPipeline:
    def __init__(self, bg_model, min_depth=None, max_depth=None, emulate_disk_hop=True):
        assert isinstance(bg_model, BGModel)
        self.bg = bg_model
        self.min_depth = bg_model.min_depth if min_depth is None else min_depth
        self.max_depth = bg_model.max_depth if max_depth is None else max_depth
        self.emulate_disk_hop = emulate_disk_hop
        self._lib = _lib.lib()
        self._ws = None

    def warp(self, inputs, fuse_hop=False):
        """All t frames reprojected into the target frame, each in its own z-buffer
        (== t reference PCTransformModel.predict calls with only_this_ind = 0..t-1).
        Returns (seg u8 [b,t,H,W], depth f32 [b,t,H,W]); with fuse_hop=True the depth is already
        disk-hop decoded and a third value, the u8 validity mask, is returned."""
        packed = 'depth_code' in inputs
        depth = inputs['depth_code'] if packed else inputs['depth']
        if not depth.is_cuda:
            raise _lib.PFError("BGForecastPipeline needs CUDA tensors (no CPU fallback)")
        dev = depth.device
        b, t, H, W = depth.shape
        if packed and not fuse_hop:
            raise ValueError("packed inputs (depth_code / depth_lut / depth_mask_bits) need fuse_hop=True")
        K = inputs['intrinsics'].to(dev, torch.float32).contiguous()
        E = inputs['extrinsics'].to(dev, torch.float32).contiguous()
        Kinv = (inputs['intrinsics_inv'].to(dev, torch.float32) if 'intrinsics_inv' in inputs
                else torch.inverse(K)).contiguous()
        Einv = (inputs['extrinsics_inv'].to(dev, torch.float32) if 'extrinsics_inv' in inputs
                else torch.inverse(E)).contiguous()
        T = inputs['target_T'].to(dev, torch.float32).contiguous()
        if packed:
            # pf_zsplat_forward_frames_hop_packed: uint16 depth codes + caller-built table + bit-packed mask
            if depth.dtype not in (torch.uint16, torch.int16) or inputs['depth_mask_bits'].dtype != torch.uint8:
                raise TypeError("depth_code must be (u)int16 and depth_mask_bits uint8")
            depth_c = depth.contiguous()
            mask_c = inputs['depth_mask_bits'].contiguous()
            lut_c = inputs['depth_lut'].to(dev, torch.float32).contiguous()
            if lut_c.numel() != 65536 or mask_c.numel() != b * t * (H * W // 8) or (H * W) % 8:
                raise ValueError("depth_lut must have 65536 entries and depth_mask_bits b*t*H*W/8 bytes")
        else:
            depth_c = depth.to(torch.float32).contiguous()
            mask = inputs['depth_mask'].contiguous()
            mask_c = mask.view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8)
        seg_c = inputs['seg'].contiguous()
        if seg_c.dtype != torch.uint8:
            raise TypeError("seg must be uint8")
        nbytes = self._lib.pf_zsplat_workspace_bytes(b, t, H, W)
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out_seg = torch.empty((b, t, H, W), dtype=torch.uint8, device=dev)
        out_depth = torch.empty((b, t, H, W), dtype=torch.float32, device=dev)
        if fuse_hop:
            out_mask = torch.empty((b, t, H, W), dtype=torch.uint8, device=dev)
            if packed:
                with torch.cuda.device(dev):
                    stream = torch.cuda.current_stream(dev).cuda_stream
                    rc = self._lib.pf_zsplat_forward_frames_hop_packed(
                        depth_c.data_ptr(), lut_c.data_ptr(), mask_c.data_ptr(), seg_c.data_ptr(), K.data_ptr(),
                        Kinv.data_ptr(), E.data_ptr(), Einv.data_ptr(), T.data_ptr(), b, t, H, W, None,
                        out_seg.data_ptr(), out_depth.data_ptr(), out_mask.data_ptr(), float(self.min_depth),
                        float(self.max_depth), self._ws.data_ptr(), self._ws.numel(), stream)
                _lib.check(rc, "pf_zsplat_forward_frames_hop_packed")
                return out_seg, out_depth, out_mask
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev).cuda_stream
                rc = self._lib.pf_zsplat_forward_frames_hop(
                    depth_c.data_ptr(), mask_c.data_ptr(), seg_c.data_ptr(), K.data_ptr(), Kinv.data_ptr(),
                    E.data_ptr(), Einv.data_ptr(), T.data_ptr(), b, t, H, W, None,
                    out_seg.data_ptr(), out_depth.data_ptr(), out_mask.data_ptr(), float(self.min_depth),
                    float(self.max_depth), self._ws.data_ptr(), self._ws.numel(), stream)
            _lib.check(rc, "pf_zsplat_forward_frames_hop")
            return out_seg, out_depth, out_mask
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = self._lib.pf_zsplat_forward_frames(
                depth_c.data_ptr(), mask_c.data_ptr(), seg_c.data_ptr(), K.data_ptr(), Kinv.data_ptr(),
                E.data_ptr(), Einv.data_ptr(), T.data_ptr(), b, t, H, W, 1, None,
                out_seg.data_ptr(), out_depth.data_ptr(), None, self._ws.data_ptr(), self._ws.numel(), stream)
        _lib.check(rc, "pf_zsplat_forward_frames")
        return out_seg, out_depth

    def decode_depth(self, depth):
        """Exporter uint16 quantisation + BGDataset decode/clamp, or (emulate_disk_hop=False) only
        the dataset's mask/clamp rule on the raw warped depth."""
        dev = depth.device
        out = torch.empty_like(depth)
        mask = torch.empty(depth.shape, dtype=torch.uint8, device=dev)
        if self.emulate_disk_hop:
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev).cuda_stream
                rc = self._lib.pf_depth_disk_hop(depth.data_ptr(), out.data_ptr(), mask.data_ptr(), depth.numel(),
                                                 float(self.min_depth), float(self.max_depth), stream)
            _lib.check(rc, "pf_depth_disk_hop")
            return out, mask
        m = depth > 0
        d = torch.where(m, depth.clamp(self.min_depth, self.max_depth), torch.full_like(depth, -1))
        return d, m.to(torch.uint8)

    def forecast(self, inputs):
        """inputs: the PCTransformModel input dict (pc_transform_model.py:27-32).
        Returns the BGModel.predict dict plus 'warped_seg' / 'warped_depth' / 'warped_mask'."""
        if self.emulate_disk_hop or 'depth_code' in inputs:
            seg, d, m = self.warp(inputs, fuse_hop=True)      # hop fused into the resolve kernel
        else:
            seg, depth = self.warp(inputs)
            d, m = self.decode_depth(depth)
        out = self.bg.predict({'seg': seg, 'depth': d, 'depth_mask': m}, {})
        out['warped_seg'], out['warped_depth'], out['warped_mask'] = seg, d, m
        return out

    def forecast_panoptic(self, inputs, mask_preds, pred_bboxes, orig_classes, pred_depths=None,
                          background_depths=None, background_depth_masks=None, use_depth_sorting=True,
                          use_bbox_ulbr=True):
        """bg forecast followed by the fg -> bg merge of `FGModel.predict_panoptic` (fg_model.py:515-518,557-588),
        both on the device: the label map never leaves HBM (the reference writes it to a PNG that FGSceneDataset
        reads back, fg_scene_dataset.py:501-510).  The instance arguments are the forecaster's outputs as in
        `panoptic.merge_instances`.  Returns the forecast dict plus 'panoptic' int64 [b, H, W]."""
        from . import panoptic
        out = self.forecast(inputs)
        out['panoptic'] = panoptic.merge_instances(
            mask_preds, pred_bboxes, orig_classes, pred_depths, background=out['seg'],
            background_depths=background_depths, background_depth_masks=background_depth_masks,
            use_depth_sorting=use_depth_sorting, use_bbox_ulbr=use_bbox_ulbr)['seg']
        return out


def bind_to_gpu_numa(device_index):
    """Pins the calling process to the CPU cores NVML reports as local to GPU `device_index` (its NUMA node), so that
    the pinned staging buffers allocated afterwards are first-touched on that node and host->device copies do not
    cross the socket interconnect.  With 8 ranks uploading ~35 GB/s each this is what the aggregate rate depends on.
    Returns the core list (empty if NVML is unavailable or reports nothing)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if visible:
            ids = [x.strip() for x in visible.split(",") if x.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cores = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        cores = [c for c in cores if c in os.sched_getaffinity(0)]
        if cores:
            os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return []


class PipelinedForecaster:
    """Host-buffer front end for throughput runs: uploads of batch i+1 (copy stream) overlap the
    forecast of batch i (compute stream) and the download of batch i-1's label map.

    Every submitted batch still pays its own host->device copy (from pinned memory) and its own
    device->host read of the uint8 label map; they are simply overlapped across batches, which is
    how an export loop over a dataset runs (reference loop: export_cityscapes_segmentation_results.py:75-110).
    """

    def __init__(self, pipe, depth=3):
        self.pipe = pipe
        self.depth = depth
        self.copy_stream = torch.cuda.Stream()
        self.compute_stream = torch.cuda.Stream()
        self.down_stream = torch.cuda.Stream()   # label-map download: off the compute stream (0.65 ms per 16-frame batch)
        self.out_dev = [None] * depth
        self.computed = [torch.cuda.Event() for _ in range(depth)]
        self.slots = [None] * depth          # device input dicts
        self.out_host = [None] * depth       # pinned uint8 label maps
        self.h2d_done = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.n_submitted = 0
        self.n_collected = 0

    def submit(self, host_inputs):
        """host_inputs: dict of pinned CPU tensors (PCTransformModel input dict)."""
        s = self.n_submitted % self.depth
        if self.n_submitted - self.n_collected >= self.depth:
            raise RuntimeError("pipeline full: collect() before submitting more")
        dev = self.pipe.bg.depth_mean.device if self.pipe.bg.use_depth_inps else torch.device("cuda")
        if self.slots[s] is None:
            self.slots[s] = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host_inputs.items()}
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[s])          # slot's previous consumer finished
            for k, v in host_inputs.items():
                self.slots[s][k].copy_(v, non_blocking=True)
            self.h2d_done[s].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.h2d_done[s])
            out = self.pipe.forecast(self.slots[s])
            self.free[s].record(self.compute_stream)
            seg = out['seg']
            self.out_dev[s] = seg                              # keeps the tensor alive until its download is enqueued
            self.computed[s].record(self.compute_stream)
        if self.out_host[s] is None or self.out_host[s].shape != seg.shape:
            self.out_host[s] = torch.empty(seg.shape, dtype=seg.dtype).pin_memory()
        with torch.cuda.stream(self.down_stream):
            self.down_stream.wait_event(self.computed[s])
            self.out_host[s].copy_(seg, non_blocking=True)
            seg.record_stream(self.down_stream)
            self.done[s].record(self.down_stream)
        self.n_submitted += 1

    def collect(self):
        """Blocks until the oldest in-flight batch is on the host; returns its pinned label map
        (valid until the slot is reused `depth` submits later)."""
        if self.n_collected >= self.n_submitted:
            raise RuntimeError("nothing in flight")
        s = self.n_collected % self.depth
        self.done[s].synchronize()
        self.n_collected += 1
        return self.out_host[s]
