"""Sharded bg-forecast export: validation items are independent, so rank r processes items
i = r (mod world) with no data-path collective; the only exchange is ONE gather of the per-rank
uint8 label maps to rank 0 at the end.

The reference's export loop is single-process (experiments/export_cityscapes_segmentation_results.py:53-127,
scripts/bg/run_export_bg_val.sh:7,17); this module is its multi-GPU counterpart for the hot path.
One process per GPU (torch.distributed, backend nccl on GPUs / gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """Items owned by `rank`: i = rank (mod world), in increasing order."""
    return list(range(rank, n_items, world))


def padded_local_count(n_items, world):
    return (n_items + world - 1) // world


class ShardedExporter:
    """forecast_fn(list_of_item_indices) -> uint8 tensor [len, H, W] on `device`."""

    def __init__(self, forecast_fn, n_items, height, width, device, batch_size=2, rank=None, world=None):
        self.forecast_fn = forecast_fn
        self.n_items = n_items
        self.h, self.w = height, width
        self.device = torch.device(device)
        self.batch_size = batch_size
        if rank is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank, self.world = rank, world
        self.local_items = shard_indices(n_items, rank, world)
        self.n_pad = padded_local_count(n_items, world)
        # every rank holds the same padded number of maps so the gather is one fixed-size collective
        self.maps = torch.zeros((self.n_pad, height, width), dtype=torch.uint8, device=self.device)

    def run(self):
        for s in range(0, len(self.local_items), self.batch_size):
            idx = self.local_items[s:s + self.batch_size]
            out = self.forecast_fn(idx)
            if out.dtype != torch.uint8 or tuple(out.shape) != (len(idx), self.h, self.w):
                raise ValueError("forecast_fn must return uint8 [%d,%d,%d]" % (len(idx), self.h, self.w))
            self.maps[s:s + len(idx)].copy_(out)
        return self

    def gather(self):
        """ONE collective.  Rank 0 returns uint8 [n_items, H, W] in item order; other ranks None."""
        if self.world == 1:
            return self.maps[:self.n_items]
        bufs = [torch.empty_like(self.maps) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(self.maps, bufs, dst=0)
        if self.rank != 0:
            return None
        out = torch.empty((self.n_items, self.h, self.w), dtype=torch.uint8, device=self.device)
        for r in range(self.world):
            items = shard_indices(self.n_items, r, self.world)
            if items:
                out[torch.as_tensor(items, device=self.device)] = bufs[r][:len(items)]
        return out


# ----------------------------------------------------------------------------------------------------------------------
# The export caller: sharded, pipelined, resumable replacement for the reference's
# `export_results` (experiments/export_cityscapes_segmentation_results.py:53-166).
# ----------------------------------------------------------------------------------------------------------------------
def _to_device(item, device, stream):
    """train_utils.batch2gpu (training/train_utils.py:45-61) on a copy stream, from pinned host memory."""
    if isinstance(item, dict):
        return {k: _to_device(v, device, stream) for k, v in item.items()}
    if isinstance(item, list):
        return [_to_device(v, device, stream) for v in item]
    if torch.is_tensor(item):
        with torch.cuda.stream(stream):
            return item.pin_memory().to(device, non_blocking=True) if not item.is_cuda else item
    return item


def export_results(model, dataset, split, params, rank=None, world=None, writer_workers=8, skip_existing=True,
                   label_lut=None):
    """Same contract as the reference loop -- reads `params` the same way (:54-74), calls `model.predict(inputs,
    labels)` under no_grad (:84-85), writes `<base>/<city>/<city>_<seq>_<target_frame:06d>_gtFine_labelIds.png`
    (:93-110) and, with `save_depth` + `save_depth_as_png`, the uint16 depth PNG (:119-124), then fills missing
    files (:131-166) -- but:
      * items are sharded: rank r handles i = r (mod world) (no collective; every rank writes its own files; rank 0
        runs the filler after a barrier);
      * the loop is pipelined: the DataLoader prefetches, the next batch's host->device copy runs on a copy stream
        under the current batch's kernels, PNG encoding runs on a thread pool;
      * it is resumable: batches whose label PNGs all exist are not forecast again (the reference has this only for
        the pc_transform dataset, pc_transform_dataset.py:95-100).
    `no_convert` must be set (as scripts/bg/run_export_bg_val.sh does) unless `label_lut` (256-entry trainId -> id
    table, the reference's `convert_labels`) is given.  Returns (n_written, n_skipped, n_filled)."""
    import numpy as np
    from torch.utils.data import DataLoader, Subset
    from . import disk_io

    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    no_gpu = params.get('no_gpu', False)
    batch_size = params['training']['batch_size']
    collate_fn = params.get('collate_fn', None)
    num_workers = params['training'].get('num_data_workers', 0)
    no_convert = params.get('no_convert')
    if params.get('viz') or params.get('is_img') or params.get('save_disp_as_png'):
        raise NotImplementedError("viz / is_img / disparity exports are outside the bg-forecast hot path")
    if not no_convert and label_lut is None:
        raise ValueError("pass --no_convert (as run_export_bg_val.sh does) or a 256-entry label_lut")
    save_depth = params.get('save_depth')
    save_png = params.get('save_depth_as_png')
    export_name = params.get('export_name')
    base = os.path.join(params['working_dir'], export_name if export_name is not None else 'exported_predictions', split)
    writer = disk_io.ExportWriter(base, workers=writer_workers, skip_existing=False)
    lut = None if label_lut is None else np.asarray(label_lut, np.uint8)

    local = shard_indices(len(dataset), rank, world)
    loader = DataLoader(Subset(dataset, local), batch_size=batch_size, collate_fn=collate_fn,
                        num_workers=num_workers, pin_memory=False)
    use_cuda = not no_gpu
    copy_stream = torch.cuda.Stream() if use_cuda else None
    n_written = n_skipped = 0

    def stage(batch):
        """host batch -> (device inputs, labels, meta, event): the copy is enqueued, not waited for"""
        meta = batch['meta']
        names = [(meta['city'][i], meta['seq'][i], int(meta['target_frame'][i])) for i in range(len(meta['city']))]
        if skip_existing and all(writer.exists(*n) for n in names):
            return None, names
        if not use_cuda:
            return (batch['inputs'], batch['labels'], None), names
        dev = torch.device('cuda', torch.cuda.current_device())
        inputs = _to_device(batch['inputs'], dev, copy_stream)
        labels = _to_device(batch['labels'], dev, copy_stream)
        ev = torch.cuda.Event()
        ev.record(copy_stream)
        return (inputs, labels, ev), names

    def finish(preds, names):
        nonlocal n_written
        seg = preds['seg']
        seg = seg.to('cpu', non_blocking=False) if torch.is_tensor(seg) else seg
        seg = seg.numpy().astype(np.uint8)
        depth = preds['depth'].float().cpu().numpy() if (save_depth and 'depth' in preds) else None
        for i, (city, seq, frame) in enumerate(names):
            s = seg[i] if lut is None else lut[seg[i]]
            if depth is not None and not save_png:
                os.makedirs(os.path.join(base, city), exist_ok=True)
                np.save(os.path.join(base, city, '%s_%s_%06d_depths.npy' % (city, seq, frame)), depth[i])
            writer.submit(s, city, seq, frame, depth=depth[i] if (depth is not None and save_png) else None)
            n_written += 1

    it = iter(loader)
    nxt = stage(next(it)) if len(local) else None
    while nxt is not None:
        cur = nxt
        try:
            nxt = stage(next(it))                    # next batch's copy overlaps this batch's kernels
        except StopIteration:
            nxt = None
        staged, names = cur
        if staged is None:
            n_skipped += len(names)
            continue
        inputs, labels, ev = staged
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        with torch.no_grad():
            preds = model.predict(inputs, labels)
        finish(preds, names)
    writer.close()

    if world > 1 and dist.is_available() and dist.is_initialized():
        dist.barrier()
    n_filled = 0
    cityscapes_dir = params['data'].get('cityscapes_dir')
    if rank == 0 and cityscapes_dir is not None:
        import glob
        gt_dir = os.path.join(cityscapes_dir, 'gtFine', getattr(dataset, 'split', split))
        cities = params['data'].get('cities')
        filler = disk_io.ExportWriter(base, workers=1)
        for city in sorted(os.listdir(gt_dir)):
            if cities is not None and city not in cities:
                continue
            for path in sorted(glob.glob(os.path.join(gt_dir, city, '*_gtFine_labelIds.png'))):
                out_name = os.path.join(base, city, os.path.basename(path))
                if not os.path.exists(out_name):
                    # :146-163 (no background_dir on the bg path): all 255 under --no_convert, zeros otherwise
                    blank = np.full((1024, 2048), 255 if no_convert else 0, dtype=np.uint8)
                    filler._save(blank, out_name)
                    n_filled += 1
        filler.close()
    return n_written, n_skipped, n_filled
