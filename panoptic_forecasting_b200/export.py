"""Sharded bg-forecast export: validation items are independent, so rank r processes items
i = r (mod world) with no data-path collective; the only exchange is ONE gather of the per-rank
uint8 label maps to rank 0 at the end.

The reference's export loop is single-process (experiments/export_cityscapes_segmentation_results.py:53-127,
scripts/bg/run_export_bg_val.sh:7,17); this module is its multi-GPU counterpart for the hot path.
One process per GPU (torch.distributed, backend nccl on GPUs / gloo in the CPU tests).
"""
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
