"""On-disk formats on either side of the hot path (host-side).

Writer side == what the reference exporter produces
(experiments/export_cityscapes_segmentation_results.py:93-124):
  labels  <base>/<city>/<city>_<seq>_<frame:06d>_gtFine_labelIds.png   uint8 PNG
  depth   <base>/<city>/<city>_<seq>_<frame:06d>_depths.png            uint16 PNG = round(clamp(d+1,0,255)*256)
plus its missing-file filler (:131-166: all-255 label maps under --no_convert, zeros otherwise) and
the skip-if-exists filter that makes an interrupted export resumable (pc_transform_dataset.py:95-100).
Reader side == what BGDataset decodes (data/datasets/bg_dataset.py:172-232): label PNG -> integer
map, uint16 depth -> d = u16/256 - 1, mask = d > 0, d[~mask] = -1, clamp to [min_depth, max_depth].

PNG encoding runs on a thread pool so it overlaps the GPU work of the next batches.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
from PIL import Image


def label_path(base_dir, city, seq, frame):
    return os.path.join(base_dir, city, '%s_%s_%06d_gtFine_labelIds.png' % (city, seq, frame))


def depth_path(base_dir, city, seq, frame):
    return os.path.join(base_dir, city, '%s_%s_%06d_depths.png' % (city, seq, frame))


def encode_depth_u16(depth):
    """float32 depth -> exporter's uint16 (torch.round == round-half-even == np.rint)."""
    d = np.asarray(depth, np.float32)
    return np.rint(np.clip(d + np.float32(1), 0, 255) * np.float32(256)).astype(np.uint16)


def decode_depth_u16(u16, min_depth=0.1, max_depth=200.0):
    """BGDataset decode: returns (depth float32, mask bool)."""
    d = u16.astype(np.float32) / np.float32(256.0) - np.float32(1)
    mask = d > 0
    d[~mask] = -1
    d[mask & (d > max_depth)] = max_depth
    d[mask & (d < min_depth)] = min_depth
    return d, mask


class ExportWriter:
    """Asynchronous writer of label / depth PNGs with the reference's naming."""

    def __init__(self, base_dir, workers=8, skip_existing=False):
        self.base_dir = base_dir
        self.skip_existing = skip_existing
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.futures = []
        self.n_skipped = 0

    def exists(self, city, seq, frame):
        return os.path.exists(label_path(self.base_dir, city, seq, frame))

    @staticmethod
    def _save(arr, path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + '.tmp.png'
        Image.fromarray(arr).save(tmp)
        os.replace(tmp, path)          # atomic: a killed export never leaves a truncated PNG behind

    def submit(self, seg_u8, city, seq, frame, depth=None):
        """seg_u8: [H,W] uint8 numpy (a private copy is taken); depth: optional [H,W] float32."""
        if self.skip_existing and self.exists(city, seq, frame):
            self.n_skipped += 1
            return
        self.futures.append(self.pool.submit(self._save, np.array(seg_u8, dtype=np.uint8, copy=True),
                                             label_path(self.base_dir, city, seq, frame)))
        if depth is not None:
            self.futures.append(self.pool.submit(self._save, encode_depth_u16(depth),
                                                 depth_path(self.base_dir, city, seq, frame)))

    def fill_missing(self, expected, height=1024, width=2048, no_convert=True):
        """expected: iterable of (city, seq, frame).  Writes the reference's blank map for absent files."""
        blank = np.full((height, width), 255 if no_convert else 0, dtype=np.uint8)
        n = 0
        for city, seq, frame in expected:
            if not self.exists(city, seq, frame):
                self._save(blank, label_path(self.base_dir, city, seq, frame))
                n += 1
        return n

    def close(self):
        for f in self.futures:
            f.result()
        self.futures = []
        self.pool.shutdown(wait=True)


def read_bg_inputs(label_dirs, depth_dirs, city, seq, frame, min_depth=0.1, max_depth=200.0):
    """Reads the t reprojected label PNGs (one directory per input frame, the `..._ind{i}_all` layout
    of configs/bg/bg_val_mid.yaml:12-14) and the matching uint16 depth PNGs; returns the BGModel input
    dict as numpy arrays: seg uint8 [t,H,W], depth float32 [t,H,W], depth_mask bool [t,H,W]."""
    segs, deps, masks = [], [], []
    for ld, dd in zip(label_dirs, depth_dirs):
        segs.append(np.array(Image.open(label_path(ld, city, seq, frame)), dtype=np.uint8))
        d, m = decode_depth_u16(np.array(Image.open(depth_path(dd, city, seq, frame))), min_depth, max_depth)
        deps.append(d)
        masks.append(m)
    return {'seg': np.stack(segs), 'depth': np.stack(deps), 'depth_mask': np.stack(masks)}
