"""On-disk formats on either side of the hot path (host-side).

Writer side == what the reference exporter produces
(experiments/export_cityscapes_segmentation_results.py:93-124):
  labels  <base>/<city>/<city>_<seq>_<frame:06d>_gtFine_labelIds.png   uint8 PNG
  depth   <base>/<city>/<city>_<seq>_<frame:06d>_depths.png            uint16 PNG = round(clamp(d+1,0,255)*256)
plus its missing-file filler (:131-166: all-255 label maps under --no_convert, zeros otherwise) and
the skip-if-exists filter that makes an interrupted export resumable (pc_transform_dataset.py:95-100).
Reader side == what BGDataset decodes (data/datasets/bg_dataset.py:172-232): label PNG -> integer
map, uint16 depth -> d = u16/256 - 1, mask = d > 0, d[~mask] = -1, clamp to [min_depth, max_depth].

The reference's bg dataset reads the depth triplet from ONE HDF5 file (`depths_decompressed_..._val.h5`, key
`<city>/<seq>/<frame:06d>/<start_fr>`, `[H, W, 3]` uint16: bg_dataset.py:184-196) that an unpublished script repacks from
the three depth-PNG exports; `repack_depth_pngs_to_h5` is that step and `read_bg_inputs_h5` the matching reader
(h5lite.py: pure-Python HDF5 subset, h5py is not in the image).

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


def repack_depth_pngs_to_h5(depth_dirs, items, h5_path):
    """items: iterable of (city, seq, frame, start_fr).  Stacks the t `..._depths.png` exports of each item (one
    directory per input frame) into `[H, W, t]` uint16 under `<city>/<seq>/<frame:06d>/<start_fr>`."""
    from . import h5lite
    tree = {}
    for city, seq, frame, start_fr in items:
        planes = [np.array(Image.open(depth_path(d, city, seq, frame))) for d in depth_dirs]
        tree.setdefault(city, {}).setdefault(seq, {}).setdefault('%06d' % frame, {})[str(start_fr)] = \
            np.stack(planes, axis=-1).astype(np.uint16)
    h5lite.write(h5_path, tree)


def read_bg_inputs_h5(label_dirs, h5_file, city, seq, frame, start_fr, min_depth=0.1, max_depth=200.0):
    """BGDataset.__getitem__ (bg_dataset.py:172-232): t label PNGs + the `[H, W, t]` uint16 depth stack of the HDF5
    container (an open h5lite.File or a path); same return value as read_bg_inputs."""
    from . import h5lite
    f = h5lite.File(h5_file) if isinstance(h5_file, str) else h5_file
    stack = f['%s/%s/%06d/%s' % (city, seq, frame, start_fr)][()]
    segs = [np.array(Image.open(label_path(ld, city, seq, frame)), dtype=np.uint8) for ld in label_dirs]
    deps, masks = [], []
    for i in range(stack.shape[-1]):
        d, m = decode_depth_u16(stack[..., i], min_depth, max_depth)
        deps.append(d)
        masks.append(m)
    return {'seg': np.stack(segs), 'depth': np.stack(deps), 'depth_mask': np.stack(masks)}
