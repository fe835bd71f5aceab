"""CPU: the on-disk formats either side of the hot path (writer naming / uint16 depth quantisation /
skip-if-exists / missing-file filler / BGDataset decode), checked against the reference arithmetic
restated in the oracle's disk hop."""
import os

import numpy as np
import torch
from PIL import Image

from oracle import cpu_port
from panoptic_forecasting_b200 import disk_io


def test_depth_png_round_trip_equals_oracle_disk_hop(tmp_path):
    rng = np.random.default_rng(0)
    depth = (rng.uniform(-3, 300, size=(3, 40, 64))).astype(np.float32)
    depth[0, 0, :6] = [-1.0, 0.0, 0.001953125, 199.99, 254.0, 1e9]
    seg = rng.integers(0, 19, size=(3, 40, 64), dtype=np.uint8)
    dirs = [str(tmp_path / ("ind%d" % i)) for i in range(3)]
    for i, d in enumerate(dirs):
        w = disk_io.ExportWriter(d, workers=2)
        w.submit(seg[i], "aachen", "000001", 19, depth=depth[i])
        w.close()
        assert os.path.basename(disk_io.label_path(d, "aachen", "000001", 19)) == "aachen_000001_000019_gtFine_labelIds.png"
        assert np.array(Image.open(disk_io.depth_path(d, "aachen", "000001", 19))).dtype == np.uint16
    got = disk_io.read_bg_inputs(dirs, dirs, "aachen", "000001", 19)
    ref_d, ref_m = cpu_port.disk_hop(torch.from_numpy(depth))
    assert np.array_equal(got["seg"], seg)
    assert np.array_equal(got["depth"], ref_d.numpy())
    assert np.array_equal(got["depth_mask"], ref_m.numpy())


def test_skip_existing_and_missing_filler(tmp_path):
    base = str(tmp_path / "out")
    a = np.full((8, 16), 3, np.uint8)
    w = disk_io.ExportWriter(base, workers=1)
    w.submit(a, "ulm", "000002", 7)
    w.close()
    w2 = disk_io.ExportWriter(base, workers=1, skip_existing=True)
    w2.submit(np.zeros((8, 16), np.uint8), "ulm", "000002", 7)          # already there -> untouched
    w2.submit(a + 1, "ulm", "000003", 7)
    n = w2.fill_missing([("ulm", "000002", 7), ("ulm", "000004", 7)], height=8, width=16, no_convert=True)
    w2.close()
    assert w2.n_skipped == 1 and n == 1
    assert (np.array(Image.open(disk_io.label_path(base, "ulm", "000002", 7))) == 3).all()
    assert (np.array(Image.open(disk_io.label_path(base, "ulm", "000003", 7))) == 4).all()
    assert (np.array(Image.open(disk_io.label_path(base, "ulm", "000004", 7))) == 255).all()
