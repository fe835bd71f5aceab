"""CPU: the export caller (panoptic_forecasting_b200.export.export_results) -- the sharded, resumable replacement of
the reference loop experiments/export_cityscapes_segmentation_results.py:53-166 -- at world size 1 and 2 (gloo): the
PNG set is byte-identical, an interrupted export resumes without recomputing, the missing-file filler writes the
reference's blank maps.  The model is a CPU stand-in with the reference `predict` surface; the CUDA models are driven
through the same loop by tests/test_export_gpu.py."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from PIL import Image

from panoptic_forecasting_b200 import export


class FakeItems(torch.utils.data.Dataset):
    """Items shaped like BGDataset's (bg_dataset.py:235-261): {'inputs','labels','meta'}."""
    split = "val"

    def __init__(self, n, h=12, w=20):
        self.n, self.h, self.w = n, h, w

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(i)
        seg = torch.randint(0, 19, (3, self.h, self.w), generator=g, dtype=torch.uint8)
        return {"inputs": {"seg": seg}, "labels": {},
                "meta": {"city": "city%d" % (i % 3), "seq": "%06d" % i, "frame": i, "target_frame": 19 + i}}


class FakeModel:
    calls = 0

    def predict(self, inputs, labels):
        FakeModel.calls += 1
        seg = (inputs["seg"].long().sum(1) * 7 + 3) % 19
        return {"seg": seg, "depth": seg.float() * 1.5 - 1}


def params_for(tmp, **extra):
    p = {"no_gpu": True, "training": {"batch_size": 2, "num_data_workers": 0}, "working_dir": str(tmp),
         "no_convert": True, "viz": False, "is_img": False, "save_depth": False, "save_disp_as_png": False,
         "save_depth_as_png": False, "disp_factor": None, "export_name": "exported", "data": {"num_classes": 11}}
    p.update(extra)
    return p


def digest(base):
    out = {}
    for root, _, files in os.walk(base):
        for f in sorted(files):
            path = os.path.join(root, f)
            out[os.path.relpath(path, base)] = hashlib.sha1(open(path, "rb").read()).hexdigest()
    return out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        export.export_results(FakeModel(), FakeItems(n), "val", params_for(tmp, save_depth=True, save_depth_as_png=True))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 2])
def test_world2_png_set_is_byte_identical_to_world1(tmp_path, n):
    a, b = tmp_path / "w1", tmp_path / "w2"
    export.export_results(FakeModel(), FakeItems(n), "val", params_for(a, save_depth=True, save_depth_as_png=True), rank=0, world=1)
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, b, n)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    da, db = digest(a / "exported" / "val"), digest(b / "exported" / "val")
    assert len(da) == 2 * n and da == db
    name = "city1/city1_000001_000020_gtFine_labelIds.png"
    assert name in da                                       # reference naming: <city>_<seq>_<target_frame:06d>
    item = FakeItems(n)[1]
    want = ((item["inputs"]["seg"].long().sum(0) * 7 + 3) % 19).numpy().astype(np.uint8)
    assert np.array_equal(np.array(Image.open(a / "exported" / "val" / name)), want)
    d = np.array(Image.open(a / "exported" / "val" / "city1/city1_000001_000020_depths.png"))
    assert d.dtype == np.uint16 and np.array_equal(d, np.rint(np.clip(want * 1.5, 0, 255) * 256).astype(np.uint16))


def test_resume_skips_finished_batches_and_filler_writes_blanks(tmp_path):
    n = 6
    p = params_for(tmp_path)
    FakeModel.calls = 0
    w, s, f = export.export_results(FakeModel(), FakeItems(n), "val", p, rank=0, world=1)
    assert (w, s, f) == (n, 0, 0) and FakeModel.calls == 3
    base = tmp_path / "exported" / "val"
    before = digest(base)
    os.remove(base / "city2" / "city2_000002_000021_gtFine_labelIds.png")      # an interrupted export: one file missing
    FakeModel.calls = 0
    w, s, f = export.export_results(FakeModel(), FakeItems(n), "val", p, rank=0, world=1)
    assert FakeModel.calls == 1 and w == 2 and s == 4                          # only the batch holding item 2 is redone
    assert digest(base) == before
    # missing-file filler (:131-166): ground-truth names without a prediction get an all-255 map under --no_convert
    gt = tmp_path / "cs" / "gtFine" / "val" / "city0"
    gt.mkdir(parents=True)
    for name in ("city0_000000_000019_gtFine_labelIds.png", "city0_000099_000019_gtFine_labelIds.png"):
        Image.fromarray(np.zeros((4, 4), np.uint8)).save(gt / name)
    p["data"]["cityscapes_dir"] = str(tmp_path / "cs")
    w, s, f = export.export_results(FakeModel(), FakeItems(n), "val", p, rank=0, world=1)
    assert (w, s, f) == (0, n, 1)
    blank = np.array(Image.open(base / "city0" / "city0_000099_000019_gtFine_labelIds.png"))
    assert blank.shape == (1024, 2048) and (blank == 255).all()


def test_requires_no_convert_or_lut(tmp_path):
    p = params_for(tmp_path, no_convert=False)
    with pytest.raises(ValueError):
        export.export_results(FakeModel(), FakeItems(2), "val", p, rank=0, world=1)
    lut = (np.arange(256) + 7).astype(np.uint8)
    export.export_results(FakeModel(), FakeItems(2), "val", p, rank=0, world=1, label_lut=lut)
    got = np.array(Image.open(tmp_path / "exported" / "val" / "city0" / "city0_000000_000019_gtFine_labelIds.png"))
    item = FakeItems(2)[0]
    assert np.array_equal(got, ((item["inputs"]["seg"].long().sum(0) * 7 + 3) % 19).numpy().astype(np.uint8) + 7)
