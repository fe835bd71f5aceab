"""CPU: the HDF5 subset behind the bg dataset's depth container (panoptic_forecasting_b200/h5lite.py; reference
data/datasets/bg_dataset.py:184-196 reads it through h5py, which this image does not have)."""
import os

import numpy as np
import pytest
from PIL import Image

from panoptic_forecasting_b200 import disk_io, h5lite

SCIPY_H5 = os.path.join(os.path.dirname(np.__file__), "..", "scipy", "io", "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")


def test_reads_a_file_written_by_libhdf5():
    """scipy ships a MATLAB 7.3 file = HDF5 written by libhdf5 1.6 behind a 512-byte user block (superblock v0, base
    address, symbol-table group, v1 object header, layout message v2): variable `testdouble` = 0:pi/4:2*pi."""
    if not os.path.exists(SCIPY_H5):
        pytest.skip("scipy test data not installed")
    f = h5lite.File(SCIPY_H5)
    assert f.keys() == ["testdouble"]
    d = f["testdouble"]
    assert d.shape == (9, 1) and d.dtype == np.dtype("<f8")
    assert np.array_equal(d[()].ravel(), np.arange(9) * (np.pi / 4))


@pytest.mark.parametrize("chunks,deflate,shuffle", [(None, False, False), ((5, 8, 3), False, False), ((4, 16, 2), True, True)])
def test_round_trip_groups_dtypes_layouts(tmp_path, chunks, deflate, shuffle):
    rng = np.random.default_rng(1)
    tree = {"aachen": {"%06d" % s: {"000019": {"4": rng.integers(0, 65536, (11, 20, 3), dtype=np.uint16)}} for s in range(3)},
            "f32": rng.normal(size=(7,)).astype(np.float32), "i64": np.arange(-4, 4, dtype=np.int64),
            "u8": rng.integers(0, 255, (3, 5), dtype=np.uint8), "f64": rng.normal(size=(2, 2, 2, 2))}
    for i in range(50):                                           # more entries than libhdf5's default leaf node holds
        tree["g%03d" % i] = {"x": np.full((2,), i, np.int32)}
    path = str(tmp_path / "t.h5")
    h5lite.write(path, tree, chunks=chunks, deflate=deflate, shuffle=shuffle)
    f = h5lite.File(path)
    assert len(f.keys()) == 55 and "aachen" in f and "nope" not in f
    for s in range(3):
        ref = tree["aachen"]["%06d" % s]["000019"]["4"]
        got = f["aachen/%06d/000019/4" % s]
        assert got.shape == ref.shape and got.dtype == ref.dtype and np.array_equal(got[()], ref)
        assert np.array_equal(h5lite.read_bg_depth(path, "aachen", "%06d" % s, 19, 4), ref)
    for k in ("f32", "i64", "u8", "f64"):
        assert np.array_equal(f[k][()], tree[k]) and f[k].dtype == tree[k].dtype
    assert int(f["g049/x"][()][1]) == 49
    with pytest.raises(KeyError):
        f["aachen/000009"]


def test_depth_png_exports_repacked_into_the_bg_container(tmp_path):
    """exporter PNGs (ind0/1/2) -> [H,W,3] uint16 HDF5 -> BGDataset decode == the PNG reader's result."""
    rng = np.random.default_rng(2)
    dirs = [str(tmp_path / ("ind%d" % i)) for i in range(3)]
    depth = rng.uniform(-3, 280, size=(3, 24, 40)).astype(np.float32)
    seg = rng.integers(0, 19, size=(3, 24, 40), dtype=np.uint8)
    for i, d in enumerate(dirs):
        w = disk_io.ExportWriter(d, workers=1)
        w.submit(seg[i], "ulm", "000007", 19, depth=depth[i])
        w.close()
    h5 = str(tmp_path / "depths_val.h5")
    disk_io.repack_depth_pngs_to_h5(dirs, [("ulm", "000007", 19, 4)], h5)
    a = disk_io.read_bg_inputs(dirs, dirs, "ulm", "000007", 19)
    b = disk_io.read_bg_inputs_h5(dirs, h5, "ulm", "000007", 19, 4)
    for k in ("seg", "depth", "depth_mask"):
        assert np.array_equal(a[k], b[k])
    assert h5lite.File(h5)["ulm/000007/000019/4"].shape == (24, 40, 3)
