"""GPU parity: Stage A (pf_zsplat_forward through PCTransformModel / the C ABI) vs the oracle.
Integer/byte outputs and the winning depths must be BIT-EXACT."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import pc_params
from oracle import pc_transform_oracle
from panoptic_forecasting_b200 import synthetic
from panoptic_forecasting_b200.models import build_model
from test_oracle import load_pc_case

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def run_gpu(inp_np, ind, is_img=False, **extra):
    model = build_model(dict(pc_params(ind, is_img or None, **extra), no_gpu=False))
    inp = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in inp_np.items()}
    out = model.predict(inp, {})
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def with_inverses(inp_t):
    npin = {k: v.numpy() for k, v in inp_t.items()}
    npin["intrinsics_inv"] = torch.inverse(inp_t["intrinsics"]).numpy()
    npin["extrinsics_inv"] = torch.inverse(inp_t["extrinsics"]).numpy()
    return npin


def assert_same(out, ref):
    assert np.array_equal(out["seg"], ref["seg"])
    assert np.array_equal(out["depth"].view(np.uint32), ref["depth"].view(np.uint32))
    if "result2d" in out:
        assert np.array_equal(out["result2d"], ref["result2d"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "pc_*.npz"))), ids=os.path.basename)
def test_golden_vectors_bit_exact(pf_lib, path):
    inp, ind, is_img, z = load_pc_case(path)
    out = run_gpu(inp, ind, is_img)
    ref = {"seg": z["out_seg"], "depth": z["out_depth"], "result2d": z["out_result2d"].astype(np.int64)}
    assert_same(out, ref)


@pytest.mark.parametrize("dist", ["R", "U"])
@pytest.mark.parametrize("ind", [0, 2, None])
@pytest.mark.parametrize("shape", [(2, 3, 96, 160), (1, 3, 256, 512), (1, 2, 50, 67)])
def test_matches_oracle_bit_exact(pf_lib, dist, ind, shape):
    b, t, h, w = shape
    if ind is not None and ind >= t:
        pytest.skip("frame index outside t")
    npin = with_inverses(synthetic.make_pc_inputs(b=b, t=t, h=h, w=w, dist=dist, seed=100 + h))
    assert_same(run_gpu(npin, ind), pc_transform_oracle.predict(npin, only_this_ind=ind))


def test_edge_cases_bit_exact(pf_lib):
    """ties, invalid-only cells, points behind the camera, everything masked, 1x1 image,
    clamp-to-border (large ego motion), floor==ceil duplicates (identity warp)."""
    rng = np.random.default_rng(0)
    H, W = 16, 24
    eye = lambda n: np.eye(n, dtype=np.float32)[None]
    base = {"intrinsics": eye(3), "extrinsics": eye(4), "target_T": np.eye(4, dtype=np.float32)[None, None].repeat(2, 1),
            "depth": np.full((1, 2, H, W), 5.0, np.float32), "depth_mask": np.ones((1, 2, H, W), bool),
            "seg": rng.integers(1, 19, (1, 2, H, W), dtype=np.uint8)}
    cases = [dict(base)]                                           # identity warp: integer coords, frame ties
    c = dict(base); c["depth_mask"] = np.zeros((1, 2, H, W), bool); cases.append(c)      # nothing valid
    c = dict(base); c["depth"] = -base["depth"]; cases.append(c)                          # behind the camera
    c = dict(base); T = base["target_T"].copy(); T[0, :, 0, 3] = 40.0; c["target_T"] = T; cases.append(c)  # off-image -> clamp
    c = dict(base); m = np.ones((1, 2, H, W), bool); m[0, :, ::2] = False; c["depth_mask"] = m; cases.append(c)
    c = dict(base); d = base["depth"].copy(); d[0, 1] = 4.0; c["depth"] = d; cases.append(c)  # frame 1 nearer
    c = dict(base); E = base["extrinsics"].copy(); E[0, 3] = [0.0, 0.0, 0.01, 1.0]; c["extrinsics"] = E   # projective last row:
    c["extrinsics_inv"] = np.linalg.inv(E).astype(np.float32); c["intrinsics_inv"] = eye(3); cases.append(c)  # w != 1 path
    for c in cases:
        for ind in (None, 1):
            assert_same(run_gpu(c, ind), pc_transform_oracle.predict(c, only_this_ind=ind))
    one = {"intrinsics": eye(3), "extrinsics": eye(4), "target_T": np.eye(4, dtype=np.float32)[None, None],
           "depth": np.full((1, 1, 1, 1), 2.0, np.float32), "depth_mask": np.ones((1, 1, 1, 1), bool),
           "seg": np.full((1, 1, 1, 1), 9, np.uint8)}
    assert_same(run_gpu(one, None), pc_transform_oracle.predict(one))


def test_label_lut_and_no_coords(pf_lib):
    npin = with_inverses(synthetic.make_pc_inputs(b=1, t=3, h=64, w=128, dist="R", seed=5))
    lut = (np.arange(256) * 7 % 256).astype(np.uint8)
    out = run_gpu(npin, 0, label_lut=lut, return_result2d=False)
    ref = pc_transform_oracle.predict(npin, only_this_ind=0, lut=lut)
    assert "result2d" not in out
    assert_same(out, ref)


def test_host_buffer_entry_point(pf_lib):
    """pf_zsplat_forward_host (host-buffer convenience call) == device-pointer path == oracle."""
    npin = with_inverses(synthetic.make_pc_inputs(b=2, t=1, h=64, w=96, dist="U", seed=9))
    ref = pc_transform_oracle.predict(npin)
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    d, mk, sg = c(npin["depth"], np.float32), c(npin["depth_mask"], np.uint8), c(npin["seg"], np.uint8)
    K, Ki = c(npin["intrinsics"], np.float32), c(npin["intrinsics_inv"], np.float32)
    E, Ei = c(npin["extrinsics"], np.float32), c(npin["extrinsics_inv"], np.float32)
    T = c(npin["target_T"], np.float32)
    oseg = np.empty((2, 64, 96), np.uint8)
    odep = np.empty((2, 64, 96), np.float32)
    p = lambda a: a.ctypes.data
    rc = pf_lib.pf_zsplat_forward_host(p(d), p(mk), p(sg), p(K), p(Ki), p(E), p(Ei), p(T), 2, 1, 64, 96, 1, None,
                                       p(oseg), p(odep))
    assert rc == 0, pf_lib.pf_last_error()
    assert_same({"seg": oseg, "depth": odep}, ref)


@pytest.mark.parametrize("dist", ["R", "U"])
@pytest.mark.parametrize("ind", [0, None])
@pytest.mark.parametrize("shape", [(2, 3, 64, 128), (1, 3, 128, 256), (3, 2, 32, 384), (2, 3, 100, 96), (1, 2, 4, 32)])
@pytest.mark.parametrize("l2_mb", [None, "1"])
def test_fast_point_kernel_bit_exact(pf_lib, monkeypatch, dist, ind, shape, l2_mb):
    """W % 32 == 0, H % 4 == 0 and no result2d: the packed-FFMA2 point kernel with horizontal (warp shuffle) and
    vertical (in-thread) candidate merging, the slab reused by several groups (PF_ZSPLAT_L2_MB=1: one z-buffer per
    group, so the call-wide sentinel is patched in after the last group) -- same bits as the oracle, and as the generic
    kernel (the default; PF_ZSPLAT_FAST=1 selects the FFMA2 kernel) in both work-space schemes."""
    b, t, h, w = shape
    if l2_mb is not None:
        monkeypatch.setenv("PF_ZSPLAT_L2_MB", l2_mb)
    monkeypatch.setenv("PF_ZSPLAT_FAST", "1")
    monkeypatch.setenv("PF_ZSPLAT_MODE", "slab")
    npin = with_inverses(synthetic.make_pc_inputs(b=b, t=t, h=h, w=w, dist=dist, seed=7 + h))
    out = run_gpu(npin, ind, return_result2d=False)
    assert_same(out, pc_transform_oracle.predict(npin, only_this_ind=ind))
    monkeypatch.delenv("PF_ZSPLAT_FAST")
    assert_same(run_gpu(npin, ind, return_result2d=False), out)
    monkeypatch.setenv("PF_ZSPLAT_FAST", "1")
    monkeypatch.setenv("PF_ZSPLAT_MODE", "full")
    assert_same(run_gpu(npin, ind, return_result2d=False), out)


@pytest.mark.parametrize("ind", [1, None])
@pytest.mark.parametrize("l2_mb", [None, "1"])
def test_full_mode_bit_exact(pf_lib, monkeypatch, ind, l2_mb):
    """PF_ZSPLAT_MODE=full: a z-buffer per plane of the call, point kernels per L2-sized group, one resolve at the
    end with the sentinel known (no patch pass) -- the A/B alternative to the slab scheme."""
    if l2_mb is not None:
        monkeypatch.setenv("PF_ZSPLAT_L2_MB", l2_mb)
    npin = with_inverses(synthetic.make_pc_inputs(b=3, t=3, h=64, w=128, dist="R", seed=31))
    monkeypatch.setenv("PF_ZSPLAT_MODE", "slab")
    assert_same(run_gpu(npin, ind, return_result2d=False), pc_transform_oracle.predict(npin, only_this_ind=ind))
    monkeypatch.setenv("PF_ZSPLAT_MODE", "full")
    assert_same(run_gpu(npin, ind, return_result2d=False), pc_transform_oracle.predict(npin, only_this_ind=ind))
    assert_same(run_gpu(npin, ind), pc_transform_oracle.predict(npin, only_this_ind=ind))       # generic kernel + coords


def hop_np(d, mn=0.1, mx=200.0):
    q = ((torch.from_numpy(d) + 1).clamp(0, 255) * 256).round().numpy().astype(np.uint16)
    r = torch.from_numpy(q.astype(np.float32)) / 256.0 - 1
    m = r > 0
    r[~m] = -1
    r[m & (r > mx)] = mx
    r[m & (r < mn)] = mn
    return r.numpy(), m.numpy()


@pytest.mark.parametrize("dist", ["R", "U"])
@pytest.mark.parametrize("shape", [(2, 3, 64, 128), (2, 3, 40, 96), (2, 3, 30, 72), (1, 3, 256, 512)])
@pytest.mark.parametrize("fast,pipe", [("0", "1"), ("0", "0"), ("1", "1")], ids=["generic-pipelined", "generic-plain", "fast"])
def test_packed_inputs_bit_exact(pf_lib, bg_shapes, monkeypatch, dist, shape, fast, pipe):
    """pf_zsplat_forward_frames_hop_packed (uint16 depth code + table, 1-bit mask) == the reference-format entry
    point on depth = lut[code] == oracle + disk hop, bit for bit (fast kernel for W % 128 == 0, generic otherwise)."""
    from conftest import bg_params
    from panoptic_forecasting_b200.pipeline import BGForecastPipeline
    monkeypatch.setenv("PF_ZSPLAT_FAST", fast)
    monkeypatch.setenv("PF_ZSPLAT_PIPE", pipe)       # pipelined code / table loads of the generic point kernel (default on)
    b, t, h, w = shape
    packed, unpacked = synthetic.pack_pc_inputs(synthetic.make_pc_inputs(b=b, t=t, h=h, w=w, dist=dist, seed=3 + w))
    npin = with_inverses(unpacked)
    bg = build_model(dict(bg_params(), no_gpu=False)).eval()
    pipe = BGForecastPipeline(bg)
    inv = {"intrinsics_inv": torch.from_numpy(npin["intrinsics_inv"]), "extrinsics_inv": torch.from_numpy(npin["extrinsics_inv"])}
    got_p = [x.cpu().numpy() for x in pipe.warp({k: v.cuda() for k, v in dict(packed, **inv).items()}, fuse_hop=True)]
    got_u = [x.cpu().numpy() for x in pipe.warp({k: v.cuda() for k, v in dict(unpacked, **inv).items()}, fuse_hop=True)]
    for a, c in zip(got_p, got_u):
        assert np.array_equal(a.view(np.uint8), c.view(np.uint8))
    for ind in range(t):
        ref = pc_transform_oracle.predict(npin, only_this_ind=ind)
        d, m = hop_np(ref["depth"])
        assert np.array_equal(got_p[0][:, ind], ref["seg"])
        assert np.array_equal(got_p[1][:, ind].view(np.uint32), d.view(np.uint32))
        assert np.array_equal(got_p[2][:, ind].astype(bool), m)


def test_full_size_properties(pf_lib):
    """BASELINE size (1024x2048): size-independent properties instead of the (slow) oracle:
    idempotence of the call, agreement of t=1 slices with only_this_ind, depth/label consistency,
    and a checksum against the oracle on a 128-row band rendered separately."""
    inp = synthetic.make_pc_inputs(b=1, t=3, h=1024, w=2048, dist="R", seed=0)
    npin = with_inverses(inp)
    a = run_gpu(npin, 0)
    b = run_gpu(npin, 0)
    assert_same(a, b)                                              # deterministic despite atomics
    sl = {k: (v[:, 0:1] if k in ("depth", "depth_mask", "target_T", "seg") else v) for k, v in npin.items()}
    assert_same(run_gpu(sl, None), a)                              # only_this_ind == explicit t=1 slice
    hit = a["depth"] >= 0
    assert hit.mean() > 0.5
    assert (a["seg"][~hit] == 0).all()
    sentinel = a["depth"].max()
    inv_only = a["depth"] == sentinel
    assert (a["seg"][inv_only] == 0).all()
    ref = pc_transform_oracle.predict(npin, only_this_ind=0)       # ~3 s on one core
    assert_same(a, ref)


def test_disk_hop(pf_lib):
    g = torch.Generator().manual_seed(0)
    d = (torch.rand(100000, generator=g) * 300 - 20)
    d[:10] = torch.tensor([-1.0, 0.0, 0.001953125, 0.1, 199.99, 200.0, 254.0, 255.0, 1e9, -5.0])
    x = d.cuda()
    out = torch.empty_like(x)
    m = torch.empty(x.shape, dtype=torch.uint8, device="cuda")
    rc = pf_lib.pf_depth_disk_hop(x.data_ptr(), out.data_ptr(), m.data_ptr(), x.numel(), 0.1, 200.0, None)
    assert rc == 0
    torch.cuda.synchronize()
    # exporter + BGDataset arithmetic (export_cityscapes_segmentation_results.py:119-122, bg_dataset.py:223-230,166-170)
    q = ((d + 1).clamp(0, 255) * 256).round().numpy().astype(np.uint16)
    r = torch.from_numpy(q.astype(np.float32)) / 256.0 - 1
    mask = r > 0
    r[~mask] = -1
    r[mask & (r > 200.0)] = 200.0
    r[mask & (r < 0.1)] = 0.1
    assert torch.equal(out.cpu(), r)
    assert torch.equal(m.cpu().bool(), mask)
