"""GPU parity of the tensor-core (tcgen05, split-bf16 3-pass) Stage B path vs the oracle and vs
the fp32 SIMT path.  Tolerances: north_star asks <= 1e-3 relative on logits; the split-bf16
scheme is expected (oracle/precision study in DESIGN.md) to land near 3e-5, asserted <= 3e-4."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from oracle import bg_oracle
from panoptic_forecasting_b200 import _lib, synthetic
from test_bgnet_gpu import check_against, gpu_model

pytestmark = pytest.mark.gpu


def layer_sweep(pf_lib, bg_shapes, tol, hw=(24, 40), b=2):
    sd = synthetic.make_bg_state_dict(bg_shapes, seed=2)
    m = gpu_model(sd, precision="tc")
    m._upload(torch.device("cuda", torch.cuda.current_device()))
    n = pf_lib.pf_bgnet_num_convs(m._net)
    info = _lib.ConvInfo()
    g = torch.Generator().manual_seed(0)
    worst = 0.0
    for i in range(1, n + 1):
        assert pf_lib.pf_bgnet_conv_info(m._net, i, C.byref(info)) == 0
        name = info.name.decode()
        H, W = hw if info.stride == 1 else (hw[0], hw[1] + 8)
        x = torch.randn(b, info.cin, H, W, generator=g).relu()
        if i < n:
            ref = bg_oracle.conv_layer(sd, name, x, info.ksize, info.stride)
        else:
            ref = F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"])
        y = torch.empty(ref.shape, device="cuda")
        rc = pf_lib.pf_bgnet_debug_conv(m._net, i, x.cuda().data_ptr(), b, H, W, y.data_ptr(), None)
        assert rc == 0, (name, pf_lib.pf_last_error())
        err = (y.cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
        assert err <= tol, (i, name, info.cin, info.cout, info.ksize, err)
        worst = max(worst, err)
    return worst


def test_split_storage_simt_kernels(pf_lib, bg_shapes, monkeypatch):
    """Split-bf16 storage with every conv forced onto the SIMT kernels (isolates storage from UMMA)."""
    monkeypatch.setenv("PF_TC_FORCE_SIMT", "1")
    layer_sweep(pf_lib, bg_shapes, 1e-4)


def test_every_conv_layer_tcgen05(pf_lib, bg_shapes, monkeypatch):
    monkeypatch.delenv("PF_TC_FORCE_SIMT", raising=False)
    worst = layer_sweep(pf_lib, bg_shapes, 1e-4)
    print("worst per-layer relative error (tcgen05 split-bf16):", worst)


def test_every_conv_layer_tcgen05_many_tiles(pf_lib, bg_shapes, monkeypatch):
    """Enough pixel tiles (> 148) that layers keep their full N per CTA: the n = 32 folded form, its 32-channel-chunk
    and N-split fallbacks, 48..128-cout unfolded layers with resident / streamed weights.  Odd sizes: ragged 8x14 tiles."""
    monkeypatch.delenv("PF_TC_FORCE_SIMT", raising=False)
    layer_sweep(pf_lib, bg_shapes, 1e-4, hw=(100, 170))


@pytest.mark.parametrize("env", [{"PF_HALO_EPI8": "0"}, {"PF_HALO_EPI16": "0"}, {"PF_HALO_EPI16": "1"}, {"PF_HALO_FOLD_TEAMS": "1"},
                                 {"PF_HALO_ALT": "7"}, {"PF_HALO_ALT": "0"}],
                         ids=lambda e: ",".join("%s=%s" % kv for kv in e.items()))
def test_every_conv_layer_tcgen05_epilogue_teams(pf_lib, bg_shapes, monkeypatch, env):
    """The epilogue-team variants of the halo kernel (conv_halo_kernel<224,0> only / two teams everywhere / four teams
    also for the fused conv1x1_up layers / two teams on the folded layers / two teams on alternate tiles for every layer
    with an N tile <= 32 / never): same results, enough tiles that every layer
    keeps its full N tile."""
    monkeypatch.delenv("PF_TC_FORCE_SIMT", raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    layer_sweep(pf_lib, bg_shapes, 1e-4, hw=(100, 170))


@pytest.mark.parametrize("cluster", ["1", "0"])
def test_every_conv_layer_tcgen05_cta_pairs(pf_lib, bg_shapes, monkeypatch, cluster):
    """Streamed-weight layers as CTA pairs (cluster 2x1x1, weight tiles loaded half by each CTA and multicast to both) and
    without.  One image of 100 x 162: 147 tiles of 16 x 8 -> 146 CTAs, so CTA 1's second round is a ghost tile (its pair
    partner still has a real one); the two-image sweeps above cover the even case."""
    monkeypatch.delenv("PF_TC_FORCE_SIMT", raising=False)
    monkeypatch.setenv("PF_HALO_CLUSTER", cluster)
    layer_sweep(pf_lib, bg_shapes, 1e-4, hw=(100, 162), b=1)


def test_every_conv_layer_tcgen05_unfolded(pf_lib, bg_shapes, monkeypatch):
    """The 3x3 layers on the one-MMA-pair-per-tap form of the halo kernel (the dx-folded form is the default)."""
    monkeypatch.delenv("PF_TC_FORCE_SIMT", raising=False)
    monkeypatch.setenv("PF_HALO_NO_FOLD", "1")
    layer_sweep(pf_lib, bg_shapes, 1e-4)


@pytest.mark.parametrize("shape,final", [((1, 3, 64, 128), None), ((2, 3, 128, 192), (256, 384)), ((1, 3, 256, 512), None)])
def test_whole_net_tcgen05(pf_lib, bg_shapes, shape, final):
    b, t, h, w = shape
    sd = synthetic.make_bg_state_dict(bg_shapes, seed=h)
    pc = synthetic.make_pc_inputs(b, 3, h, w, "R", seed=h)
    inp = {"seg": pc["seg"].long(), "depth": pc["depth"].clamp(0.1, 200), "depth_mask": pc["depth_mask"]}
    q = bg_oracle.predict(sd, inp, final)["orig_size_logits"]
    sd["model.finalConv.bias"] = sd["model.finalConv.bias"] - q.mean((0, 2, 3))
    ref = bg_oracle.predict(sd, inp, final)
    cu = {k: v.cuda() for k, v in inp.items()}
    out = gpu_model(sd, final, precision="tc").predict(cu, {})
    err, mism = check_against(out, ref, rel_tol=3e-4)
    print("tcgen05 path: logits rel err %.2e, label mismatches %d of %d" % (err, mism, ref["seg"].numel()))
    # against the fp32 SIMT path on the same device
    out32 = gpu_model(sd, final, precision="fp32").predict(cu, {})
    scale = out32["logits"].abs().max().item()
    assert (out["logits"] - out32["logits"]).abs().max().item() <= 3e-4 * scale


def test_every_conv_layer_tcgen05_per_tap_kernel(pf_lib, bg_shapes, monkeypatch):
    """PF_TC_NO_HALO=1: the one-TMA-box-per-tap tcgen05 kernel (conv_tc.cu) stays correct."""
    monkeypatch.delenv("PF_TC_FORCE_SIMT", raising=False)
    monkeypatch.setenv("PF_TC_NO_HALO", "1")
    layer_sweep(pf_lib, bg_shapes, 1e-4)


@pytest.mark.parametrize("gather", ["0", "1"])
@pytest.mark.parametrize("b,h,w", [(1, 128, 256), (3, 192, 320), (2, 256, 512)])
def test_fused_conv1x1_up_path(pf_lib, bg_shapes, monkeypatch, gather, b, h, w):
    """PF_TC_FUSE_UP=1: conv1x1_up commuted with the bilinear TransitionUp -- a low-resolution 1x1 over the upsampled
    slices into an fp32 buffer, then the high-resolution 1x1 over the skip slices adds its interpolation in the epilogue,
    from a low-resolution patch staged per tile by TMA (default) or by per-pixel global gathers (PF_TC_FUSE_UP_GATHER=1).
    No upsampled tensor is written."""
    monkeypatch.setenv("PF_TC_FUSE_UP", "1")
    monkeypatch.setenv("PF_TC_FUSE_UP_GATHER", gather)
    sd = synthetic.make_bg_state_dict(bg_shapes, seed=7)
    pc = synthetic.make_pc_inputs(b, 3, h, w, "R", seed=7)
    inp = {"seg": pc["seg"].long(), "depth": pc["depth"].clamp(0.1, 200), "depth_mask": pc["depth_mask"]}
    ref = bg_oracle.predict(sd, inp, None)
    out = gpu_model(sd, None, precision="tc").predict({k: v.cuda() for k, v in inp.items()}, {})
    scale = ref["logits"].abs().max().item()
    assert (out["logits"].cpu() - ref["logits"]).abs().max().item() <= 3e-4 * scale
    monkeypatch.setenv("PF_TC_FUSE_UP", "0")
    plain = gpu_model(sd, None, precision="tc").predict({k: v.cuda() for k, v in inp.items()}, {})
    assert (out["logits"] - plain["logits"]).abs().max().item() <= 3e-4 * scale


def test_label_only_fast_path_matches_interpolated_argmax(pf_lib, bg_shapes, monkeypatch):
    """return_logits=False: strips of the full-resolution map whose quarter-resolution source pixels agree on the class
    skip the interpolation (the head conv parks each source pixel's argmax in the padding channel).  Must equal the
    argmax of the interpolated logits (the return_logits=True path) and the hint-free kernel (PF_TC_NO_AMAX=1)."""
    h, w = 256, 512
    sd = synthetic.make_bg_state_dict(bg_shapes, seed=12)
    pc = synthetic.make_pc_inputs(2, 3, h, w, "R", seed=12)
    inp = {"seg": pc["seg"].long(), "depth": pc["depth"].clamp(0.1, 200), "depth_mask": pc["depth_mask"]}
    q = bg_oracle.predict(sd, inp, (2 * h, 2 * w))["orig_size_logits"]
    sd["model.finalConv.bias"] = sd["model.finalConv.bias"] - q.mean((0, 2, 3))       # class-balanced logits
    cu = {k: v.cuda() for k, v in inp.items()}
    full = gpu_model(sd, (2 * h, 2 * w), precision="tc", return_logits=True).predict(cu, {})
    fast = gpu_model(sd, (2 * h, 2 * w), precision="tc", return_logits=False, seg_dtype="uint8").predict(cu, {})["seg"]
    monkeypatch.setenv("PF_TC_NO_AMAX", "1")
    slow = gpu_model(sd, (2 * h, 2 * w), precision="tc", return_logits=False, seg_dtype="uint8").predict(cu, {})["seg"]
    assert len(torch.unique(fast)) >= 5                                               # a real multi-class map
    assert torch.equal(slow.long(), full["seg"].long())
    n_diff = int((fast.long() != full["seg"].long()).sum())
    assert n_diff <= 1e-5 * fast.numel(), n_diff


@pytest.mark.parametrize("precision,tol", [("tc", 3e-4), ("fp32", 1e-4)])
def test_poisoned_workspace(pf_lib, bg_shapes, precision, tol):
    """Every byte of the activation arena that a kernel reads must have been written by the same forward: the work space
    is filled with 0xFF (NaN patterns in bf16 and fp32) between two calls.  A NaN accumulator becomes 0 behind ReLU
    (fmaxf), so an uninitialised padding channel is a silently wrong result, not a NaN -- hence the comparison with the
    oracle, at a size where the layers keep their full N tiles (N = 24 single blocks, 2 x 24 splits, pooled epilogues)."""
    b, h, w = 2, 256, 512
    sd = synthetic.make_bg_state_dict(bg_shapes, seed=21)
    pc = synthetic.make_pc_inputs(b, 3, h, w, "R", seed=21)
    inp = {"seg": pc["seg"].long(), "depth": pc["depth"].clamp(0.1, 200), "depth_mask": pc["depth_mask"]}
    ref = bg_oracle.predict(sd, inp, None)
    m = gpu_model(sd, None, precision=precision)
    cu = {k: v.cuda() for k, v in inp.items()}
    first = m.predict(cu, {})
    m._ws.fill_(0xFF)
    out = m.predict(cu, {})
    assert torch.isfinite(out["logits"]).all()
    check_against(out, ref, rel_tol=tol)
    assert torch.equal(out["logits"], first["logits"]) or (out["logits"] - first["logits"]).abs().max().item() <= tol * ref["logits"].abs().max().item()
