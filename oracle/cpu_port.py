"""ORACLE / CPU BASELINE (test infrastructure, not product code): the reference's composite
bg-forecast path restated with the same multi-threaded torch CPU operators the reference calls,
so that its speed is representative of the reference's own CPU path:

  PCTransformModel.predict   pc_transform_model.py:26-150  (batched small-matrix `@`, scatter_min)
  disk hop                   export_cityscapes_segmentation_results.py:119-122, bg_dataset.py:223-230
  BGModel.predict            bg_model.py:91-102 (via oracle/bg_oracle.py)

torch_scatter.scatter_min (pytorch_scatter 2.0.5, un-vendored) is restated with
scatter_reduce_('amin') twice (value pass, then arg pass on lowest source index), i.e. its CPU
tie rule.  Results are checked against oracle/pc_transform_oracle.py in tests/test_oracle.py.

Used only by bench.py (`cpu_baseline` leg and `--impl reference`) and tests/.
"""
import torch

from . import bg_oracle


def pc_predict(inputs, only_this_ind=None):
    K, E = inputs["intrinsics"], inputs["extrinsics"]
    depth, mask, T, seg = inputs["depth"], inputs["depth_mask"], inputs["target_T"], inputs["seg"]
    if only_this_ind is not None:
        s = slice(only_this_ind, only_this_ind + 1)
        depth, mask, T, seg = depth[:, s], mask[:, s], T[:, s], seg[:, s]
    b, t, H, W = depth.shape
    N = H * W
    dev = depth.device          # CPU for the baseline; bench.py's library_baseline leg runs the same ops on the GPU
    vs, us = torch.meshgrid(torch.arange(H, dtype=torch.float, device=dev), torch.arange(W, dtype=torch.float, device=dev),
                            indexing="ij")
    pix = torch.stack([us.reshape(-1), vs.reshape(-1), torch.ones(N, device=dev)], -1).expand(b, N, 3)
    rays = (torch.inverse(K).reshape(b, 1, 3, 3) @ pix.unsqueeze(-1)).squeeze(-1)          # :51-54
    pc = rays.unsqueeze(1) * depth.reshape(b, t, N, 1)                                      # :55
    pc = torch.cat([pc, torch.ones(b, t, N, 1, device=dev)], -1).unsqueeze(-1)              # :56-59
    pv = E.view(b, 1, 1, 4, 4) @ pc                                                         # :63
    pt = T.unsqueeze(2) @ pv                                                                # :68
    qc = torch.inverse(E).reshape(b, 1, 1, 4, 4) @ pt                                       # :71
    qc = qc[:, :, :, :3] / qc[:, :, :, 3:4]                                                 # :72
    z = qc[:, :, :, 2].squeeze(-1)
    uv = K.view(b, 1, 1, 3, 3) @ qc                                                         # :74
    uv = (uv[:, :, :, :2] / uv[:, :, :, 2:3]).squeeze(-1)                                   # :75
    inb = (uv[..., 0] >= 0) & (uv[..., 0] < W) & (uv[..., 1] >= 0) & (uv[..., 1] < H)       # :83-86
    valid = (mask.view(b, t, N) * (z > 0) & inb).reshape(b, t * N)                          # :87-89
    z = z.reshape(b, t * N).clone()
    z[~valid] = z.max() + 1                                                                 # :105
    uv = uv.reshape(b, t * N, 2)
    fu, cu_ = uv[..., 0].floor().long(), uv[..., 0].ceil().long()
    fv, cv = uv[..., 1].floor().long(), uv[..., 1].ceil().long()
    xs = torch.cat([fu, fu, cu_, cu_], 1).clamp_(0, W - 1)                                  # :107-114
    ys = torch.cat([fv, cv, fv, cv], 1).clamp_(0, H - 1)
    zs = z.repeat(1, 4)
    cell = ys * W + xs                                                                      # :117
    Etot = 4 * t * N
    mn = torch.full((b, N), float("inf"), device=dev).scatter_reduce_(1, cell, zs, "amin", include_self=True)   # :118
    e = torch.arange(Etot, device=dev).expand(b, Etot)
    cand = torch.where(zs == mn.gather(1, cell), e, torch.full_like(e, Etot))
    arg = torch.full((b, N), Etot, dtype=torch.long, device=dev).scatter_reduce_(1, cell, cand, "amin", include_self=True)
    hit = arg < Etot                                                                        # :120
    src = (arg % (t * N)).clamp_(max=t * N - 1)
    segf = seg.reshape(b, t * N).clone()
    segf[~valid] = 0                                                                        # :133
    out_seg = torch.where(hit, segf.gather(1, src), torch.zeros_like(segf[:, :N]))          # :134
    out_depth = torch.where(hit, z.gather(1, src), torch.full((b, N), -1.0, device=dev))    # :136-139
    return {"seg": out_seg.view(b, H, W), "depth": out_depth.view(b, H, W)}


def disk_hop(depth, min_depth=0.1, max_depth=200.0):
    q = ((depth + 1).clamp(0, 255) * 256).round()
    q = q.to(torch.int32).to(torch.float32)            # uint16 PNG round trip
    d = q / 256.0 - 1
    m = d > 0
    d[~m] = -1
    d[m & (d > max_depth)] = max_depth
    d[m & (d < min_depth)] = min_depth
    return d, m


def composite_predict(sd, pc_inputs, final_size=None):
    """One target frame per batch item from t input frames: t single-frame reprojections ->
    disk hop -> BGModel.predict."""
    t = pc_inputs["depth"].shape[1]
    segs, deps, masks = [], [], []
    with torch.no_grad():
        for ind in range(t):
            r = pc_predict(pc_inputs, only_this_ind=ind)
            d, m = disk_hop(r["depth"])
            segs.append(r["seg"]); deps.append(d); masks.append(m)
        bg_in = {"seg": torch.stack(segs, 1).long(), "depth": torch.stack(deps, 1), "depth_mask": torch.stack(masks, 1)}
        out = bg_oracle.predict(sd, bg_in, final_size)
    out["warped_seg"], out["warped_depth"] = bg_in["seg"], bg_in["depth"]
    return out
