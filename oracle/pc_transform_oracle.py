"""ORACLE (test infrastructure, not product code): CPU restatement of
``PCTransformModel.predict`` -- unproject -> rigid chain -> reproject -> 4-way splat ->
nearest-depth (z-buffer) selection -> label/depth gather.

Follows /root/reference/panoptic_forecasting/models/pc_transform/pc_transform_model.py:26-150
step by step; every function cites the lines it restates.  The scatter itself
(``torch_scatter.scatter_min``, pytorch_scatter pinned at 2.0.5 by the reference README.md:23,
source not vendored in the reference tree) is restated from its published CPU algorithm:
sequential strict-``<`` update, i.e. ties go to the lowest source index; untouched cells keep
``arg == dim_size``.

Floating point: numpy float32, one rounding per multiply and per add, dot products summed
left to right starting from the first product (this reproduces ATen's small-matrix CPU
``bmm`` kernel, which is what ``@`` dispatches to for 3x3 / 4x4 operands).  The inverses of K
and E are taken as inputs (the reference calls ``torch.inverse``; callers pass that result).

Pinning: checked bit-for-bit against the unmodified reference run in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz; tests/test_oracle.py).  The reference
ships no tests or golden vectors of its own for this path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import numpy as np

F32 = np.float32


def _dot_rows(M, cols):
    """rows of M (k x k float32) times k column arrays -> list of k float32 arrays.
    Left-to-right accumulation, separate mul and add (pc_transform_model.py:54,63,68,71,74)."""
    out = []
    k = len(cols)
    for i in range(M.shape[0]):
        acc = M[i, 0] * cols[0]
        for j in range(1, k):
            acc = acc + M[i, j] * cols[j]
        out.append(acc)
    return out


def reproject_points(K, Kinv, E, Einv, T, depth, height, width):
    """One (batch, frame): returns (u', v', z') float32 arrays of length H*W.
    pc_transform_model.py:41-75."""
    K, Kinv, E, Einv, T = (np.asarray(a, F32) for a in (K, Kinv, E, Einv, T))
    vs, us = np.meshgrid(np.arange(height, dtype=F32), np.arange(width, dtype=F32), indexing="ij")
    us, vs = us.reshape(-1), vs.reshape(-1)
    one = np.ones_like(us)
    d = np.asarray(depth, F32).reshape(-1)
    # :54  K_inv @ [u, v, 1]
    r = _dot_rows(Kinv, [us, vs, one])
    # :55  * depth ; :56-59 homogeneous 1
    pc = [r[0] * d, r[1] * d, r[2] * d, one]
    pv = _dot_rows(E, pc)        # :63
    pt = _dot_rows(T, pv)        # :68
    qc = _dot_rows(Einv, pt)     # :71
    with np.errstate(divide="ignore", invalid="ignore"):
        x, y, z = qc[0] / qc[3], qc[1] / qc[3], qc[2] / qc[3]   # :72
        p = _dot_rows(K, [x, y, z])                              # :74
        u2, v2 = p[0] / p[2], p[1] / p[2]                        # :75
    return u2, v2, z


def predict(inputs, only_this_ind=None, is_img=False, lut=None):
    """Restates PCTransformModel.predict (pc_transform_model.py:26-150) for numpy inputs.

    inputs: intrinsics [b,3,3], extrinsics [b,4,4], depth [b,t,H,W] f32, depth_mask [b,t,H,W] bool,
            target_T [b,t,4,4], seg [b,t,H,W] (or [b,t,H,W,3] if is_img),
            optional intrinsics_inv / extrinsics_inv (default: numpy float32 LAPACK inverse).
    returns dict seg [b,H,W(,3)], depth [b,H,W] f32, result2d [b,t,H,W,2] int64.
    """
    K = np.asarray(inputs["intrinsics"], F32)
    E = np.asarray(inputs["extrinsics"], F32)
    Kinv = np.asarray(inputs["intrinsics_inv"], F32) if "intrinsics_inv" in inputs else np.linalg.inv(K).astype(F32)
    Einv = np.asarray(inputs["extrinsics_inv"], F32) if "extrinsics_inv" in inputs else np.linalg.inv(E).astype(F32)
    depth = np.asarray(inputs["depth"], F32)
    mask = np.asarray(inputs["depth_mask"]).astype(bool)
    T = np.asarray(inputs["target_T"], F32)
    seg = np.asarray(inputs["seg"])
    if only_this_ind is not None:                                   # :33-37
        s = slice(only_this_ind, only_this_ind + 1)
        depth, mask, T, seg = depth[:, s], mask[:, s], T[:, s], seg[:, s]
    b, t, H, W = depth.shape
    N = H * W
    u2 = np.empty((b, t * N), F32)
    v2 = np.empty((b, t * N), F32)
    z2 = np.empty((b, t * N), F32)
    for bi in range(b):
        for ti in range(t):
            u, v, z = reproject_points(K[bi], Kinv[bi], E[bi], Einv[bi], T[bi, ti], depth[bi, ti], H, W)
            sl = slice(ti * N, (ti + 1) * N)
            u2[bi, sl], v2[bi, sl], z2[bi, sl] = u, v, z
    with np.errstate(invalid="ignore"):
        inb = (u2 >= 0) & (u2 < W) & (v2 >= 0) & (v2 < H)            # :83-86
        valid = (mask.reshape(b, t * N) & (z2 > 0)) & inb            # :87-89
    sentinel = F32(z2.max() + F32(1))                               # :105 (whole tensor, all batches)
    z2 = np.where(valid, z2, sentinel)
    with np.errstate(invalid="ignore"):
        fu, cu = np.floor(u2).astype(np.int64), np.ceil(u2).astype(np.int64)   # :107-110
        fv, cv = np.floor(v2).astype(np.int64), np.ceil(v2).astype(np.int64)
    xs = np.concatenate([fu, fu, cu, cu], axis=1).clip(0, W - 1)     # :112-114 (replica order)
    ys = np.concatenate([fv, cv, fv, cv], axis=1).clip(0, H - 1)
    zs = np.tile(z2, (1, 4))                                         # :115
    cell = ys * W + xs                                               # :117
    Etot = 4 * t * N
    if is_img:
        out_seg = np.zeros((b, N, 3), seg.dtype)
        seg_flat = seg.reshape(b, t * N, 3)
    else:
        out_seg = np.zeros((b, N), seg.dtype)
        seg_flat = seg.reshape(b, t * N)
    out_depth = np.full((b, N), -1, F32)                             # :136-138
    for bi in range(b):
        # scatter_min (:118-119): min depth per cell, ties -> lowest source index e.
        order = np.lexsort((np.arange(Etot), zs[bi], cell[bi]))      # by cell, then depth, then e
        c_sorted = cell[bi][order]
        first = np.ones(Etot, bool)
        first[1:] = c_sorted[1:] != c_sorted[:-1]
        win_e = order[first]
        win_c = c_sorted[first]
        src = win_e % (t * N)
        ok = valid[bi][src]                                          # :133 zero labels of invalid points
        if is_img:
            out_seg[bi, win_c] = np.where(ok[:, None], seg_flat[bi][src], 0)
        else:
            lab = seg_flat[bi][src]
            if lut is not None:
                lab = np.asarray(lut, seg.dtype)[lab]
            out_seg[bi, win_c] = np.where(ok, lab, 0)                # :134
        out_depth[bi, win_c] = zs[bi][win_e]                         # :139
    res2d = np.stack([xs[:, :t * N], ys[:, :t * N]], axis=-1).reshape(b, t, H, W, 2)   # :147
    return {
        "seg": out_seg.reshape((b, H, W, 3) if is_img else (b, H, W)),
        "depth": out_depth.reshape(b, H, W),
        "result2d": res2d,
    }
