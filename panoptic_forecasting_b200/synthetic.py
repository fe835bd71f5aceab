"""Seeded synthetic inputs for the bg-forecast hot path (no dataset, no network).

Shapes/dtypes follow what the reference datasets hand to the models:
  PCTransformModel.predict inputs  (pc_transform_model.py:27-32)
  BGModel.predict inputs           (bg_model.py:91-95, bg_dataset.py:223-261)
Camera/ego formulas follow data/data_utils.py:74-78 (extrinsics = vehicle_T_camera @ flu_T_rdf),
:117-165 (unicycle ego step, inverted) and :170-203 (yaw/pitch/roll rotation); they are
re-derived here because the reference helpers use the removed ``np.float``.
"""
import numpy as np
import torch

CITYSCAPES_K = (2262.52, 2265.3017905988554, 1096.98, 513.137)  # fx, fy, u0, v0


def intrinsics_mat(h=1024, w=2048):
    fx, fy, u0, v0 = CITYSCAPES_K
    sx, sy = w / 2048.0, h / 1024.0
    K = np.eye(3)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = fx * sx, fy * sy, u0 * sx, v0 * sy
    return K


def extrinsics_mat(yaw=0.0, pitch=0.038, roll=0.0, t=(1.7, 0.1, 1.22)):
    sy_, cy = np.sin(yaw), np.cos(yaw)
    sp, cp = np.sin(pitch), np.cos(pitch)
    sr, cr = np.sin(roll), np.cos(roll)
    R = np.array([[cy * cp, cy * sp * sr - sy_ * cr, cy * sp * cr + sy_ * sr],
                  [sy_ * cp, sy_ * sp * sr + cy * cr, sy_ * sp * cr - cy * sr],
                  [-sp, cp * sr, cp * cr]])
    V = np.eye(4)
    V[:3, :3] = R
    V[:3, 3] = t
    flu_T_rdf = np.eye(4)
    flu_T_rdf[:3, :3] = np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], dtype=np.float64)
    return V @ flu_T_rdf


def vehicle_now_T_prev(speed, yaw_rate, dt):
    if abs(yaw_rate) < 0.000175:
        x, y, th = dt * speed, 0.0, 0.0
    else:
        r = speed / yaw_rate
        wt = yaw_rate * dt
        x, y, th = r * np.sin(wt), r - r * np.cos(wt), wt
    T = np.eye(4)
    T[:3, :3] = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    T[:3, 3] = [x, y, 0]
    return np.linalg.inv(T)


def target_transforms(steps=(15, 12, 9), speed=10.0, yaw_rate=0.01, dt=1.0 / 17.0, rng=None):
    """target_T[i]: source-frame-i vehicle coords -> target-frame vehicle coords
    (pc_transform_dataset.py:165-186 accumulation)."""
    out = []
    for n in steps:
        T = np.eye(4)
        for k in range(n):
            s = speed if rng is None else speed + rng.normal(0, 0.3)
            yr = yaw_rate if rng is None else yaw_rate + rng.normal(0, 0.003)
            T = vehicle_now_T_prev(s, yr, dt) @ T
        out.append(T)
    return np.stack(out)


def _scene_realistic(rng, h, w, K, E):
    """Ground plane + fronto-parallel boxes + far wall; labels piecewise constant."""
    vs, us = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    fx, fy, u0, v0 = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    cam_h = 1.22
    ry = (vs - v0) / fy
    with np.errstate(divide="ignore", invalid="ignore"):
        ground = np.where(ry > 1e-3, cam_h / ry, np.inf)
    depth = np.minimum(ground, 120.0)
    seg = np.where(ground < 120.0, 0, 10).astype(np.uint8)  # road / sky-wall
    seg[(ground < 120.0) & (np.abs((us - u0) / fx * depth) > 4.0)] = 1  # sidewalk
    nbox = 20
    for _ in range(nbox):
        z = rng.uniform(6.0, 80.0)
        bw, bh = rng.uniform(1.0, 8.0), rng.uniform(1.5, 10.0)
        x0 = rng.uniform(-25.0, 25.0)
        uL, uR = (x0 * fx / z + u0), ((x0 + bw) * fx / z + u0)
        vB = cam_h * fy / z + v0
        vT = (cam_h - bh) * fy / z + v0
        m = (us >= uL) & (us < uR) & (vs >= vT) & (vs < vB) & (depth > z)
        depth[m] = z
        seg[m] = rng.integers(2, 19)
    depth = depth * (1.0 + rng.normal(0, 2e-3, size=depth.shape))
    return depth.astype(np.float32), seg


def _blob_mask(rng, h, w, frac):
    """~frac of pixels invalid, in rectangular blobs (moving-object masks)."""
    mask = np.ones((h, w), dtype=bool)
    target = frac * h * w
    covered = 0
    while covered < target:
        bh = int(rng.integers(max(2, h // 32), max(3, h // 6)))
        bw = int(rng.integers(max(2, w // 64), max(3, w // 8)))
        y0 = int(rng.integers(0, h - bh + 1))
        x0 = int(rng.integers(0, w - bw + 1))
        covered += mask[y0:y0 + bh, x0:x0 + bw].sum()
        mask[y0:y0 + bh, x0:x0 + bw] = False
    return mask


def make_pc_inputs(b=1, t=3, h=1024, w=2048, dist="R", seed=0, steps=(15, 12, 9), device="cpu"):
    """Inputs dict for PCTransformModel.predict. dist: 'R' realistic, 'U' adversarial iid."""
    rng = np.random.default_rng(seed)
    K = intrinsics_mat(h, w)
    E = extrinsics_mat()
    depth = np.empty((b, t, h, w), np.float32)
    seg = np.empty((b, t, h, w), np.uint8)
    mask = np.empty((b, t, h, w), bool)
    Ts = np.empty((b, t, 4, 4), np.float64)
    for bi in range(b):
        Ts[bi] = target_transforms(tuple(steps[i % len(steps)] for i in range(t)), rng=rng)
        for ti in range(t):
            if dist == "R":
                depth[bi, ti], seg[bi, ti] = _scene_realistic(rng, h, w, K, E)
                mask[bi, ti] = _blob_mask(rng, h, w, 0.15)
            else:
                depth[bi, ti] = rng.uniform(2.0, 82.0, size=(h, w)).astype(np.float32)
                seg[bi, ti] = rng.integers(0, 19, size=(h, w), dtype=np.uint8)
                mask[bi, ti] = rng.uniform(size=(h, w)) > 0.15
    dev = torch.device(device)
    return {
        "intrinsics": torch.from_numpy(np.broadcast_to(K, (b, 3, 3)).copy()).float().to(dev),
        "extrinsics": torch.from_numpy(np.broadcast_to(E, (b, 4, 4)).copy()).float().to(dev),
        "depth": torch.from_numpy(depth).to(dev),
        "depth_mask": torch.from_numpy(mask).to(dev),
        "target_T": torch.from_numpy(Ts).float().to(dev),
        "seg": torch.from_numpy(seg).to(dev),
    }


CITYSCAPES_BASELINE_M = 0.209313          # stereo baseline: depth = baseline * fx / disparity


def disparity_lut(fx=CITYSCAPES_K[0], baseline=CITYSCAPES_BASELINE_M):
    """65536-entry float32 table code -> depth for Cityscapes-style uint16 disparity PNGs: disparity =
    (code - 1) / 256 for code > 0, depth = baseline * fx / disparity (computed in float64, rounded once);
    code 0 (no measurement) and code 1 (zero disparity) map to depth 0."""
    code = np.arange(65536, dtype=np.float64)
    with np.errstate(divide="ignore"):
        depth = baseline * fx / ((code - 1.0) / 256.0)
    depth[:2] = 0.0
    return depth.astype(np.float32)


def pack_pc_inputs(inputs, lut=None):
    """Packed form of a PCTransformModel input dict (pf_zsplat_forward_frames_hop_packed): the float depth is
    replaced by the nearest uint16 disparity code of `lut` (so it becomes a value of the table, as real
    Cityscapes depth is) and the bool mask by 1 bit per pixel.  Returns (packed dict, unpacked dict whose
    'depth' is exactly lut[code] -- the same frames in the reference's formats)."""
    lut = disparity_lut() if lut is None else np.asarray(lut, np.float32)
    depth = inputs["depth"].cpu().numpy()
    fx, base = CITYSCAPES_K[0], CITYSCAPES_BASELINE_M
    with np.errstate(divide="ignore"):
        code = np.rint(base * fx / depth.astype(np.float64) * 256.0 + 1.0)
    code = np.clip(np.nan_to_num(code, nan=0.0, posinf=65535.0), 2, 65535).astype(np.uint16)
    mask = inputs["depth_mask"].cpu().numpy().astype(bool)
    b, t, h, w = depth.shape
    bits = np.packbits(mask.reshape(b, t, h * w), axis=-1, bitorder="little")
    packed = {k: v for k, v in inputs.items() if k not in ("depth", "depth_mask")}
    packed["depth_code"] = torch.from_numpy(code.view(np.int16))
    packed["depth_lut"] = torch.from_numpy(lut.copy())
    packed["depth_mask_bits"] = torch.from_numpy(bits)
    unpacked = dict(inputs)
    unpacked["depth"] = torch.from_numpy(lut[code])
    return packed, unpacked


def make_bg_inputs(b=1, t=3, h=512, w=1024, seed=0, device="cpu", label_dtype=torch.int64):
    """Inputs dict for BGModel.predict: labels iid {0..18},
    depth U(0,100), mask = depth > 5."""
    g = torch.Generator().manual_seed(seed)
    seg = torch.randint(0, 19, (b, t, h, w), generator=g, dtype=torch.int64).to(label_dtype)
    depth = torch.rand((b, t, h, w), generator=g) * 100.0
    mask = depth > 5.0
    dev = torch.device(device)
    return {"seg": seg.to(dev), "depth": depth.to(dev), "depth_mask": mask.to(dev)}


def make_bg_dense_inputs(b=1, t=3, h=64, w=128, seed=0, num_classes=11):
    """Inputs for BGModel with `convert2onehot` off (bg_model.py:61-65): soft per-class planes [b,t,C,h,w]
    (softmax of seeded noise), depth / mask as in make_bg_inputs."""
    base = make_bg_inputs(b, t, h, w, seed=seed)
    g = torch.Generator().manual_seed(seed + 1000)
    scores = torch.softmax(2.0 * torch.randn(b, t, num_classes, h, w, generator=g), dim=2)
    return {"seg": scores, "depth": base["depth"], "depth_mask": base["depth_mask"]}


def make_loss_target(pred_seg, seed=0, num_classes=11):
    """Label map for BGModel.loss: the predicted map with 30 % of the pixels re-drawn and a band of ignore (255)."""
    g = torch.Generator().manual_seed(seed + 2000)
    keep = torch.rand(pred_seg.shape, generator=g) < 0.7
    target = torch.where(keep, pred_seg.long(), torch.randint(0, num_classes, pred_seg.shape, generator=g))
    target[:, :5] = 255
    return target


def make_bg_state_dict(ref_state_dict_like, seed=0):
    """Seeded synthetic weights with randomised BatchNorm statistics (so folding is
    exercised). ``ref_state_dict_like``: a state_dict giving names and shapes."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(ref_state_dict_like.keys()):      # sorted: independent of module registration order
        v = ref_state_dict_like[k]
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros_like(v)
        elif k in ("depth_mean",):
            out[k] = torch.full_like(v, 30.0)
        elif k in ("depth_std",):
            out[k] = torch.full_like(v, 25.0)
        elif k.endswith("norm.weight") or k.endswith("running_var"):
            out[k] = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith("norm.bias") or k.endswith("running_mean"):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif k.endswith("conv.weight") or k.endswith("finalConv.weight"):
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            out[k] = torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif k.endswith("finalConv.bias"):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        else:
            raise KeyError(k)
    return out


def make_merge_inputs(b, n_per_item, h, w, seed=0, use_bbox_ulbr=True, mask_size=28):
    """Synthetic inputs of the fg -> bg panoptic merge (fg_model.py:515-588): per batch item a piecewise-constant
    background label map (ids 0..18, so the `>= 11 -> 255` rule fires), a background depth map with an invalid
    band, and `n` instances (blob-shaped mask logits, boxes that overlap each other and the image border,
    classes 0..7, depths straddling the background's).  numpy only; lists indexed by batch item."""
    rng = np.random.RandomState(seed)
    out = {k: [] for k in ("background", "bg_depth", "bg_depth_mask", "mask_logits", "bboxes", "classes", "depths")}
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    my, mx = np.mgrid[0:mask_size, 0:mask_size].astype(np.float32)
    for i in range(b):
        n = int(n_per_item[i])
        bg = ((yy // max(1, h // 6)).astype(np.int64) * 3 + (xx // max(1, w // 8)).astype(np.int64)) % 19
        out["background"].append(bg.astype(np.int64))
        depth = (8.0 + 60.0 * (1.0 - yy / h) + 5.0 * np.sin(xx / w * 12.0)).astype(np.float32)
        out["bg_depth"].append(depth)
        mask = np.ones((h, w), dtype=bool)
        mask[h // 3: h // 3 + max(1, h // 16), :] = False
        mask[:, w // 5: w // 5 + max(1, w // 32)] &= rng.rand(h, 1) > 0.5
        out["bg_depth_mask"].append(mask)
        cx = rng.uniform(0.0, w, n)
        cy = rng.uniform(0.2 * h, 0.9 * h, n)
        bw = rng.uniform(0.04 * w, 0.3 * w, n)
        bh = rng.uniform(0.08 * h, 0.5 * h, n)
        if n >= 2:                                   # force one heavy overlap
            cx[1], cy[1] = cx[0] + 0.2 * bw[0], cy[0] + 0.1 * bh[0]
        if use_bbox_ulbr:
            boxes = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
        else:
            boxes = np.stack([cx, cy, bw, bh], 1)
        out["bboxes"].append(boxes.astype(np.float32))
        logits = np.empty((n, mask_size, mask_size), dtype=np.float32)
        for k in range(n):
            r = rng.uniform(0.25, 0.55) * mask_size
            oy, ox = rng.uniform(0.35, 0.65, 2) * mask_size
            logits[k] = 6.0 * (1.0 - np.sqrt((my - oy) ** 2 + (mx - ox) ** 2) / r) + rng.normal(0, 0.7, (mask_size, mask_size))
        out["mask_logits"].append(logits)
        out["classes"].append(rng.randint(0, 8, n).astype(np.int64))
        out["depths"].append(rng.uniform(5.0, 70.0, n).astype(np.float32))
    return out
