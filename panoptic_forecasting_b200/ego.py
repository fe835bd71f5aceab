"""Ego-motion transforms for the reprojection stage, built with torch (any device).

The reference builds `target_T` on the host per item: one unicycle-model step per recorded frame
(data/data_utils.py:117-165, `get_vehicle_now_T_prev`) accumulated over the frames between a source
frame and the target frame (data/datasets/pc_transform_dataset.py:165-186).  This module does the
same arithmetic batched, in float64, and returns float32 matrices for `pf_zsplat_forward*`.
"""
import torch

ANGLE_RAD_EPS = 0.000175   # data_utils.py:139: below this yaw rate the vehicle is treated as moving straight


def step_now_T_prev(speed, yaw_rate, dt):
    """[...]-shaped float64 tensors -> [...,4,4]: vehicle(now) <- vehicle(previous frame)."""
    speed, yaw_rate, dt = (torch.as_tensor(a, dtype=torch.float64) for a in (speed, yaw_rate, dt))
    straight = yaw_rate.abs() < ANGLE_RAD_EPS
    safe_rate = torch.where(straight, torch.ones_like(yaw_rate), yaw_rate)
    r = speed / safe_rate
    wt = yaw_rate * dt
    x = torch.where(straight, dt * speed, r * torch.sin(wt))
    y = torch.where(straight, torch.zeros_like(wt), r - r * torch.cos(wt))
    th = torch.where(straight, torch.zeros_like(wt), wt)
    c, s = torch.cos(th), torch.sin(th)
    # prev_T_now = [R(th) | (x, y, 0)]; the reference returns its inverse: [R^T | -R^T t]
    T = torch.zeros(speed.shape + (4, 4), dtype=torch.float64, device=speed.device)
    T[..., 0, 0], T[..., 0, 1] = c, s
    T[..., 1, 0], T[..., 1, 1] = -s, c
    T[..., 2, 2] = 1
    T[..., 3, 3] = 1
    T[..., 0, 3] = -(c * x + s * y)
    T[..., 1, 3] = -(-s * x + c * y)
    return T


def target_T_from_odometry(speed, yaw_rate, dt):
    """speed / yaw_rate / dt: [..., n] per-frame odometry between a source frame and the target frame
    (oldest first).  Returns float32 [...,4,4] = step_n @ ... @ step_1 (target <- source)."""
    steps = step_now_T_prev(speed, yaw_rate, dt)
    T = torch.eye(4, dtype=torch.float64, device=steps.device).expand(steps.shape[:-3] + (4, 4)).clone()
    for k in range(steps.shape[-3]):
        T = steps[..., k, :, :] @ T
    return T.float()
