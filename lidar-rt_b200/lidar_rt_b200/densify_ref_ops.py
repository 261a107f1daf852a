"""The mask the reference computes for actors whose Gaussians leave their tracking box (gaussian_model.py:377-400, after Street
Gaussians): mask logic on the host with torch, like the reference; the row moves it triggers go through densify.prune_points."""
import torch


def _build_rotation(r):
    q = r / torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])[:, None]
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.zeros((q.size(0), 3, 3), device=r.device)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - w * z); R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y); R[:, 2, 1] = 2 * (y * z + w * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def points_outside_box(model, bb, repeat_num: int = 2):
    stds = model.get_scaling
    if getattr(model, "dimension", 2) == 2:
        stds = torch.cat([stds, 0 * torch.ones_like(stds[:, :1])], dim=-1)
    stds = stds[:, None, :].expand(-1, repeat_num, -1)
    xyz = model._xyz.detach()
    means = torch.zeros_like(xyz)[:, None, :].expand(-1, repeat_num, -1)
    samples = torch.normal(mean=means, std=stds)
    rots = _build_rotation(model._rotation.detach())[:, None, :, :].expand(-1, repeat_num, -1, -1)
    pts = torch.matmul(rots, samples.unsqueeze(-1)).squeeze(-1) + xyz[:, None, :].expand(-1, repeat_num, -1)
    n = xyz.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.bool, device=xyz.device)
    inside = torch.logical_and(torch.all((pts >= bb.min_xyz).view(n, -1), dim=-1), torch.all((pts <= bb.max_xyz).view(n, -1), dim=-1))
    return torch.logical_not(inside)
