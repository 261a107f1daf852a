"""Drop-in for the reference render glue `lib.gaussian_renderer`
(/root/reference/lib/gaussian_renderer/__init__.py:15-181): same `raytracing()` signature and
returned dict; `render` is provided as an alias (BASELINE.json names it).

Differences underneath: the proxy mesh (build2DRectangle, :142) is never materialised — the LBVH is
built straight from the concatenated Gaussian parameters — and `pipe.compute_cov3D_python` /
`pipe.convert_SHs_python` / `override_color` (paths the reference's device code does not implement
either) raise instead of silently rendering garbage.

Assets are duck-typed like the reference's GaussianModel (lib/scene/gaussian_model.py:112-148):
`get_world_xyz(frame)`, `get_opacity`, `get_scaling`, `get_rotation(frame) -> (obj_quat, local_quat)`,
`get_features`, `active_sh_degree`. Sensors: anything with `get_range_rays(frame)` and
`sensor_center[frame]` (lib/scene/lidar_sensor.py:395-434), a pinhole camera shaped like lib/scene/cameras.py::Camera
(`camera_center`, `image_width`, `image_height`, `FoVx`, `world_view_transform`; reference :31-41), or a tuple
(rays_o, rays_d, centre).
"""
import math

import torch
import torch.nn.functional as F

from diff_lidar_tracer import Tracer, TracingSettings
from lidar_rt_b200.prepare import fused_prepare

tracer_2dgs = Tracer()              # module-global, constructed at import like the reference (:11)


def quaternion_raw_multiply(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Hamilton product, real part first (lib/utils/general_utils.py:156-174)."""
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz,
                        aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw), -1)


def get_rays(K, c2w):
    """Pinhole rays of the Camera branch (lib/utils/graphics_utils.py:88-95, called at lib/gaussian_renderer/__init__.py:41):
    one ray per pixel through K, rotated into the world by c2w (3,4); directions are NOT normalised (z = 1 in the camera
    frame), exactly as the reference hands them to its tracer. rays_o is the camera centre as a stride-0 view (the reference
    materialises the expansion; the values are the same and the tracer takes the shared-origin path)."""
    W, H = int(K[0][2] * 2), int(K[1][2] * 2)
    dev = c2w.device
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(H - 1, 0, H), indexing="ij")
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], torch.ones_like(i)], -1).to(dev)
    rays_d = dirs @ c2w.T[:3, :3]
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def _is_camera(sensor) -> bool:
    return all(hasattr(sensor, a) for a in ("camera_center", "image_width", "image_height", "FoVx", "world_view_transform"))


def _arg(args, group, name, default):
    g = getattr(args, group, None) if args is not None else None
    return getattr(g, name, default) if g is not None else default


def _assemble(frame, gaussian_assets, dynamic, decomp):
    """The reference's accessor loop and concatenations (:76-134), op for op."""
    all_means, all_opac, all_scales, all_shs, obj_rot, rot_local = [], [], [], [], [], []
    for pc in gaussian_assets:
        all_means.append(pc.get_world_xyz(frame))
        all_opac.append(pc.get_opacity)
        all_scales.append(pc.get_scaling)
        r1, r2 = pc.get_rotation(frame)
        obj_rot.append(r1.expand(r2.shape[0], -1))
        rot_local.append(r2)
        all_shs.append(pc.get_features)
    _cat = lambda xs: xs[0] if len(xs) == 1 else torch.cat(xs, 0)      # one asset: no 232 B/Gaussian copy (+ its backward)
    means3D = _cat(all_means)
    opacity = _cat(all_opac)
    scales = _cat(all_scales)
    shs = _cat(all_shs)
    if decomp == "background" or not dynamic:                      # reference :117-130
        rotations = rot_local[0] if len(rot_local) == 1 else torch.cat(rot_local, 0)
    elif decomp == "object":
        rotations = quaternion_raw_multiply(torch.cat(obj_rot, 0), F.normalize(torch.cat(rot_local, 0), dim=1))
    else:
        rot_act = quaternion_raw_multiply(torch.cat(obj_rot[1:], 0), F.normalize(torch.cat(rot_local[1:], 0), dim=1)) \
            if len(rot_local) > 1 else rot_local[0][:0]
        rotations = torch.cat([rot_local[0], rot_act], 0)

    return means3D, opacity, scales, rotations, shs


def raytracing(frame, gaussian_assets, sensor, background, args, scaling_modifier=1.0, override_color=None,
               decomp=False):
    if decomp == "background":
        gaussian_assets = gaussian_assets[:1]
    elif decomp == "object":
        gaussian_assets = gaussian_assets[1:]

    if isinstance(sensor, tuple):
        rays_o, rays_d, sensor_center = sensor[0], sensor[1], sensor[2]
    elif _is_camera(sensor):                                               # reference :31-41
        sensor_center = sensor.camera_center
        focal = 0.5 * sensor.image_width / math.tan(0.5 * sensor.FoVx)
        K = [[focal, 0, 0.5 * sensor.image_width], [0, focal, 0.5 * sensor.image_height], [0, 0, 1]]
        rays_o, rays_d = get_rays(K, sensor.world_view_transform.T.inverse()[:3, :4])
    elif hasattr(sensor, "get_range_rays"):
        rays_o, rays_d = sensor.get_range_rays(frame)
        sensor_center = sensor.sensor_center[frame]
    else:
        raise ValueError("sensor type not supported")

    if override_color is not None or _arg(args, "pipe", "convert_SHs_python", False):
        raise NotImplementedError("precomputed colours are not implemented by the tracer (nor by the reference's device code)")
    if _arg(args, "pipe", "compute_cov3D_python", False):
        raise NotImplementedError("precomputed covariances are not implemented by the tracer")

    dev = rays_d.device
    tracer_settings = TracingSettings(
        image_height=None, image_width=None, tanfovx=None, tanfovy=None,
        bg=background.to(dev), scale_modifier=1.0,
        viewmatrix=torch.empty(0, device=dev), projmatrix=torch.empty(0, device=dev),
        sh_degree=gaussian_assets[0].active_sh_degree, campos=sensor_center.to(dev),
        prefiltered=False, debug=False)

    dynamic = bool(getattr(args, "dynamic", False)) if args is not None else False
    fused = None
    if _arg(args, "pipe", "fused_prepare", True):
        # one native call instead of the accessor loop + torch.cat below (same values, same gradients); applies when
        # the assets expose the GaussianModel leaves
        # SH coefficients are not concatenated either: the tracer reads features_dc / features_rest in place and writes their
        # gradients in place (2 x 0.92 GB of copies per training step at 2.4 M Gaussians otherwise); pipe.sh_in_place=False restores the copy
        fused = fused_prepare(gaussian_assets, frame, dynamic, decomp, tracer_2dgs.optix_context.ctx,
                              sh_in_place=bool(_arg(args, "pipe", "sh_in_place", True)))
    if fused is not None:
        means3D, opacity, scales, rotations, shs = fused
    else:
        means3D, opacity, scales, rotations, shs = _assemble(frame, gaussian_assets, dynamic, decomp)

    grads3D = torch.zeros_like(means3D, requires_grad=True)
    try:
        means3D.retain_grad()
    except Exception:
        pass

    tracer = tracer_2dgs
    tracer.build_acceleration_structure(None, None, rebuild=True)      # reference :145 (rebuild every call)
    rendered, accum_w = tracer(ray_o=rays_o, ray_d=rays_d, mesh_normals=None, means3D=means3D, grads3D=grads3D,
                               shs=shs, colors_precomp=None, opacities=opacity, scales=scales, rotations=rotations,
                               cov3Ds_precomp=None, tracer_settings=tracer_settings)

    intensities = rendered[..., 0:1]
    rayhit_logits = rendered[..., 1:2]
    raydrop_logits = rendered[..., 2:3]
    depth = rendered[..., 3:4]
    # the reference reads args.opt.use_rayhit unconditionally (:168); with args=None (this repository's benches) its
    # configured value applies (configs/exp.yaml:44: True)
    if args.opt.use_rayhit if (args is not None and hasattr(args, "opt")) else True:
        prob = F.softmax(torch.cat([rayhit_logits, raydrop_logits], dim=-1), dim=-1)
        raydrop_prob = prob[..., 1:2]
    else:
        raydrop_prob = torch.sigmoid(raydrop_logits)
    return {"depth": depth, "intensity": intensities, "raydrop": raydrop_prob, "means3D": means3D,
            "accum_gaussian_weight": accum_w.unsqueeze(-1)}


render = raytracing
