"""BASELINE config #5: the shape of the reference's train.py iteration (train.py:120-290) on a synthetic Waymo-dynamic scene:
random frame -> raytracing() -> depth / intensity / ray-drop / Chamfer losses -> backward -> Adam step on every Gaussian parameter.
Reports iterations/s on one GPU. Data loading, densification and logging are out of scope.
   python scripts/train_loop.py [--gaussians 2000000] [--actors 40] [--iters 60] [--frames 16]"""
import argparse, os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
import numpy as np, torch
from lidar_rt_b200 import synthetic as syn
from lidar_rt_b200.scene import GaussianAsset, LidarSensor
import lib.gaussian_renderer as gr
from lib.utils.chamfer3D.dist_chamfer_3D import chamfer_3DDist

ap = argparse.ArgumentParser()
ap.add_argument("--gaussians", type=int, default=2_000_000)
ap.add_argument("--actors", type=int, default=40)
ap.add_argument("--iters", type=int, default=60)
ap.add_argument("--warmup", type=int, default=8)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--no-fused-prepare", action="store_true")
ap.add_argument("--no-chamfer", action="store_true", help="leave out the Chamfer term of train.py:196-207")
ap.add_argument("--torch-adam", action="store_true", help="torch.optim.Adam(fused=True) over all tensors instead of lrt_adam_step")
ap.add_argument("--profile", action="store_true", help="print the torch profiler's top CUDA kernels of 10 iterations")
a = ap.parse_args()
dev = torch.device("cuda", 0)
P = a.gaussians + a.actors * 10000
sc = syn.make_street_scene(P, seed=1, n_actors=a.actors, per_actor=10000)
assets = []
for k in [-1] + list(range(a.actors)):
    m = sc.actor_id == k
    sub = syn.Scene(sc.means[m], sc.scales[m], sc.rots[m], sc.opac[m], sc.shs[m], sc.actor_id[m], 3)
    poses = None
    if k >= 0:          # actor Gaussians live in the actor's frame-0 pose; later frames = rigid translation (synthetic.actor_transform)
        poses = {f: (torch.tensor(syn.actor_transform(k, f)[1], device=dev), torch.tensor([[1.0, 0, 0, 0]], device=dev)) for f in range(a.frames)}
    assets.append(GaussianAsset(sub, device=dev, actor_poses=poses))
sensor = LidarSensor(device=dev)
for f in range(a.frames):
    sensor.add_frame(f, syn.sensor_pose(f))
H, W = sensor.H, sensor.W
bg = torch.tensor([0.0, 0.0, 1.0], device=dev)
args = types.SimpleNamespace(dynamic=True, pipe=types.SimpleNamespace(fused_prepare=not a.no_fused_prepare, compute_cov3D_python=False, convert_SHs_python=False),
                             opt=types.SimpleNamespace(use_rayhit=True))
# "ground truth": the initial render of every frame, perturbed
gt = {}
with torch.no_grad():
    for f in range(a.frames):
        pkg = gr.raytracing(f, assets, sensor, bg, args)
        gt[f] = (pkg["depth"] * (1 + 0.02 * torch.randn_like(pkg["depth"])), (pkg["intensity"] + 0.05 * torch.randn_like(pkg["intensity"])).clamp(0, 1),
                 (pkg["raydrop"] > 0.5).float())
groups = []
lrs = dict(_xyz=1.6e-4, _features_dc=2.5e-3, _features_rest=1.25e-4, _opacity=0.05, _scaling=5e-3, _rotation=1e-3)     # configs/exp.yaml
for name, lr in lrs.items():
    groups.append({"params": [getattr(x, name) for x in assets], "lr": lr})
if a.torch_adam:
    opt = torch.optim.Adam(groups, eps=1e-15, fused=True)
else:
    from lidar_rt_b200.optim import FusedAdam       # every tensor of every asset in one launch (lrt_adam_step)
    opt = FusedAdam(groups, eps=1e-15)
rng = np.random.default_rng(0)
chamLoss = chamfer_3DDist()

def iteration():
    f = int(rng.integers(0, a.frames))
    pkg = gr.raytracing(f, assets, sensor, bg, args)
    d_gt, i_gt, r_gt = gt[f]
    loss = (pkg["depth"] - d_gt).abs().mean() * 0.1 + (pkg["intensity"] - i_gt).abs().mean() + \
        torch.nn.functional.binary_cross_entropy(pkg["raydrop"].clamp(1e-6, 1 - 1e-6), r_gt) * 0.1
    if not a.no_chamfer:          # train.py:196-207 (lambda_cd = 0.01)
        mask = r_gt[..., 0] < 0.5 if r_gt.dim() == 3 else r_gt < 0.5
        gt_pts = sensor.inverse_projection_with_range(f, d_gt, mask)
        pred_pts = sensor.inverse_projection_with_range(f, pkg["depth"], mask)
        dist1, dist2, _, _ = chamLoss(pred_pts[None, ...], gt_pts[None, ...])
        loss = loss + 0.01 * (dist1 + dist2).mean() * 0.5
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss

for _ in range(a.warmup):
    iteration()
if a.profile:
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=False) as prof:
        for _ in range(10):
            iteration()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
    print(prof.key_averages(group_by_stack_n=0).table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=70))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(a.iters):
    loss = iteration()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"config #5 (synthetic): P = {P} Gaussians ({a.actors} actors), {H} x {W} rays, SH degree 3, fused_prepare={not a.no_fused_prepare}, chamfer={not a.no_chamfer}, adam={'torch fused' if a.torch_adam else 'lrt_adam_step'}: "
      f"{a.iters / dt:.1f} it/s ({1e3 * dt / a.iters:.2f} ms/it, {H * W * a.iters / dt / 1e6:.1f} Mrays/s), final loss {float(loss):.4f}")
