"""BASELINE config #5: the shape of the reference's train.py iteration (train.py:120-290) on a synthetic Waymo-dynamic scene:
random frame -> raytracing() -> depth / intensity / ray-drop / Chamfer losses -> backward -> Adam step on every Gaussian parameter.
Reports iterations/s. Data loading and logging are out of scope.
   python scripts/train_loop.py [--gaussians 2000000] [--actors 40] [--iters 60] [--frames 16] [--json OUT]
   torchrun --nproc-per-node N scripts/train_loop.py        DATA-PARALLEL training over frames: every rank renders and
                                                            back-propagates its own random frame of the same Gaussians, the gradients of all
                                                            246 tensors travel as ONE packed NCCL all-reduce (optim.all_reduce_gradients), every rank
                                                            takes the same optimiser step. Reports global iterations/s and frames/s.
The reference side of the comparison (its own tracer on OptiX, its Chamfer extension, per-asset torch Adam) is oracle/run_ref_train.py."""
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
ap.add_argument("--json", default=None, help="write the result as JSON (rank 0)")
a = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
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
rng = np.random.default_rng(rank)                 # every rank draws its own frames
chamLoss = chamfer_3DDist()
all_params = [p for g_ in groups for p in g_["params"]]
if world > 1:
    from lidar_rt_b200.optim import all_reduce_gradients

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
    if world > 1:
        all_reduce_gradients(all_params, world)   # one packed NCCL all-reduce (average); every rank then takes the same step
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
if world > 1:
    dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(a.iters):
    loss = iteration()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
if world > 1:
    t_ = torch.tensor([dt], device=dev); dist.all_reduce(t_, op=dist.ReduceOp.MAX); dt = float(t_.item())
    # the replicas must still agree after the run: same parameters on every rank
    chk = torch.stack([p.detach().double().sum() for p in all_params[:6]])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "data-parallel replicas diverged"
if rank == 0 and a.json:
    import json
    json.dump({"config": 5, "P": P, "actors": a.actors, "rays": H * W, "world": world, "iters": a.iters, "it_per_s": a.iters / dt, "ms_per_it": 1e3 * dt / a.iters,
               "frames_per_s": world * a.iters / dt, "fused_prepare": not a.no_fused_prepare, "chamfer": not a.no_chamfer,
               "adam": "torch fused" if a.torch_adam else "lrt_adam_step", "gradient_all_reduce": "one packed NCCL all-reduce per iteration" if world > 1 else None,
               "final_loss": float(loss), "gpu": torch.cuda.get_device_name(dev)}, open(a.json, "w"), indent=1)
if rank == 0:
  print(f"config #5 (synthetic): P = {P} Gaussians ({a.actors} actors), {H} x {W} rays, SH degree 3, fused_prepare={not a.no_fused_prepare}, chamfer={not a.no_chamfer}, adam={'torch fused' if a.torch_adam else 'lrt_adam_step'}: "
      f"world {world}: {a.iters / dt:.1f} it/s ({1e3 * dt / a.iters:.2f} ms/it, {world * a.iters / dt:.1f} frames/s, {world * H * W * a.iters / dt / 1e6:.1f} Mrays/s), final loss {float(loss):.4f}")
if world > 1:
    dist.destroy_process_group()
