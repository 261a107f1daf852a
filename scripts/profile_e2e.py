"""Where does the end-to-end step (raytracing() + loss.backward()) spend GPU time outside our kernels? (development aid)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from lidar_rt_b200 import synthetic as syn
from lidar_rt_b200.scene import GaussianAsset
import lib.gaussian_renderer as gr

dev = torch.device("cuda", 0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
sc = syn.make_street_scene(P, seed=1)
asset = GaussianAsset(sc, device=dev)
H, W = 64, 2650
bg = torch.tensor([0.0, 0.0, 1.0], device=dev)
frames = []
for f in range(6):
    o, d = syn.lidar_rays(H, W, syn.waymo_inclinations(), syn.sensor_pose(f))
    frames.append((torch.as_tensor(o, device=dev).reshape(3), torch.as_tensor(d, device=dev), torch.randn(H, W, 4, device=dev)))

def step(i):
    centre, rd, dL = frames[i]
    ro = centre[None, None].expand(H, W, 3)
    for p in asset.parameters():
        p.grad = None
    pkg = gr.raytracing(i, [asset], (ro, rd, centre), bg, None)
    loss = (pkg["intensity"] * dL[..., 0:1]).sum() + (pkg["depth"] * dL[..., 3:4]).sum() + (pkg["raydrop"] * dL[..., 2:3]).sum()
    loss.backward()

for i in range(3):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3, 6):
        step(i)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    ct = getattr(e, "device_time_total", None)
    if ct is None:
        ct = getattr(e, "cuda_time_total", 0)
    if ct > 0 and e.device_type.name == "CUDA" if hasattr(e, "device_type") else ct > 0:
        rows.append((ct / 3.0, e.count // 3, e.key[:90]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"GPU time per step: {tot/1e3:.3f} ms over {len(rows)} kernel kinds")
for ct, n, k in rows[:14]:
    print(f"{ct/1e3:8.3f} ms x{n:3d}  {k}")
cpu = sorted(((e.self_cpu_time_total / 3.0, e.count // 3, e.key[:80]) for e in prof.key_averages() if e.self_cpu_time_total > 0), reverse=True)
print(f"CPU self time per step: {sum(c[0] for c in cpu)/1e3:.3f} ms")
for ct, n, k in cpu[:22]:
    print(f"{ct/1e3:8.3f} ms x{n:3d}  {k}")
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for rep in range(4):
    for i in range(6):
        step(i)
torch.cuda.synchronize(); print(f"wall per step without profiler: {(time.perf_counter() - t0) / 24 * 1e3:.3f} ms")
t0 = time.perf_counter()
for rep in range(4):
    for i in range(6):
        centre, rd, dL = frames[i]
        ro = centre[None, None].expand(H, W, 3)
        with torch.no_grad():
            pkg = gr.raytracing(i, [asset], (ro, rd, centre), bg, None)
torch.cuda.synchronize(); print(f"wall per forward-only raytracing(): {(time.perf_counter() - t0) / 24 * 1e3:.3f} ms")
