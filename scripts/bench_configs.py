"""The other BASELINE.json configs, timed on one GPU with CUDA events (bench.py measures the headline config):

  #2 Waymo static single frame   P = 1 M, 64 x 2650 rays, rebuild + forward + backward
  #3 KITTI-360 dynamic           P = 0.5 M + 20 actors x 10 k, 66 x 1030 rays, 50 frames, actors move every frame:
                                 acceleration structure REFIT per frame (lrt_refit; a full rebuild every `--rebuild-every`
                                 frames), forward + backward; the refit frame is checked against a rebuild of the same frame
  #4 Waymo dynamic sweep         P = 2 M + 40 actors x 10 k, this rank's share of a 200-frame sweep (25 frames on 8 GPUs),
                                 rebuild + forward per frame through lidar_rt_b200.sweep.render_sweep (+ gather when run
                                 under torchrun)

   python scripts/bench_configs.py [--configs 2,3,4] [--out gpurun_out/configs.json]
   torchrun --nproc-per-node N scripts/bench_configs.py --configs 4
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
import numpy as np, torch
from lidar_rt_b200 import native, synthetic as syn
from lidar_rt_b200.sweep import frames_of_rank, render_sweep

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="2,3,4")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
ap.add_argument("--rebuild-every", type=int, default=10)
ap.add_argument("--sweep-frames", type=int, default=200)
ap.add_argument("--sweep-world", type=int, default=8, help="config #4: world size the single-process run stands for")
a = ap.parse_args()
want = {int(x) for x in a.configs.split(",")}
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
BG = cu(np.array([0, 0, 1], np.float32))
ev = lambda: torch.cuda.Event(enable_timing=True)
res = {"gpu": torch.cuda.get_device_name(dev), "world": world}


def upload(sc):
    return tuple(map(cu, (sc.means, sc.scales, sc.rots, sc.opac, sc.shs)))


if 2 in want and rank == 0:
    sc = syn.make_street_scene(1_000_000, seed=2)
    g = upload(sc)
    ctx = native.Context(dev)
    rng = np.random.default_rng(0)
    ts = []
    for f in range(13):
        o, d = syn.lidar_rays(64, 2650, syn.waymo_inclinations(), syn.sensor_pose(f))
        dL = np.zeros((64, 2650, 9), np.float32); dL[..., :4] = rng.standard_normal((64, 2650, 4))
        ro, rd, gl = cu(o), cu(d), cu(dL)
        torch.cuda.synchronize()
        e = [ev() for _ in range(4)]
        e[0].record(); ctx.build(*g[:4])
        e[1].record(); r = ctx.forward(ro, rd, BG, *g, 3)
        e[2].record(); ctx.backward(ro, rd, BG, *g, 3, r["out"], gl, hits=r)
        e[3].record(); torch.cuda.synchronize()
        if f >= 3:
            ts.append([e[i].elapsed_time(e[i + 1]) for i in range(3)])
    t = np.median(np.array(ts), 0)
    R = 64 * 2650
    res["config2_waymo_static_1M"] = {"P": sc.P, "rays": R, "ms": {"build": t[0], "forward": t[1], "backward": t[2], "step": float(t.sum())},
                                      "mrays_per_s_fwd_bwd": R / t.sum() / 1e3, "contributing_hits_per_ray": float(r["hit_cnt"].float().mean())}
    print(json.dumps(res["config2_waymo_static_1M"]), flush=True)
    ctx.close(); del g

if 3 in want and rank == 0:
    H, W = syn.KITTI_H, 1030
    sc0 = syn.make_street_scene(500_000 + 20 * 10_000, seed=3, n_actors=20, per_actor=10_000)
    ctx = native.Context(dev); chk = native.Context(dev)
    inc = syn.kitti_inclinations(H)
    rng = np.random.default_rng(1)
    ts, worst = [], 0.0
    shs = cu(sc0.shs); scales, rots, opac = cu(sc0.scales), cu(sc0.rots), cu(sc0.opac)
    n_frames = 50
    for f in range(n_frames):
        sc = syn.scene_at_frame(sc0, f)
        means = cu(sc.means)
        o, d = syn.lidar_rays(H, W, inc, syn.sensor_pose(f), pixel_offset=0.0)
        dL = np.zeros((H, W, 9), np.float32); dL[..., :4] = rng.standard_normal((H, W, 4))
        ro, rd, gl = cu(o), cu(d), cu(dL)
        refit = f % a.rebuild_every != 0
        torch.cuda.synchronize()                 # the frame's uploads are not part of the step
        e = [ev() for _ in range(4)]
        import time as _t
        h = [_t.perf_counter()]
        na = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        e[0].record(); ctx.build(means, scales, rots, opac, refit=refit); h.append(_t.perf_counter())
        e[1].record(); r = ctx.forward(ro, rd, BG, means, scales, rots, opac, shs, 3); h.append(_t.perf_counter())
        e[2].record(); gr = ctx.backward(ro, rd, BG, means, scales, rots, opac, shs, 3, r["out"], gl, hits=r); h.append(_t.perf_counter())
        e[3].record(); torch.cuda.synchronize()
        if os.environ.get("LRT_VERBOSE") == "2":          # which kernel makes a frame slow: re-run its forward with per-kernel timing
            fw = e[1].elapsed_time(e[2])
            if fw > 4.0 or f in (5, 20, 35):
                import ctypes as _ct
                ctx.set_option(native.OPT_KERNEL_TIMING, 1); ctx.kernel_times()
                r3 = ctx.forward(ro, rd, BG, means, scales, rots, opac, shs, 3, want_slots=True); torch.cuda.synchronize()
                kt = {k: round(v[0], 2) for k, v in ctx.kernel_times().items() if v[0] > 0.05}
                ctx.set_option(native.OPT_KERNEL_TIMING, 0)
                c = (_ct.c_int * 16)(); ctx.lib.lrt_debug_counters(ctx._h, c)
                sl = r3["slot_cnt"]; hcn = r3["hit_cnt"]
                print("cfg3 frame", f, "fwd ms", round(fw, 2), kt, "fallback rays", c[8], "heavy items", c[9], "slots/ray mean", round(float(sl.float().mean()), 1),
                      "max", int(sl.max()), ">256:", int((sl > 256).sum()), ">512:", int((sl > 512).sum()), "hits max", int(hcn.max()), flush=True)
                del r3
        if os.environ.get("LRT_VERBOSE") == "1" and f < 14:
            print("cfg3 frame", f, "refit" if refit else "build", "device ms", [round(e[i].elapsed_time(e[i + 1]), 2) for i in range(3)],
                  "host ms", [round(1e3 * (h[i + 1] - h[i]), 2) for i in range(3)], "cudaMalloc", torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - na, flush=True)
        if f >= 2:
            ts.append([refit] + [e[i].elapsed_time(e[i + 1]) for i in range(3)])
        if refit and f % 7 == 3:        # the refit structure must give what a fresh build of the same frame gives
            chk.build(means, scales, rots, opac)
            r2 = chk.forward(ro, rd, BG, means, scales, rots, opac, shs, 3)
            assert torch.equal(r["out"], r2["out"]) and torch.equal(r["hit_cnt"], r2["hit_cnt"]), f"frame {f}: refit differs from rebuild"
            worst = max(worst, float((r["out"] - r2["out"]).abs().max()))
    ts = np.array(ts, np.float64)
    rf, rb = ts[ts[:, 0] == 1], ts[ts[:, 0] == 0]
    R = H * W
    step = float(np.mean(ts[:, 1:].sum(1)))
    res["config3_kitti_dynamic_refit"] = {"P": sc0.P, "actors": 20, "rays": R, "frames": n_frames, "rebuild_every": a.rebuild_every,
                                          "ms": {"refit": float(np.median(rf[:, 1])), "rebuild": float(np.median(rb[:, 1])) if len(rb) else None,
                                                 "forward": float(np.median(ts[:, 2])), "backward": float(np.median(ts[:, 3])), "step_mean": step},
                                          "mrays_per_s_fwd_bwd": R / step / 1e3, "refit_equals_rebuild_bitwise": True,
                                          "contributing_hits_per_ray": float(r["hit_cnt"].float().mean())}
    print(json.dumps(res["config3_kitti_dynamic_refit"]), flush=True)
    ctx.close(); chk.close()

if 4 in want:
    sc0 = syn.make_street_scene(2_000_000 + 40 * 10_000, seed=4, n_actors=40, per_actor=10_000)
    ctx = native.Context(dev)
    shs = cu(sc0.shs); scales, rots, opac = cu(sc0.scales), cu(sc0.rots), cu(sc0.opac)
    means0 = cu(sc0.means); aid = cu(sc0.actor_id.astype(np.int64))
    inc = syn.waymo_inclinations()
    eff_world = world if world > 1 else a.sweep_world
    mine = frames_of_rank(a.sweep_frames, rank if world > 1 else 0, eff_world)
    # host-side inputs of this rank's frames prepared up front (data loading is outside the path)
    rays, shift = {}, {}
    for f in mine:
        o, d = syn.lidar_rays(64, 2650, inc, syn.sensor_pose(f))
        rays[f] = (cu(o), cu(d))
        t = np.zeros((41, 3), np.float32)
        for k in range(40):
            t[k] = syn.actor_transform(k, f)[1]
        shift[f] = cu(t)
    bg = BG

    def render(f):
        means = means0 + shift[f][aid]          # actors translate rigidly; background rows index the zero row (-1 -> last)
        ctx.build(means, scales, rots, opac)
        return ctx.forward(rays[f][0], rays[f][1], bg, means, scales, rots, opac, shs, 3, record_hits=False)["out"]

    for f in mine[:2]:
        render(f)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = ev(), ev()
    e0.record()
    out = render_sweep(render, a.sweep_frames, rank if world > 1 else 0, eff_world, gather=world > 1)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    if rank == 0:
        R = 64 * 2650
        n_done = a.sweep_frames if world > 1 else len(mine)
        res["config4_waymo_dynamic_sweep"] = {"P": sc0.P, "actors": 40, "rays_per_frame": R, "sweep_frames": a.sweep_frames, "world": eff_world,
                                              "measured_ranks": world, "frames_rendered_in_timed_region": n_done, "ms_sweep_share": ms,
                                              "ms_per_frame_per_gpu": ms / len(mine), "forward_only": True, "gathered": world > 1,
                                              "mrays_per_s": n_done * R / ms / 1e3}
        print(json.dumps(res["config4_waymo_dynamic_sweep"]), flush=True)
    ctx.close()

if rank == 0:
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
