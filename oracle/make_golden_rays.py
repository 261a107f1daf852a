#!/usr/bin/env python
"""oracle/make_golden_rays.py — TEST INFRASTRUCTURE. Generates tests/golden/ref_range2point.npz by executing the
reference's own LiDARSensor.range2point (lib/scene/lidar_sensor.py:325-393) unmodified on the CPU (function body
lifted with `ast`, device="cuda" redirected — see make_golden.py), for a Waymo-style inclination table and a
KITTI-style pair of bounds. Run in the build container (needs /root/reference)."""
import os, sys, types
import numpy as np, torch
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
from oracle.make_golden import REF, cuda_to_cpu, lift
from lidar_rt_b200 import synthetic as syn

ns = lift(f"{REF}/lib/scene/lidar_sensor.py", ["LiDARSensor.range2point"])
rng = np.random.default_rng(17)
out = {}
with cuda_to_cpu():
    for tag, H, W, inc, off, aoff in [("waymo", 8, 16, syn.waymo_inclinations(8).tolist(), 0.5, 0.0),
                                      ("kitti", 6, 12, [float(np.radians(-24.9)), float(np.radians(2.0))], 0.0, 0.1)]:
        pose = torch.from_numpy(syn.sensor_pose(7))
        me = types.SimpleNamespace(inclination_bounds=inc, sensor2world={0: pose}, sensor_center={0: pose[:3, 3]},
                                   H=H, W=W, pixel_offset=off, angle_offset=aoff)
        rmap = rng.uniform(0.5, 80.0, (H, W)).astype(np.float32)
        pts = ns["range2point"](me, 0, torch.from_numpy(rmap))
        out[f"{tag}_range"] = rmap; out[f"{tag}_points"] = pts.numpy().astype(np.float32); out[f"{tag}_pose"] = pose.numpy()
        out[f"{tag}_inc"] = np.asarray(inc, np.float32); out[f"{tag}_offsets"] = np.array([off, aoff], np.float32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_range2point.npz"), **out)
print({k: v.shape for k, v in out.items()})
