"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list
(scripts/profile_frame.py --frames N) into per-kernel shares of a frame, and refresh profiles/traffic.json.
   python scripts/summarise_launches.py gpurun_out/r1_h_launches.csv 2 profiles/r1_h_launches_final"""
import csv, json, re, sys, collections, os
src, frames, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    name = r[ix["Kernel Name"]]
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name)
    if "cub" not in name and "at::" not in name:
        name = re.sub(r"<.*", "", re.sub(r"^void ", "", name)).strip()      # templates: k_bw_hits<(bool)0> -> k_bw_hits
    if "cub::" in name or name.startswith("void cub") or "DeviceRadixSort" in name or "DeviceScan" in name:
        name = "cub radix sort / scan kernels"
    elif name.startswith("void at::") or "at::native" in name:
        name = "torch (harness) kernels"
    d = per.setdefault(name, {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
    m, v = r[ix["Metric Name"]], float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    if m == "gpu__time_duration.sum":
        d["n"] += 1; d["ns"] += v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
    elif m == "dram__bytes_read.sum":
        d["rd"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    elif m == "dram__bytes_write.sum":
        d["wr"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
per = {k: v for k, v in per.items() if not k.startswith("torch")}
tot = sum(v["ns"] for v in per.values())
lines = [f"{'kernel':40s} {'launches/frame':>14s} {'share':>7s} {'ms/frame':>9s}   DRAM read + write per frame (MB)"]
for k, v in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
    lines.append(f"{k:40s} {v['n'] / frames:14.1f} {100 * v['ns'] / tot:6.1f}% {v['ns'] / frames / 1e6:9.3f}   {v['rd'] / frames / 1e6:8.1f} + {v['wr'] / frames / 1e6:8.1f}")
lines.append(f"{'total':40s} {sum(v['n'] for v in per.values()) / frames:14.1f} {100.0:6.1f}% {tot / frames / 1e6:9.3f}   "
             f"{sum(v['rd'] for v in per.values()) / frames / 1e6:8.1f} + {sum(v['wr'] for v in per.values()) / frames / 1e6:8.1f}")
open(out + ".txt", "a").write("\n".join(lines) + "\n")
print("\n".join(lines))
tj = {"kernels": {k: (v["rd"] + v["wr"]) / frames for k, v in per.items()},
      "step_total": sum(v["rd"] + v["wr"] for v in per.values()) / frames,
      "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum per frame: {os.path.basename(out)}.csv"}
json.dump(tj, open(os.path.join(os.path.dirname(out), "traffic.json"), "w"), indent=1)
