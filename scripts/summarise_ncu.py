"""Text summary of an `ncu --set full` report (one block per kernel launch): the metrics DESIGN.md / BASELINE.md quote.
   python scripts/summarise_ncu.py gpurun_out/r1_h_full.ncu-rep > profiles/r1_h_top_kernels_ncu.txt"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct"]
stall = [h for h in hdr if "average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
for r in data:
    print("=" * 100)
    print(r[col["Kernel Name"]])
    for w in want:
        if w in col:
            print(f"  {w:70s} {r[col[w]]:>18s} {units[col[w]]}")
    top = sorted(((float(r[col[h]].replace(",", "")), h.split("stalled_")[1].split("_per")[0]) for h in stall if r[col[h]]), reverse=True)[:6]
    print("  warp stall reasons (warps per issue-active cycle):", ", ".join(f"{n} {v:.2f}" for v, n in top))
