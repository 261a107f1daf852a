"""Instructions with the most warp-stall samples in an `ncu -i X.ncu-rep --page source --csv [-k regex:kernel]` export (development aid).
   ncu -i prof.ncu-rep --page source --csv -k regex:k_sp_slots > src.csv; python scripts/top_stalls.py src.csv [N]"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[h]; data = []
for r in rows[h + 1:]:                       # first launch only (a report may hold several launches of the kernel)
    if r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        data.append(r)
ia = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed"); ith = hdr.index("Avg. Threads Executed")
il = hdr.index("stall_long_sb"); iss = hdr.index("stall_short_sb")
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "instr", len(data), "warp-inst", sum(int(r[iex]) for r in data))
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]
for i in sorted(top):
    r = data[i]
    print(i, r[isamp], f"{100 * int(r[isamp]) / tot:.1f}%", "exec", r[iex], "thr", r[ith], "lsb", r[il], "ssb", r[iss], "|", r[ia].strip()[:90])
