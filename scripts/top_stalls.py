"""Instructions with the most warp-stall samples in an `ncu -i X.ncu-rep --page source --csv` export (development aid).
   ncu -i prof.ncu-rep --page source --csv > src.csv; python scripts/top_stalls.py src.csv [N]"""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ia=hdr.index("Source"); isamp=hdr.index("# Samples"); iex=hdr.index("Instructions Executed"); ith=hdr.index("Avg. Threads Executed")
il=hdr.index("stall_long_sb"); iss=hdr.index("stall_short_sb")
tot=sum(int(r[isamp]) for r in data)
print("total samples",tot,"instr",len(data), "warp-inst", sum(int(r[iex]) for r in data))
top=sorted(range(len(data)), key=lambda i:-int(data[i][isamp]))[:int(sys.argv[2]) if len(sys.argv)>2 else 30]
for i in sorted(top):
    r=data[i]
    print(i, r[isamp], f"{100*int(r[isamp])/tot:.1f}%", "exec",r[iex],"thr",r[ith],"lsb",r[il],"ssb",r[iss], "|", r[ia].strip()[:90])
