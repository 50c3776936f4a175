"""Summarise an ncu source-page CSV: instructions with the most stall samples, and totals by stall reason.
usage: ncu_src_top.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
data = rows[2:]
ia = hdr.index("Address"); isrc = hdr.index("Source"); ismp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for _, h in stall_cols}
recs = []
total_samples = 0; total_inst = 0
for k, r in enumerate(data):
    if len(r) < len(hdr) - 5: continue
    try: s = int(r[ismp])
    except: continue
    total_samples += s; total_inst += int(r[iex] or 0)
    st = {}
    for i, h in stall_cols:
        v = int(r[i] or 0); tot[h] += v
        if v: st[h.replace("stall_", "")] = v
    recs.append((s, k, r[isrc].strip(), int(r[iex] or 0), st))
print("total samples", total_samples, "warp-instructions", total_inst)
for h, v in sorted(tot.items(), key=lambda x: -x[1]):
    if v: print(f"  {h:28s} {v:8d} {100*v/total_samples:5.1f}%")
print("top instructions by samples:")
for s, k, src, ex, st in sorted(recs, reverse=True)[:top]:
    print(f"{s:6d} {100*s/total_samples:4.1f}% line {k:5d} exec {ex:9d}  {src[:60]:60s} {dict(sorted(st.items(), key=lambda x:-x[1])[:3])}")
