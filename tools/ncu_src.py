"""Summarise an ncu --page source --csv dump: top SASS lines by samples, with stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print("total samples", tot, "warp-instructions", inst)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("stall mix:", {k: round(100 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    s = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[k]] or 0), k) for k in stalls), reverse=True)[:2]
    print(f"{100*s/tot:5.1f}%  exec={r[ix['Instructions Executed']]:>10} thr={r[ix['Avg. Threads Executed']]:>5}  {r[ix['Source']][:90]:90s} {st}")
