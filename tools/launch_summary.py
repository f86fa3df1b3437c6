"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X) -> per-kernel totals.
usage: launch_summary.py launches.csv out.txt "command line that was profiled" """
import csv, re, sys
src, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = {}
for r in rows:
    name = re.sub(r"\(.*", "", r[ik])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv]) / 1e6
tot = sum(v[1] for v in agg.values())
with open(out, "w") as f:
    f.write(f"# {cmd}\n# ncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are cold-cache and serialised:\n"
            f"# compare SHARES, not absolutes)\n# total kernel time {tot:.1f} ms over {len(rows)} launches\n")
    f.write("launches  total_ms  share  kernel\n")
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{n:8d}  {ms:8.2f}  {100 * ms / tot:4.1f}%  {name}\n")
print(open(out).read())
