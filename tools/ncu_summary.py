"""Turn an .ncu-rep into the small text summary committed under profiles/.
usage: ncu_summary.py report.ncu-rep out.txt ["free-text note"]"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2:]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
with open(out, "w") as f:
    f.write(f"# {rep}\n# {note}\n")
    for v in vals:
        f.write(f"\nkernel: {v[hdr.index('Kernel Name')]}  grid {v[hdr.index('Grid Size')]} block {v[hdr.index('Block Size')]}\n")
        for k in KEYS:
            if k in hdr:
                f.write(f"  {k:70s} {v[hdr.index(k)]:>18s} {units[hdr.index(k)]}\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(src)))
    if len(r) > 3:
        # several kernels: keep the first kernel's block (blocks are separated by "Kernel Name" rows)
        ends = [i for i, row in enumerate(r) if row and row[0] == "Kernel Name"]
        stop = ends[1] if len(ends) > 1 else len(r)
        h, data = r[1], [row for row in r[2:stop] if len(row) == len(r[1])]
        ix = {x: i for i, x in enumerate(h)}
        tot = sum(int(x[ix["# Samples"]] or 0) for x in data) or 1
        stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
        agg = sorted(((sum(int(x[ix[s]] or 0) for x in data), s) for s in stalls), reverse=True)[:6]
        f.write("\nwarp-stall mix (% of samples): " + ", ".join(f"{s[6:]} {100*v/tot:.1f}" for v, s in agg) + "\n")
        f.write("top SASS lines by samples (%, executions, avg active threads, instruction):\n")
        for x in sorted(data, key=lambda x: -int(x[ix["# Samples"]] or 0))[:12]:
            f.write(f"  {100*int(x[ix['# Samples']] or 0)/tot:5.1f}%  {x[ix['Instructions Executed']]:>11}  {x[ix['Avg. Threads Executed']]:>3}  {x[ix['Source']][:80]}\n")
print("wrote", out)
