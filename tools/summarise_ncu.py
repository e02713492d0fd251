"""Summarise ncu outputs into the small text/JSON files committed under profiles/.

  summarise_ncu.py launches <launches.csv> <out.md>          per-kernel totals + shares of a launch list
  summarise_ncu.py full <report.ncu-rep> <out.md> [frames]   key metrics of one --set full capture
"""
import collections, csv, io, json, subprocess, sys

mode = sys.argv[1]
if mode == "launches":
    src, dst = sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(open(src, errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for r in rows:
        if r is hdr or len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        unit = r[hdr.index("Metric Unit")]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = r[ik].split("(")[0]
        tot[name] += v; cnt[name] += 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"| `{k[:90]}` | {cnt[k]} | {v:.3f} | {100 * v / total:.1f}% |\n")
        f.write(f"\ntotal device time in the captured window: {total:.3f} ms "
                "(ncu serialises launches and runs them cold-cache: compare shares, not absolutes)\n")
    print(open(dst).read())
else:
    rep, dst = sys.argv[2], sys.argv[3]
    frames = float(sys.argv[4]) if len(sys.argv) > 4 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h]
    with open(dst, "w") as f:
        for n, r in enumerate(rows[2:]):
            d = dict(zip(hdr, r))
            f.write(f"### launch {n}: `{d.get('Kernel Name', '')[:100]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in keys:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            if frames:
                tr = sum(float(d[k]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index(k)]]
                         for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                f.write(f"| DRAM traffic per frame | {tr / frames:.0f} | byte |\n")
                f.write(f"| warp instructions per frame | {float(d['smsp__inst_executed.sum']) / frames:.0f} | inst |\n")
            tot = sum(float(d[s] or 0) for s in stall)
            f.write("\nwarp stall samples: " + ", ".join(
                f"{s.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * float(d[s] or 0) / tot:.1f}%"
                for s in sorted(stall, key=lambda s: -float(d[s] or 0))[:8]) + "\n\n")
    print(open(dst).read())
