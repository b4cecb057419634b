"""Summarise ncu outputs into profiles/: python tools/ncu_summarise.py <launches.csv> <prof.ncu-rep> <tag>"""
import collections, csv, subprocess, sys, io, json
launches, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
out = []
if launches != "-":
    rows = [r for r in csv.reader(open(launches, errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0].replace("void bnl::<unnamed>::", "").replace("void bnl::", "")
        try:
            t = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    out.append(f"## launch list ({launches}): {sum(a[0] for a in agg.values())} launches, {tot/1e6:.1f} ms under ncu (cold-cache, serialised)\n")
    out.append("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {c} | {t/1e6:.3f} | {100*t/tot:.2f}% | {t/c/1e3:.1f} |")
if rep != "-":
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__grid_size", "launch__block_size",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
    idx = {h: i for i, h in enumerate(hdr)}
    out.append(f"\n## ncu --set full ({rep})\n")
    for n, r in enumerate(rows[2:]):
        out.append(f"### launch {n}")
        for w in want:
            if w in idx:
                out.append(f"- {w} = {r[idx[w]]} {units[idx[w]]}")
open(f"profiles/{tag}.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
