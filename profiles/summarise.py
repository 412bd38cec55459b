#!/usr/bin/env python
"""Turns the artefacts of one GPU run (bench JSON line, ncu launch list CSV, `ncu --set full` report) into the
markdown summary committed under profiles/.  Usage:
  python profiles/summarise.py TAG bench.json launches.csv full.ncu-rep [ref.json] > profiles/TAG.md
The .ncu-rep is read with `ncu -i ... --page raw --csv` (ncu is present in the build container, no GPU needed)."""
import csv, io, json, subprocess, sys, collections

tag, bench, launches, rep = sys.argv[1:5]
ref = sys.argv[5] if len(sys.argv) > 5 else None
d = json.loads(open(bench).read().strip().splitlines()[-1])
out = []
P = out.append
P("# %s" % tag)
P("")
P("## bench.py (one B200, %d steps, %d warm-up)" % (d["steps"], d["warmup"]))
P("")
c = d["config"]
P("Workload: %s; %d nodes, %d arcs, %.2f bits/arc, max outdegree %d." % (c["workload"], c["nodes"], c["arcs"], c["bits_per_arc"], c["max_outdegree"]))
P("")
P("| quantity | value |")
P("|---|---|")
P("| value (stream resident in HBM) | %.4g edges/s, %.3f ms/step |" % (d["value"], d["ms_per_step"]))
r = d["roofline"]
P("| kernels per step (CUDA events on the launching stream) | %s |" % ", ".join("%s %.3f ms" % (k, v) for k, v in sorted(r["step_kernels_ms"].items(), key=lambda kv: -kv[1])))
P("| roofline, dominant kernel %s | %.1f GB/s algorithmic = %.4f of %.1f GB/s (%s); share of step %.1f %% |" % (r["kernel"], r["achieved"], r["frac"], r["peak"], r["peak_source"], 100 * r["kernel_share_of_step"]))
P("| roofline, whole step | %.4f |" % r["whole_step_frac"])
e = d["e2e"]
P("| e2e (host buffers in, result out, every step) | %.4g edges/s; H2D %d B, D2H %d B per step; %d steps |" % (e["value"], e["h2d_bytes_per_step"], e["d2h_bytes_per_step"], e["steps"]))
if d.get("cpu_baseline"):
    cb = d["cpu_baseline"]
    P("| cpu_baseline (%s, %d core) | %.4g edges/s; %s |" % (cb["kind"], cb["cores"], cb["value"], cb["sample"]))
if ref:
    rr = json.loads(open(ref).read().strip().splitlines()[-1])
    P("| --impl reference (%s, %d cores) | %.4g edges/s |" % (rr["cpu_baseline"]["kind"], rr["cpu_baseline"]["cores"], rr["value"]))
P("| clocks | SM %s MHz (max %s), reasons %s |" % (d["clocks"]["sm_mhz"], d["clocks"]["sm_max_mhz"], d["clocks"]["reasons"]))
if d.get("extra"):
    for k, v in d["extra"].items():
        if isinstance(v, dict):
            P("| extra: %s | %s |" % (k, ", ".join("%s=%s" % (a, ("%.4g" % b) if isinstance(b, float) else b) for a, b in v.items() if a != "what")))
P("")
# launch list
rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for x in rows[2:]:
    if not x[iv]:
        continue
    name = x[ik].split("(")[0].replace("void ", "").replace("bvg::", "")
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(x[iv].replace(",", "")) / 1e6
tot = sum(a[1] for a in agg.values())
P("## ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`, cold-cache and serialised: shares, not absolutes)")
P("")
P("| kernel | launches | total ms | share |")
P("|---|---|---|---|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    P("| %s | %d | %.3f | %.1f %% |" % (k, n, ms, 100 * ms / tot))
P("")
step = collections.OrderedDict((k, v) for k, v in agg.items() if k.startswith(("k_scan_", "k_long_resid", "k_long_extras", "k_long_merge", "k_long_fold", "k_reduce_slots")))
st = sum(v[1] for v in step.values())
ev = r["step_kernels_ms"]; evt = sum(ev.values())
P("Shares inside a step (scan kernels only) against bench.py's CUDA-event shares of the same kernels:")
P("")
P("| kernel | ncu share of step | CUDA-event share of step |")
P("|---|---|---|")
for k, (n, ms_) in sorted(step.items(), key=lambda kv: -kv[1][1]):
    base = k.split("<")[0].replace("_lean", "")
    P("| %s | %.1f %% | %s |" % (k, 100 * ms_ / st, ("%.1f %%" % (100 * ev[base] / evt)) if base in ev else "-"))
P("")
# full report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hh, units = rr[0], rr[1]
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"), ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"), ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 sector hit %"), ("lts__t_sector_hit_rate.pct", "L2 sector hit %"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->L1 read"), ("l1tex__m_l1tex2xbar_write_bytes.sum", "L1->L2 write"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("launch__registers_per_thread", "registers / thread"),
        ("launch__grid_size", "grid"), ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard (warps/issue)"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"), ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe"), ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected")]
P("## ncu --set full (`--clock-control none --import-source on`), one launch each")
P("")
kn = hh.index("Kernel Name")
seen = []
cols = []
for x in rr[2:]:
    name = x[kn].split("(")[0].replace("void ", "")
    cols.append((name, x))
P("| metric | " + " | ".join(n for n, _ in cols) + " |")
P("|---|" + "---|" * len(cols))
for key, label in want:
    if key in hh:
        i = hh.index(key)
        P("| %s [%s] | " % (label, units[i]) + " | ".join(x[i] for _, x in cols) + " |")
# traffic per launch of the scan kernels (read by bench.py's roofline object)
if "dram__bytes_read.sum" in hh:
    ir, iw = hh.index("dram__bytes_read.sum"), hh.index("dram__bytes_write.sum")
    def to_bytes(v, u):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    best = {}
    for name, x in cols:
        key = "k_scan_extras" if "k_scan_extras" in name else ("k_scan_merge" if "k_scan_merge" in name else name.split("<")[0])
        tr = to_bytes(x[ir], units[ir]) + to_bytes(x[iw], units[iw])
        best[key] = max(best.get(key, 0), tr)  # the largest launch of a kernel (merge level 1)
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r01_traffic.json"), "w") as f:
        json.dump({"source": os.path.basename(rep) + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch)", "kernels": best}, f, indent=1)
print("\n".join(out))
