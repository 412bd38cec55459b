#!/usr/bin/env python
"""DRAM traffic of one scan step (and of the C4 batch) from an ncu metrics pass, written to profiles/r02_traffic.json, which
bench.py's roofline.traffic reads.  Run on the GPU box (one GPU; ncu serialises and replays every launch, so no number printed by
the profiled process is a bench value):

  python profiles/traffic.py capture gpurun_out        # runs ncu over `bench.py --profile-only scan|c4` for both workloads
  python profiles/traffic.py summarise gpurun_out      # here: reads the csv logs, writes profiles/r02_traffic.json

Per kernel family: mean dram__bytes_read.sum + dram__bytes_write.sum per launch x launches per step (k_reduce_slots runs once
per scan, so launches per step = launches of the family / launches of k_reduce_slots).
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
SCAN_FAMILIES = ("k_scan_extras", "k_scan_merge", "k_long_resid", "k_long_extras", "k_long_merge", "k_reduce_slots", "k_tile_scan",
                 "k_stream_extras", "k_stream_ivfix", "k_rel_offsets", "k_extras", "k_merge", "k_halo_copy")


def capture(outdir):
    for wl in ("powerlaw", "weblike"):
        for what, extra in (("scan", ["--steps", "2", "--warmup", "1"]), ("c4", ["--random-nodes", "10000000"])):
            log = os.path.join(outdir, "r02_ncu_%s_%s.csv" % (wl, what))
            cmd = ["ncu", "--metrics", METRICS, "--clock-control", "none", "--csv", "--log-file", log,
                   sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--profile-only", what] + extra
            print(" ".join(cmd), flush=True)
            subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def family(name):
    m = re.search(r"(k_[a-z0-9_]+)", name)
    if not m:
        return None
    f = m.group(1)
    for pre in ("k_scan_extras", "k_scan_merge"):  # k_scan_extras_lean -> k_scan_extras (the profile span's name)
        if f.startswith(pre):
            return pre
    return f


def read_log(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rd:
        if len(r) < len(hdr):
            continue
        key = (r[ix["ID"]], r[ix["Kernel Name"]])
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(unit, 1)
        per.setdefault(key, {})[r[ix["Metric Name"]]] = v * mult
    for (i, name), m in per.items():
        rows.append((int(i), name, m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0), m.get("gpu__time_duration.sum", 0.0)))
    rows.sort()
    return rows


def summarise(outdir):
    out = {}
    for wl in ("powerlaw", "weblike"):
        p = os.path.join(outdir, "r02_ncu_%s_scan.csv" % wl)
        if os.path.exists(p):
            rows = read_log(p)
            fam = {}
            for _, name, rd, wr, ms in rows:
                f = family(name)
                if f in SCAN_FAMILIES:
                    a = fam.setdefault(f, [0, 0.0, 0.0, 0.0])
                    a[0] += 1; a[1] += rd; a[2] += wr; a[3] += ms
            scans = max(1, fam.get("k_reduce_slots", [1])[0])
            kernels = {f: {"launches_per_step": a[0] / scans, "read_bytes": a[1] / scans, "write_bytes": a[2] / scans, "ncu_ms": a[3] / scans}
                       for f, a in fam.items() if f != "k_reduce_slots"}
            out.setdefault(wl, {})["1"] = {"step_bytes": sum(k["read_bytes"] + k["write_bytes"] for k in kernels.values()), "kernels": kernels,
                                           "scans_captured": scans, "how": "ncu --metrics %s over bench.py --profile-only scan" % METRICS}
        p = os.path.join(outdir, "r02_ncu_%s_c4.csv" % wl)
        if os.path.exists(p):
            rows = read_log(p)
            fam = {}
            for _, name, rd, wr, ms in rows:
                f = family(name)
                if f in ("k_random", "k_query_sizes", "k_gather_rows") or (f and f.startswith("k_scan_")) or f in ("k_long_resid", "k_long_extras", "k_long_merge"):
                    a = fam.setdefault(f, [0, 0.0, 0.0, 0.0])
                    a[0] += 1; a[1] += rd; a[2] += wr; a[3] += ms
            batches = max(1, fam.get("k_random", [1])[0])
            kernels = {f: {"launches_per_batch": a[0] / batches, "read_bytes": a[1] / batches, "write_bytes": a[2] / batches, "ncu_ms": a[3] / batches} for f, a in fam.items()}
            out.setdefault(wl + "_c4", {})["1"] = {"step_bytes": sum(k["read_bytes"] + k["write_bytes"] for k in kernels.values()), "kernels": kernels,
                                                  "batches_captured": batches}
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps({k: {n: v["step_bytes"] for n, v in d.items()} for k, d in out.items()}))


if __name__ == "__main__":
    (capture if sys.argv[1] == "capture" else summarise)(sys.argv[2] if len(sys.argv) > 2 else "gpurun_out")
