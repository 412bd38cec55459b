"""Where a 1/N shard's scan step spends its time (strong scaling, BASELINE config C5), measured on ONE GPU: every shard of the
N-way plan opened in turn, scanned with CUDA events (warm, 30 repetitions) and once more with the per-kernel profile on.
Usage: python profiles/shard_profile.py [N [first-k-shards]] > profiles/r02_shards.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from webgraph_b200 import bvgraph  # noqa: E402


def main():
    nsh = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    only = int(sys.argv[2]) if len(sys.argv) > 2 else nsh   # profile the first `only` shards
    sys.argv = sys.argv[:1]
    args = bench.parse_args()
    base, meta = bench.graph_files(args, "powerlaw", 0, lambda: None)
    bounds = bvgraph.plan_shards(base, nsh)
    out = {"shards": nsh, "bounds": [int(b) for b in bounds], "rows": []}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for r in range(only):
        g = bvgraph.BVGraph.loadShard(base, int(bounds[r]), int(bounds[r + 1]))
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        for _ in range(5):
            arcs, cs = g.scanRange(lo, hi)
        ts = []
        for _ in range(30):
            torch.cuda.synchronize()
            ev[0].record()
            g.scanRange(lo, hi)
            ev[1].record()
            torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]))
        g.profile(True)
        g.scanRange(lo, hi)
        prof = g.profileRead()
        g.profile(False)
        out["rows"].append({"shard": r, "nodes": hi - lo, "arcs": arcs, "ms_median": float(np.median(ts)), "ms_min": float(np.min(ts)),
                            "kernels_ms_serial": {k: round(v["ms"], 4) for k, v in prof.items()},
                            "kernels_sum_ms": round(sum(v["ms"] for v in prof.values()), 4),
                            "launches": sum(v["launches"] for v in prof.values())})
        g.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
