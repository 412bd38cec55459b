"""C4 (random-access batches, bvg_successors_batch with device buffers) by batch size, through k_random (one thread per query)
and through the range kernels (every chain marked, rows gathered): BVG_RANDOM_RANGE_PCT in the environment picks the route
(100000 = never the range route, 0 = always).  Prints ms per batch and the per-kernel times of one batch."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from webgraph_b200 import bvgraph  # noqa: E402


def main():
    sys.argv = sys.argv[:1]
    args = bench.parse_args()
    base, st = bench.graph_files(args, "powerlaw", 0, lambda: None)
    g = bvgraph.BVGraph.load(base)
    L = bvgraph.lib()
    n = g.numNodes()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    out = {"route_pct": os.environ.get("BVG_RANDOM_RANGE_PCT", "default"), "rows": []}
    for nq in (10_000_000, 3_000_000, 1_000_000, 300_000, 100_000, 10_000):
        xs = torch.randint(0, n, (nq,), device="cuda", dtype=torch.int32, generator=gen)
        qoff = torch.zeros(nq + 1, dtype=torch.int64, device="cuda")
        bvgraph._check(L.bvg_successors_batch(g.handle, xs.data_ptr(), nq, qoff.data_ptr(), None, 0, 1))
        torch.cuda.synchronize()
        qarcs = int(qoff[-1].item())
        qout = torch.empty(max(qarcs, 1), dtype=torch.int32, device="cuda")
        ts = []
        for rep in range(8):
            ev[0].record()
            bvgraph._check(L.bvg_successors_batch(g.handle, xs.data_ptr(), nq, qoff.data_ptr(), qout.data_ptr(), qarcs, 1))
            ev[1].record()
            torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]))
        g.profile(True)
        bvgraph._check(L.bvg_successors_batch(g.handle, xs.data_ptr(), nq, qoff.data_ptr(), qout.data_ptr(), qarcs, 1))
        torch.cuda.synchronize()
        prof = {k: round(v["ms"], 3) for k, v in g.profileRead().items()}
        g.profile(False)
        out["rows"].append({"queries": nq, "arcs": qarcs, "ms": float(np.median(ts[3:])), "kernels_ms": prof})
        del xs, qoff, qout
    print(json.dumps(out))


if __name__ == "__main__":
    main()
