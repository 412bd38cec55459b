"""Arc-label stream on the device: timings of bvg_labels_decode_range (device buffers) and bvg_labels_scan_range
on a power-law graph with the labels of the reference's test (x * succ + x & mask), beside the oracle reading the same
labels node by node on one host core.  Usage: python profiles/labels_bench.py [nodes] [arcs] > profiles/r02_labels.json"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import oracle_binding as ob  # noqa: E402  (cpu_baseline leg only)
from webgraph_b200 import bvgraph, labelling, tools  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 125_000_000
    tmp = tempfile.mkdtemp(prefix="labels_")
    base = os.path.join(tmp, "g")
    st, off, succ = tools.generate_store(base, n, m, return_csr=True)
    src = np.repeat(np.arange(n, dtype=np.int64), np.diff(off))
    out = {"nodes": n, "arcs": int(off[-1]), "max_outdegree": int(np.diff(off).max()), "graph_bits_per_arc": st["graph_bits"] / off[-1]}
    L = bvgraph.lib()
    for name, kind, width, mask in (("gamma", tools.LABEL_GAMMA, 0, (1 << 15) - 1), ("fixed16", tools.LABEL_FIXED, 16, (1 << 16) - 1)):
        values = ((src * succ + src) & mask).astype(np.int32)
        lbase = base + "-" + name
        bits = tools.store_labels(lbase, "g", off, values, kind, width, threads=os.cpu_count() or 1)
        t0 = time.perf_counter()
        alg = labelling.BitStreamArcLabelledImmutableGraph.load(lbase)
        t_open = time.perf_counter() - t0
        arcs = int(off[-1])
        d_vals = torch.empty(arcs, dtype=torch.int32, device="cuda")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        res = {"label_bits": bits, "bits_per_label": bits / arcs, "open_s": t_open}
        for what in ("decode", "scan"):
            times = []
            for it in range(6):
                torch.cuda.synchronize()
                ev[0].record()
                if what == "decode":
                    bvgraph._check(L.bvg_labels_decode_range(alg._h, 0, n, None, d_vals.data_ptr(), arcs, 1, None))
                else:
                    a, nv, cs = alg.scanLabels(0, n)
                ev[1].record()
                torch.cuda.synchronize()
                times.append(ev[0].elapsed_time(ev[1]))
            ms = float(np.median(times[2:]))
            res[what + "_ms"] = ms
            res[what + "_G_labels_per_s"] = arcs / ms / 1e6
            res[what + "_stream_GBps"] = bits / 8 / ms / 1e6
        alg.g.profile(True)
        bvgraph._check(L.bvg_labels_decode_range(alg._h, 0, n, None, d_vals.data_ptr(), arcs, 1, None))
        alg.scanLabels(0, n)
        torch.cuda.synchronize()
        res["kernels_ms_decode_plus_scan"] = alg.g.profileRead()
        alg.g.profile(False)
        assert np.array_equal(d_vals.cpu().numpy(), values)
        lo = np.arange(arcs + 1, dtype=np.int64)
        assert cs == ob.label_checksum(lo, values)
        # the oracle (BitStreamLabelledArcIterator restated) on one core, first nodes holding ~20 M arcs
        orc = ob.load().load_labels(lbase, n)
        upto = int(np.searchsorted(off, min(arcs, 20_000_000)))
        t0 = time.perf_counter()
        _, tot = orc.sequential(0, upto, off, store=False)
        dt = time.perf_counter() - t0
        assert tot == int(values[:off[upto]].astype(np.int64).sum())
        res["cpu_oracle_G_labels_per_s_1core"] = int(off[upto]) / dt / 1e9
        res["cpu_sample"] = "first %d nodes / %d labels read front to back, consume only" % (upto, int(off[upto]))
        orc.close()
        alg.close()
        out[name] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
