"""EFGraph.store on the device (bvg_ef_compress) timed twice on the 4 M-node / 117 M-arc graph (second call warm)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from webgraph_b200 import tools
from webgraph_b200.efgraph import EFGraph
base = '/tmp/bvg_bench/efq'
os.makedirs('/tmp/bvg_bench', exist_ok=True)
st, off, succ = tools.generate_store(base, 4_000_000, 125_000_000, return_csr=True)
for i in range(3):
    t0 = time.perf_counter()
    bits, ms = EFGraph.store(base + '-dev', off, succ)
    print(i, bits, 'kernels %.3f ms' % ms, 'call %.3f s' % (time.perf_counter() - t0), flush=True)
