"""Open of the 1 B-arc graph WITHOUT .offsets (record boundaries found from the stream, bvg_boundaries.cuh) by sub-range size:
BVG_BND_SUB_BITS / BVG_BND_MAX_SUB / BVG_BND_LANES in the environment."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from webgraph_b200 import bvgraph
sys.argv = sys.argv[:1]
args = bench.parse_args()
base, st = bench.graph_files(args, 'powerlaw', 0, lambda: None)
graph = torch.from_numpy(np.fromfile(base + '.graph', dtype=np.uint8)).pin_memory()
L = bvgraph.lib()
ts = []
for i in range(3):
    h = C.c_void_p()
    t0 = time.perf_counter()
    bvgraph._check(L.bvg_open_memory(graph.data_ptr(), graph.numel(), None, 0, st['nodes'], st['arcs'], 7, 3, 4, 3, 0, 0, -1, C.byref(h)))
    ts.append((time.perf_counter() - t0) * 1e3)
    g = bvgraph.BVGraph(h)
    ok = g.scanRange(0, st['nodes']) == (st['arcs'], st['xor_checksum'])
    g.close()
print({k: os.environ.get(k) for k in ('BVG_BND_SUB_BITS', 'BVG_BND_MAX_SUB', 'BVG_BND_LANES', 'BVG_BND_CAP_BITS')}, [round(t, 1) for t in ts], ok)
