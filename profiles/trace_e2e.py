import sys, os, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from webgraph_b200 import bvgraph
class A: pass
args = A(); args.nodes=0; args.arcs=0; args.seed=0x5EED; args.max_degree=1<<22; args.workdir='/tmp/bvg_bench'
base, st = bench.graph_files(args, 'powerlaw', 0, lambda: None)
L = bvgraph.lib()
graph = torch.from_numpy(np.fromfile(base+'.graph', dtype=np.uint8)).pin_memory()
offs = torch.from_numpy(np.fromfile(base+'.offsets', dtype=np.uint8)).pin_memory()
n, m = st['nodes'], st['arcs']
a_out, c_out = C.c_int64(), C.c_uint64()
def one(pieces):
    t=time.perf_counter()
    bvgraph._check(L.bvg_scan_memory(graph.data_ptr(), graph.numel(), offs.data_ptr(), offs.numel(), n, m, 7,3,4,3,0,0, 0, n, pieces, C.byref(a_out), C.byref(c_out)))
    return (time.perf_counter()-t)*1e3
for _ in range(3): one(4)
os.environ['BVG_TRACE']='1'
print('traced', one(4))
del os.environ['BVG_TRACE']
print([round(one(4),2) for _ in range(3)])
