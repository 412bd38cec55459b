"""bvg_scan_memory (the e2e pass: host buffers in, result out) timed over an eighth of the 1 B-arc graph in one piece and over the
whole graph in 4 / 6 pieces, for sweeps of BVG_ONESHOT_D / BVG_ONESHOT_PART / BVG_E2E_TAPER (environment)."""
import sys, os, time, ctypes as C
sys.path.insert(0, '.')
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
bounds = bvgraph.plan_shards(base, 8)
a_out, c_out = C.c_int64(), C.c_uint64()
def one(lo, hi, pieces):
    t=time.perf_counter()
    bvgraph._check(L.bvg_scan_memory(graph.data_ptr(), graph.numel(), offs.data_ptr(), offs.numel(), n, m, 7,3,4,3,0,0, lo, hi, pieces, C.byref(a_out), C.byref(c_out)))
    assert (lo, hi) != (0, n) or (a_out.value, c_out.value) == (m, st['xor_checksum'])
    return (time.perf_counter()-t)*1e3
tag = ' '.join('%s=%s' % (k, os.environ[k]) for k in ('BVG_ONESHOT_D', 'BVG_ONESHOT_PART', 'BVG_E2E_TAPER') if k in os.environ)
for lo, hi, p in ((bounds[3], bounds[4], 1), (0, n, 3), (0, n, 4), (0, n, 5), (0, n, 6)):
    for _ in range(3): one(lo, hi, p)
    print(tag, 'range', lo, hi, 'pieces', p, [round(one(lo, hi, p),2) for _ in range(4)])
