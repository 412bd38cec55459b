"""k host threads, each draining its own cursor over a node range split as ImmutableGraph.splitNodeIterators does: the
NodeIterator route's aggregate throughput (bvg_cursor_drain in C, GIL released).  Usage: python profiles/cursor_threads.py"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import concurrent.futures as cf
import bench
from webgraph_b200 import bvgraph
class A: pass
args = A(); args.nodes=0; args.arcs=0; args.seed=0x5EED; args.max_degree=1<<22; args.workdir=os.environ.get("BVG_BENCH_DIR", "/tmp/bvg_bench")
b = bench.Bench(args, sys.argv[1] if len(sys.argv) > 1 else "powerlaw", 0, 0, 1)
g = b.open_shard()
L = b.L
nodes = b.n_total
def run(k, nodes):
    step = (nodes + k - 1) // k
    def drain(i):
        cur = C.c_void_p()
        bvgraph._check(L.bvg_cursor_open(g.handle, i * step, min(nodes, (i + 1) * step), C.byref(cur)))
        cn, ca, cc = C.c_int64(), C.c_int64(), C.c_uint64()
        bvgraph._check(L.bvg_cursor_drain(cur, -1, C.byref(cn), C.byref(ca), C.byref(cc)))
        L.bvg_cursor_close(cur)
        return ca.value, cc.value
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(k) as ex:
        res = list(ex.map(drain, range(k)))
    dt = time.perf_counter() - t0
    arcs = sum(r[0] for r in res); cs = 0
    for r in res: cs ^= r[1]
    return arcs, cs, dt
run(16, nodes)  # warm: pinned buffers of 16 cursors, schedules
for k in (1, 2, 4, 8, 16, 12, 16):
    arcs, cs, dt = run(k, nodes)
    ok = (arcs, cs) == (b.m_total, int(b.st["xor_checksum"]))
    print("%2d threads: %.2f G edges/s (%d arcs in %.1f ms) %s" % (k, arcs / dt / 1e9, arcs, dt * 1e3, "checksum ok" if ok else "MISMATCH"), flush=True)
