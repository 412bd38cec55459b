"""BVGraph.store on the device (bvg_bv_compress) on the 4 M-node / 117 M-arc power-law graph: kernels' time by range size, bits
against the host writer's (16 threads), and a decode of the result."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from webgraph_b200 import tools
from webgraph_b200.bvgraph import BVGraph
base = '/tmp/bvg_bench/bvq'
os.makedirs('/tmp/bvg_bench', exist_ok=True)
t0 = time.perf_counter()
st, off, succ = tools.generate_store(base, 4_000_000, 125_000_000, return_csr=True)
print('generate + host store (%d threads): %.2f s, %d bits' % (os.cpu_count(), time.perf_counter() - t0, st['graph_bits']), flush=True)
t0 = time.perf_counter()
hst = tools.store_csr(base + '-h', off, succ, threads=os.cpu_count())
print('host store_csr alone: %.2f s, %d bits' % (time.perf_counter() - t0, hst['graph_bits']), flush=True)
xor = st['xor_checksum']
for rn in (256, 256, 64, 1024, 4096):
    t0 = time.perf_counter()
    bits, ms = BVGraph.store(base + '-dev', off, succ, rangeNodes=rn)
    print('range_nodes %5d: %d bits (%.4f of the host writer\'s), kernels %.2f ms, call %.2f s' % (rn, bits, bits / hst['graph_bits'], ms, time.perf_counter() - t0), flush=True)
g = BVGraph.load(base + '-dev')
print('decodes back:', g.scanRange(0, g.numNodes()) == (len(succ), xor))
