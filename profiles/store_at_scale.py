"""BVGraph.store and EFGraph.store on the device at the benchmark's scale (32 M nodes, 1 B arcs: streams beyond 2^32 bits):
what they write must load and scan back to the generator's checksum."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from webgraph_b200 import tools
from webgraph_b200.bvgraph import BVGraph
from webgraph_b200.efgraph import EFGraph
base = '/tmp/bvg_bench/scale'
os.makedirs('/tmp/bvg_bench', exist_ok=True)
t0 = time.perf_counter()
st, off, succ = tools.generate_store(base, 32_000_000, 1_070_000_000, return_csr=True)
print('generated + host store: %.1f s, %d arcs, %d bits' % (time.perf_counter() - t0, len(succ), st['graph_bits']), flush=True)
want = (len(succ), st['xor_checksum'])
t0 = time.perf_counter()
bits, ms = BVGraph.store(base + '-dev', off, succ)
print('BVGraph.store on the device: %d bits (%.4f of the host writer\'s), kernels %.0f ms, call %.1f s' % (bits, bits / st['graph_bits'], ms, time.perf_counter() - t0), flush=True)
g = BVGraph.load(base + '-dev')
print('  scans back:', g.scanRange(0, g.numNodes()) == want, flush=True)
g.close()
t0 = time.perf_counter()
ebits, ems = EFGraph.store(base + '-ef', off, succ)
print('EFGraph.store on the device: %d bits (%.2f per arc), kernels %.1f ms, call %.1f s' % (ebits, ebits / len(succ), ems, time.perf_counter() - t0), flush=True)
e = EFGraph.load(base + '-ef')
t0 = time.perf_counter()
r = e.scanRange(0, e.numNodes())
print('  scans back:', r == want, 'in %.1f ms' % ((time.perf_counter() - t0) * 1e3), flush=True)
