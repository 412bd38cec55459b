import os, sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from webgraph_b200 import tools, bvgraph
from webgraph_b200.efgraph import EFGraph
base='/tmp/bvg_bench/efq'
os.makedirs('/tmp/bvg_bench', exist_ok=True)
if not os.path.exists(base+'-ef.graph'):
    st, off, succ = tools.generate_store(base, 4_000_000, 125_000_000, return_csr=True)
    tools.store_ef(base+'-ef', off, succ, threads=16)
g=EFGraph.load(base+'-ef')
n=g.numNodes()
ev=[torch.cuda.Event(enable_timing=True) for _ in range(2)]
ts=[]
for i in range(8):
    torch.cuda.synchronize(); ev[0].record(); r=g.scanRange(0,n); ev[1].record(); torch.cuda.synchronize(); ts.append(ev[0].elapsed_time(ev[1]))
print(os.environ.get('BVG_EF_SMALL'), round(float(np.median(ts[3:])),3), r)
