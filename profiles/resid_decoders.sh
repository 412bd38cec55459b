#!/bin/bash
# usage: resid_decoders.sh TAG  -- builds profiles/resid_decoders.cu, runs it on the residual sections of both benchmark graphs
# (bench.py's cached files), then once more under ncu for warp instructions / lane slots per residual.
TAG=${1:-r02_decoders}
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I webgraph_b200/csrc/cuda -I include profiles/resid_decoders.cu -o /tmp/resid_decoders
python - <<'PY'
import bench
class A: pass
a = A(); a.nodes = 0; a.arcs = 0; a.seed = 0x5EED; a.max_degree = 1 << 22; a.workdir = "/tmp/bvg_bench"
for w in ("powerlaw", "weblike"):
    print(w, bench.graph_files(a, w, 0, lambda: None)[0])
PY
PL=$(ls -d /tmp/bvg_bench/pl_*/g.graph | head -1); PL=${PL%.graph}
WEB=$(ls -d /tmp/bvg_bench/web_*/g.graph | head -1); WEB=${WEB%.graph}
for g in $PL $WEB; do
  for minrc in 1 16 64; do
    /tmp/resid_decoders $g $minrc | tee -a gpurun_out/${TAG}.jsonl
  done
done
ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:'k_lane|k_table|k_warp' --csv --log-file gpurun_out/${TAG}_ncu_pl.csv /tmp/resid_decoders $PL 64 > /dev/null
ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:'k_lane|k_table|k_warp' --csv --log-file gpurun_out/${TAG}_ncu_web.csv /tmp/resid_decoders $WEB 64 > /dev/null
