#!/bin/bash
# usage: run_variants.sh TAG "ENV1=.. ENV2=.." "ENV..." ...   -- one short bench per environment setting, summary lines.
# BENCH_ARGS overrides the bench flags (default: no extras, no CPU baseline, 2 e2e steps: ~12 s of GPU time per variant);
# e.g. BENCH_ARGS="--steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1" keeps the extras (random access, materialise,
# NodeIterator, open without .offsets) in the line.
TAG=$1; shift
BENCH_ARGS=${BENCH_ARGS:---steps 10 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 2 --no-second-workload}
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py $BENCH_ARGS > gpurun_out/${TAG}_v$i.json 2> gpurun_out/${TAG}_v$i.err
  python - "$envs" gpurun_out/${TAG}_v$i.json gpurun_out/${TAG}_v$i.err <<'PY'
import json,sys
envs,f,e=sys.argv[1:4]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    k={a:round(b,3) for a,b in d["roofline"]["step_kernels_ms"].items()}
    x=(d["roofline"].get("other_configs") or {})
    ow=x.get("open_without_offsets") or {}
    tail="  open w/o offsets %.0f ms" % ow["ms"] if "ms" in ow else ""
    print("[%s] ms/step %.3f  e2e %.2f G/s  %s%s" % (envs, d["ms_per_step"], d["e2e"]["value"]/1e9, k, tail))
except Exception as ex:
    print("[%s] ERR %r %s %s" % (envs, ex, open(f).read()[-1500:], open(e).read()[-1500:]))
PY
done
