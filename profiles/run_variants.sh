#!/bin/bash
# usage: run_variants.sh TAG "ENV1=.. ENV2=.." "ENV..." ...   -- one short bench per environment setting, summary lines
TAG=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 2 > gpurun_out/${TAG}_v$i.json 2> gpurun_out/${TAG}_v$i.err
  python - "$envs" gpurun_out/${TAG}_v$i.json gpurun_out/${TAG}_v$i.err <<'PY'
import json,sys
envs,f,e=sys.argv[1:4]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    k={a:round(b,3) for a,b in d["roofline"]["step_kernels_ms"].items()}
    print("[%s] ms/step %.3f  e2e %.2f G/s  %s" % (envs, d["ms_per_step"], d["e2e"]["value"]/1e9, k))
except Exception as ex:
    print("[%s] ERR %r %s %s" % (envs, ex, open(f).read()[-1500:], open(e).read()[-1500:]))
PY
done
