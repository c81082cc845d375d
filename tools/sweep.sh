#!/bin/bash
# lanes x min-CTAs sweep of the step kernel, cold (20 steps after reset) and warm (600 steps) regimes; under gpurun
# usage: tools/sweep.sh "<lanes list>" "<minb list>" [extra bench args]
for l in $1; do for m in $2; do
  NAV3D_MINB=$m python bench.py --steps 20 --warmup 5 --no-extras --e2e-steps 3 --lanes $l $3 > gpurun_out/sw_cold_${l}_${m}.json 2>/dev/null
  NAV3D_MINB=$m python bench.py --steps 400 --warmup 100 --no-extras --e2e-steps 3 --lanes $l $3 > gpurun_out/sw_warm_${l}_${m}.json 2>/dev/null
  python - <<PY
import json
c=json.load(open("gpurun_out/sw_cold_${l}_${m}.json")); w=json.load(open("gpurun_out/sw_warm_${l}_${m}.json"))
print("lanes $l minb $m  cold %.1f us (%.3f)  warm %.1f us (%.3f)" % (c["ms_per_step"]*1e3, c["roofline"]["frac"], w["ms_per_step"]*1e3, w["roofline"]["frac"]))
PY
done; done
