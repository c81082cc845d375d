#!/bin/bash
# usage: tools/sweep.sh "<lanes list>" "<minb list>" [extra bench args]
for l in $1; do for m in $2; do
  NAV3D_MINB=$m python bench.py --steps 200 --warmup 10 --no-extras --lanes $l $3 > gpurun_out/sw_${l}_${m}.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sw_${l}_${m}.log").read().strip().splitlines()[-1]); print("lanes",$l,"minb",$m,"steps/s %.3e"%d["value"],"ms %.4f"%d["ms_per_step"],"frac %.3f"%d["roofline"]["frac"], "e2e %.3e"%d["e2e"]["value"])
except Exception as e: print("lanes",$l,"minb",$m,"FAILED",e)
PY
done; done
