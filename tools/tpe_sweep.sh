#!/bin/bash
# thread-per-env kernels: register budget x staged/direct observation stores; fused rollout (T=32, all observations) and the
# per-step kernel in the cold / warm regimes.  usage (under gpurun): tools/tpe_sweep.sh "<minb list>"
for m in $1; do for st in 0 1; do
  r=$(NAV3D_TPE_STAGED=$st NAV3D_ROLLOUT_MINB=$m python tools/rollout_bench.py --lanes 1 --T 32 --reps 3 --only all_obs 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('%.1f' % (d['all_obs+reward+done']['ms_per_env_step_batch']*1e3))")
  NAV3D_TPE_STAGED=$st NAV3D_MINB=$m python bench.py --steps 20 --warmup 5 --no-extras --e2e-steps 3 --lanes 1 > gpurun_out/tpe_cold_${m}_${st}.json 2>/dev/null
  NAV3D_TPE_STAGED=$st NAV3D_MINB=$m python bench.py --steps 400 --warmup 100 --no-extras --e2e-steps 3 --lanes 1 > gpurun_out/tpe_warm_${m}_${st}.json 2>/dev/null
  python - <<PY
import json
c=json.load(open("gpurun_out/tpe_cold_${m}_${st}.json")); w=json.load(open("gpurun_out/tpe_warm_${m}_${st}.json"))
print("minb $m staged $st: fused all-obs $r us/step | step cold %.1f us (%.3f)  warm %.1f us (%.3f)" % (c["ms_per_step"]*1e3, c["roofline"]["frac"], w["ms_per_step"]*1e3, w["roofline"]["frac"]))
PY
done; done
