for l in 1 2; do for m in 3 4; do for sm in 0 100000 200000; do
  NAV3D_ROLLOUT_MINB=$m NAV3D_ROLLOUT_SMEM=$sm python tools/rollout_bench.py --lanes $l --T 32 --reps 3 > gpurun_out/ro2_${l}_${m}_${sm}.json 2>gpurun_out/ro2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ro2_${l}_${m}_${sm}.json"))
print("lanes $l minb $m smem $sm:", " ".join("%s %.1f us" % (k, v["ms_per_env_step_batch"]*1e3) for k,v in d.items() if isinstance(v, dict)))
PY
done; done; done
