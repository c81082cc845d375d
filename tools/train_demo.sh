#!/bin/bash
# End-to-end demonstration of BASELINE configs[4] on one GPU: train/Grid_Train.py (--native shape, 1024 envs) on
# rooms/P1_training for STEPS env steps, then evaluate the final checkpoint on rooms/P1_evaluate and run
# train/evaluate_grid.py over all checkpoints.  usage: [NPROC=8] tools/train_demo.sh [STEPS] [OUTDIR]
# NPROC > 1: data parallel under torchrun (one rank per GPU, NCCL gradient all-reduce); evaluate_grid.py is skipped.
STEPS=${1:-30000000}
OUT=${2:-gpurun_out/demo}
rm -rf "$OUT"; mkdir -p "$OUT"
T0=$(date +%s.%N)
NPROC=${NPROC:-1}
if [ "$NPROC" -gt 1 ]; then LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NPROC --master-addr 127.0.0.1 --master-port 29520 -m train.Grid_Train"; else LAUNCH="python -m train.Grid_Train"; fi
$LAUNCH --native --num-envs 1024 --steps "$STEPS" --save-dir "$OUT/ckpt" \
    --eval-freq ${EVAL_FREQ:-5000000} > "$OUT/train.log" 2>&1
echo "train wall $(python -c "import time; print(round(time.time() - $T0, 1))") s" >> "$OUT/train.log"
grep -E "^\| iter" "$OUT/train.log" | awk 'NR % 20 == 1' | cut -c1-150
grep -E "Eval num_timesteps|train wall" "$OUT/train.log"
python - "$OUT" <<'PY'
import glob, os, re, sys
sys.path.insert(0, "."); import _nav3d_path
from nav3d.experiment import evaluate_checkpoint, make_vec_env
from nav3d.ppo import RecurrentPPO
out = sys.argv[1]
files = sorted(glob.glob(os.path.join(out, "ckpt", "*.zip")), key=lambda f: int(re.search(r"_s(\d+)_", f).group(1)))
for f in (files[0], files[len(files) // 2], files[-1]):
    env = make_vec_env("./rooms/P1_evaluate", 10, 10, 0)
    model = RecurrentPPO.load(f, env=None, device=env.device)
    agg, rows = evaluate_checkpoint(model, env, os.path.basename(f)[:-4], 10)
    env.close()
    print("P1_evaluate", os.path.basename(f)[-24:], {k: round(v, 2) for k, v in agg.items()})
PY
if [ "$NPROC" -le 1 ]; then
python -m train.evaluate_grid --models-dir "$OUT/ckpt" --episodes 10 --txt "$OUT/exp3_viewDistance.txt" --csv "$OUT/exp3_viewDistance.csv" > "$OUT/eval.log" 2>&1
cat "$OUT/exp3_viewDistance.txt" | cut -c1-140
fi
rm -rf "$OUT/ckpt"
