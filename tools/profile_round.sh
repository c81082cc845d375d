#!/bin/bash
# Round-end evidence on one B200: bench line, ncu launch lists (env bench + training loop), one full capture of the step
# kernel.  usage (under gpurun): tools/profile_round.sh <tag>      -> files under gpurun_out/
TAG=${1:-r01_v6}
O=gpurun_out
python bench.py --steps 1000 --warmup 50 > $O/bench_$TAG.json 2> $O/bench_$TAG.err || exit 1
CMD="python bench.py --steps 600 --warmup 10 --no-extras --e2e-steps 3"
$CMD > $O/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 5 -c 120 --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_call -s 400 -c 1 -o $O/prof_step_$TAG $CMD > $O/ncu_f_$TAG.log 2>&1
TCMD="python tools/train_bench.py --envs 1024 --batch-envs 512 --n-steps 16 --iters 1 --epochs 1"
$TCMD > $O/plain_train_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/launches_train_$TAG.csv $TCMD > $O/ncu_lt_$TAG.log 2>&1
ls -la $O/*$TAG*
