#!/bin/bash
# ncu captures of the solver kernel (GPU box).  usage: scripts/ncu_solver.sh <tag> [schedule] [trials] [frames]
TAG=${1:-r1}; SCHED=${2:-0}; TRIALS=${3:-1000}; FRAMES=${4:-100}
mkdir -p gpurun_out
# launch list of one short bench run (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --trials $TRIALS --frames $FRAMES --schedule $SCHED --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
# full capture of the solver kernel (3rd launch = after warm-up)
ncu --set full --clock-control none --import-source on -k regex:leg_solve -s 3 -c 1 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 3 --trials $TRIALS --frames $FRAMES --schedule $SCHED --no-cpu-baseline >> gpurun_out/ncu_bench_${TAG}.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
