#!/bin/bash
# Round-2 evidence pass (GPU box): parity tests, the bench line, the ncu launch list and one full capture of the block kernel,
# and the batch-size / resident-warp sweeps.   usage: scripts/gpu_round2_evidence.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_gputests.log 2>&1
tail -3 gpurun_out/${TAG}_gputests.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
cut -c1-600 gpurun_out/${TAG}_bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_ref.err
LIGHT="--no-cpu-baseline --no-secondary --no-config4 --no-joints-e2e --no-ref-iterates"   # the headline step only: first-frame + block kernel, and the end-to-end call
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 $LIGHT > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:leg_solve_block -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_block \
    python bench.py --steps 1 --warmup 3 $LIGHT >> gpurun_out/${TAG}_ncu_bench.log 2>&1
ls -la gpurun_out/${TAG}_prof_block.ncu-rep
timeout 600 python scripts/block_bench.py quick > gpurun_out/${TAG}_block_vs_pipe.jsonl 2>&1
timeout 600 python scripts/block_bench.py r > gpurun_out/${TAG}_block_resident_sweep.jsonl 2>&1
timeout 300 python scripts/block_bench.py variants > gpurun_out/${TAG}_block_variants.jsonl 2>&1
tail -4 gpurun_out/${TAG}_block_vs_pipe.jsonl
