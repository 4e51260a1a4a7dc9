mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/x4_tests.log 2>&1; tail -3 gpurun_out/x4_tests.log
timeout 300 python scripts/stream_bench.py > gpurun_out/x4_stream.jsonl 2>&1
grep -E "leg_affine|pchip" gpurun_out/x4_stream.jsonl | cut -c1-160
timeout 600 ncu --set full --clock-control none --import-source on -k regex:leg_affine_fused -s 4 -c 1 -f -o gpurun_out/x4_prof_affine python scripts/stream_bench.py > gpurun_out/x4_ncu_affine.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/x4_launches_c5.csv python scripts/c5_bench.py > gpurun_out/x4_c5.log 2>&1
