mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/x6_gputests.log 2>&1; tail -4 gpurun_out/x6_gputests.log
timeout 300 python scripts/stream_bench.py > gpurun_out/x6_stream.jsonl 2>&1
timeout 600 python scripts/c5_bench.py > gpurun_out/x6_c5.log 2>&1; tail -1 gpurun_out/x6_c5.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/x6_launches_c5.csv python scripts/c5_bench.py > /dev/null 2>&1
timeout 900 python bench.py > gpurun_out/x6_bench_1gpu.json 2> gpurun_out/x6_bench_1gpu.err; cut -c1-300 gpurun_out/x6_bench_1gpu.json
