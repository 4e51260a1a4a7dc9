mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_generic.py -x -q > gpurun_out/x7_tests.log 2>&1; tail -15 gpurun_out/x7_tests.log
GEN_SCHED=1 GEN_DTYPES=float32 timeout 300 python scripts/generic_bench.py 100 600 6000 60000 > gpurun_out/x7_generic_lane.jsonl 2>&1
GEN_SCHED=2 GEN_CPW=0,1,2,3,4 GEN_DTYPES=float32 timeout 600 python scripts/generic_bench.py 100 600 6000 60000 > gpurun_out/x7_generic_group_f32.jsonl 2>&1
GEN_SCHED=2 GEN_CPW=0,1,2,4 GEN_DTYPES=float64 timeout 600 python scripts/generic_bench.py 100 6000 > gpurun_out/x7_generic_group_f64.jsonl 2>&1
