mkdir -p gpurun_out
# A: 96-register lean block kernel, resident warps 16..18, against the 128-register build
SEQIK_LIB=build_variants/libseqik_r96.so timeout 600 python scripts/block_bench.py r > gpurun_out/x1_block_r96.jsonl 2>&1
tail -3 gpurun_out/x1_block_r96.jsonl
# B: generic solver, chains per warp
GEN_CPW=0,1,2,3,4,6,8 GEN_DTYPES=float32 timeout 600 python scripts/generic_bench.py 100 6000 60000 > gpurun_out/x1_generic_cpw_f32.jsonl 2>&1
GEN_CPW=0,1,2,3,4,6,8 GEN_DTYPES=float64 timeout 900 python scripts/generic_bench.py 100 6000 > gpurun_out/x1_generic_cpw_f64.jsonl 2>&1
tail -2 gpurun_out/x1_generic_cpw_f64.jsonl
# C: full captures of the cold stream kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mid_quantile|leg_series|pchip' -c 12 -f -o gpurun_out/x1_prof_stream python scripts/stream_bench.py > gpurun_out/x1_ncu_stream.log 2>&1
ls -la gpurun_out/
