# round 2, GPU call f (1 GPU): kernel-variant / planner table, the planner's 256-byte-segment experiment at full size
set -x
(timeout 120 python -m pytest tests/test_gpu_jit.py -q -k "cache or asynchronous" 2>&1 | tail -5) > gpurun_out/r2f_jit_cache_tests.log 2>&1
(timeout 260 python tools/bench_diagnostics.py 30 240 | tail -1) > gpurun_out/r2f_diagnostics_n30.json 2> gpurun_out/r2f_diag.err
(QB_PLAN_MIN_LC=4 timeout 200 python bench.py --steps 2 --no-cpu-baseline --no-diagnostics --no-c3 --no-parity-check 2> gpurun_out/r2f_lc4.err | tail -1) > gpurun_out/r2f_bench_n33_lc4.json
cat gpurun_out/r2f_jit_cache_tests.log
tail -3 gpurun_out/r2f_lc4.err gpurun_out/r2f_diag.err
