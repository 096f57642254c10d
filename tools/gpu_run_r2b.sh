set -x
nvidia-smi -L
(timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2b_gpu_tests.log 2>&1
(timeout 600 python bench.py 2> gpurun_out/r2b_bench.err | tail -1) > gpurun_out/r2b_bench_n33.json
tail -3 gpurun_out/r2b_bench.err
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2b_launches_bench_n33.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-diagnostics --no-c3 --no-parity-check > gpurun_out/r2b_bench_under_ncu.log 2>&1)
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:qb_jit_pass -s 20 -c 3 -o gpurun_out/r2b_jit_pass_n28 python tools/bench_jit.py 28 layers 2 > gpurun_out/r2b_ncu_full.log 2>&1)
ls -la gpurun_out/
cat gpurun_out/r2b_gpu_tests.log
