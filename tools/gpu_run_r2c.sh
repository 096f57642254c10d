# round 2, GPU call c (1 GPU): full GPU test suite, ncu captures of the specialised kernel, diagnostics table, sanitizer passes
set -x
nvidia-smi -L
(timeout 420 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/r2c_gpu_tests.log 2>&1
(timeout 150 ncu --set full --clock-control none --import-source on -k regex:qb_jit_pass -s 20 -c 2 -f -o gpurun_out/r2c_jit_pass_n28 python tools/bench_jit.py 28 layers 2 > gpurun_out/r2c_ncu_full.log 2>&1)
(timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches_bench_n30.csv python bench.py --qubits 30 --steps 1 --warmup 3 --no-cpu-baseline --no-diagnostics --no-c3 --no-parity-check > gpurun_out/r2c_bench_under_ncu.log 2>&1)
(timeout 200 python tools/bench_diagnostics.py 30 170 | tail -1) > gpurun_out/r2c_diagnostics_n30.json 2> gpurun_out/r2c_diag.err
(timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_jit.py tests/test_gpu_fused.py -q -k "12 or 14 or 13" -x 2>&1 | tail -12) > gpurun_out/r2c_sanitizer_memcheck.log 2>&1
(timeout 200 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_jit.py -q -k "12-layers or 14-mixed" -x 2>&1 | tail -12) > gpurun_out/r2c_sanitizer_racecheck.log 2>&1
ls -la gpurun_out/
cat gpurun_out/r2c_gpu_tests.log gpurun_out/r2c_sanitizer_memcheck.log gpurun_out/r2c_sanitizer_racecheck.log
