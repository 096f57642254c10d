# round 2, GPU call j (1 GPU): the pipelined kernel shape (bulk-async tile ring) -- parity, then timing against the default shape
set -x
(timeout 200 python -m pytest tests/test_gpu_jit.py -q -k "66" -x 2>&1 | tail -6) > gpurun_out/r2j_pipelined_tests.log 2>&1
cat gpurun_out/r2j_pipelined_tests.log
(timeout 150 python tools/bench_jit.py 30 layers 2,66 2>&1 | tail -3) > gpurun_out/r2j_jit_layers30.log 2>&1
(timeout 100 python tools/bench_jit.py 30 qft 2,66 2>&1 | tail -3) > gpurun_out/r2j_jit_qft30.log 2>&1
cat gpurun_out/r2j_jit_layers30.log gpurun_out/r2j_jit_qft30.log
