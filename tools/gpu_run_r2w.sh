# round 2, GPU call w (1 GPU): the planner-facing GPU tests on the last commit (encoder-aware second-tier planner)
set -x
(timeout 500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_jit.py tests/test_gpu_fullsize.py tests/test_gpu_relabel.py tests/test_gpu_circuit.py -q 2>&1 | tail -6) > gpurun_out/r2w_planner_tests.log 2>&1
cat gpurun_out/r2w_planner_tests.log
