# round 2, GPU call x (1 GPU): the example drivers incl. the reference's shors_working test
set -x
(timeout 400 python -m pytest tests/test_gpu_examples.py tests/test_gpu_reference_suite.py -q 2>&1 | tail -5) > gpurun_out/r2x_examples_tests.log 2>&1
cat gpurun_out/r2x_examples_tests.log
