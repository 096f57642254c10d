# round 2, GPU call u (1 GPU): smoke() and the whole GPU suite on the committed tree at the end of the round
set -x
(timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3) > gpurun_out/r2u_smoke.log 2>&1
cat gpurun_out/r2u_smoke.log
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2u_gpu_tests.log 2>&1
cat gpurun_out/r2u_gpu_tests.log
