# round 2, GPU call p (1 GPU): the whole GPU suite and the default bench line on the final build of the round
set -x
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2p_gpu_tests.log 2>&1
cat gpurun_out/r2p_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r2p_bench_n33.json 2> gpurun_out/r2p_bench.err
tail -c 600 gpurun_out/r2p_bench.err
head -c 1500 gpurun_out/r2p_bench_n33.json
