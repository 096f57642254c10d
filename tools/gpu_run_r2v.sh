# round 2, GPU call v (1 GPU): the bench line with the encoder-aware second-tier planner
set -x
timeout 900 python bench.py > gpurun_out/r2v_bench_n33.json 2> gpurun_out/r2v_bench.err
tail -c 400 gpurun_out/r2v_bench.err
head -c 900 gpurun_out/r2v_bench_n33.json
