# round 2, GPU call y (1 GPU): the bench line with 20 timed steps (an even number of steps averages the two alternating gate lists)
set -x
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2y_bench_n33_20steps.json 2> gpurun_out/r2y_bench.err
tail -c 300 gpurun_out/r2y_bench.err
head -c 600 gpurun_out/r2y_bench_n33_20steps.json
