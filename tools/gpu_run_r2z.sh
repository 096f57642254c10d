# round 2, GPU call z (1 GPU): the bench line with the H / CX / CZ peephole (parity_check runs with it forced on)
set -x
timeout 600 python bench.py --no-diagnostics --no-cpu-baseline > gpurun_out/r2z_bench_n33_peephole.json 2> gpurun_out/r2z_bench.err
tail -c 400 gpurun_out/r2z_bench.err
head -c 700 gpurun_out/r2z_bench_n33_peephole.json
