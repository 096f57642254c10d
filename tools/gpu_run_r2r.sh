# round 2, GPU call r (4 GPUs): the 4-GPU bench line (33-qubit strong scaling, parity_check against the oracle, no 35-qubit target)
# on the final build of the round
set -x
(timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --no-target 2> gpurun_out/r2r_bench_w4.err | tail -1) > gpurun_out/r2r_bench_n33_4gpu.json
tail -n 4 gpurun_out/r2r_bench_w4.err
head -c 1200 gpurun_out/r2r_bench_n33_4gpu.json
