# round 2, GPU call t (8 GPUs): the driver's own N = 8 command on the final build (33-qubit strong scaling + the 35-qubit target)
set -x
(timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 8 --steps 3 --warmup 3 2> gpurun_out/r2t_bench_w8.err | tail -1) > gpurun_out/r2t_bench_n33_8gpu.json
tail -n 3 gpurun_out/r2t_bench_w8.err | cut -c1-300
head -c 700 gpurun_out/r2t_bench_n33_8gpu.json
