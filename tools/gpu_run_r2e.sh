# round 2, GPU call e (2 GPUs): placement memory (repeated circuits keep their specialised kernels under the exchange scheduler)
set -x
nvidia-smi -L
(QB_SHARDED_EXTRA=1 timeout 300 python -m pytest "tests/test_gpu_sharded.py::test_sharded_matches_oracle[2]" -q 2>&1 | tail -12) > gpurun_out/r2e_sharded_default_w2.log 2>&1
cat gpurun_out/r2e_sharded_default_w2.log
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 2> gpurun_out/r2e_bench_w2.err | tail -1) > gpurun_out/r2e_bench_n33_2gpu.json
tail -3 gpurun_out/r2e_bench_w2.err
ls -la gpurun_out/
