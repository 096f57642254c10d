# round 2, GPU call h (2 GPUs): exchanges carried by a tile pass's stores -- parity (SPMD + in-process), then the N = 2 bench line
set -x
(QB_SHARDED_EXTRA=1 timeout 300 python -m pytest "tests/test_gpu_sharded.py::test_sharded_matches_oracle[2]" -q 2>&1 | tail -12) > gpurun_out/r2h_sharded_default_w2.log 2>&1
(timeout 200 python -m pytest "tests/test_gpu_inprocess.py::test_stock_constructor_shards_in_process[2]" -q 2>&1 | tail -8) > gpurun_out/r2h_inprocess_w2.log 2>&1
cat gpurun_out/r2h_sharded_default_w2.log gpurun_out/r2h_inprocess_w2.log
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --no-target 2> gpurun_out/r2h_bench_w2.err | tail -1) > gpurun_out/r2h_bench_n33_2gpu.json
tail -n 3 gpurun_out/r2h_bench_w2.err
