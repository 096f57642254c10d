# round 2, GPU call d (2 GPUs): sharded parity (default scheduler, remote scheduler, in-process mode) and the N = 2 bench lines
set -x
nvidia-smi -L
(QB_SHARDED_EXTRA=1 timeout 300 python -m pytest "tests/test_gpu_sharded.py::test_sharded_matches_oracle[2]" -q 2>&1 | tail -12) > gpurun_out/r2d_sharded_default_w2.log 2>&1
(QB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest "tests/test_gpu_sharded.py::test_sharded_remote_passes_match_oracle[2]" -q 2>&1 | tail -12) > gpurun_out/r2d_sharded_remote_w2.log 2>&1
(timeout 300 python -m pytest tests/test_gpu_inprocess.py -q 2>&1 | tail -12) > gpurun_out/r2d_inprocess_w2.log 2>&1
cat gpurun_out/r2d_sharded_default_w2.log gpurun_out/r2d_sharded_remote_w2.log gpurun_out/r2d_inprocess_w2.log
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 2> gpurun_out/r2d_bench_w2.err | tail -1) > gpurun_out/r2d_bench_n33_2gpu.json
tail -3 gpurun_out/r2d_bench_w2.err
(QB_SHARD_REMOTE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 2> gpurun_out/r2d_bench_w2_remote.err | tail -1) > gpurun_out/r2d_bench_n33_2gpu_remote.json
tail -3 gpurun_out/r2d_bench_w2_remote.err
ls -la gpurun_out/
