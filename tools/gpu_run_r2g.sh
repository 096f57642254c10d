# round 2, GPU call g (N = 4 or 8 GPUs; usage: bash tools/gpu_run_r2g.sh N): sharded parity (SPMD default scheduler incl. placement
# memory, in-process mode) and the N-GPU bench line (33-qubit strong scaling + the 35-qubit target)
N=$1
set -x
nvidia-smi -L | head -8
(QB_SHARDED_QUICK=1 timeout 280 python -m pytest "tests/test_gpu_sharded.py::test_sharded_matches_oracle[$N]" -q 2>&1 | tail -12) > gpurun_out/r2g_sharded_default_w$N.log 2>&1
(timeout 200 python -m pytest "tests/test_gpu_inprocess.py::test_stock_constructor_shards_in_process[$N]" -q 2>&1 | tail -12) > gpurun_out/r2g_inprocess_w$N.log 2>&1
cat gpurun_out/r2g_sharded_default_w$N.log gpurun_out/r2g_inprocess_w$N.log
(timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N 2> gpurun_out/r2g_bench_w$N.err | tail -1) > gpurun_out/r2g_bench_n33_${N}gpu.json
tail -n 4 gpurun_out/r2g_bench_w$N.err
ls -la gpurun_out/ | tail -6
