# round 2, GPU call q (2 GPUs): sharded parity (SPMD default scheduler + in-process mode) on the final build -- the generic dense
# gates of the mixed circuits now run in place (K3i) on every shard
set -x
(QB_SHARDED_EXTRA=1 timeout 300 python -m pytest "tests/test_gpu_sharded.py::test_sharded_matches_oracle[2]" -q 2>&1 | tail -12) > gpurun_out/r2q_sharded_default_w2.log 2>&1
(timeout 200 python -m pytest "tests/test_gpu_inprocess.py::test_stock_constructor_shards_in_process[2]" -q 2>&1 | tail -8) > gpurun_out/r2q_inprocess_w2.log 2>&1
cat gpurun_out/r2q_sharded_default_w2.log gpurun_out/r2q_inprocess_w2.log
