# round 2, GPU call za (2 GPUs): sharded parity (SPMD, default scheduler, the minimal set) with the H / CX / CZ peephole forced on
set -x
(QB_PEEPHOLE=1 QB_SHARDED_QUICK=2 timeout 48 python -m pytest "tests/test_gpu_sharded.py::test_sharded_matches_oracle[2]" -q 2>&1 | tail -6) > gpurun_out/r2za_sharded_peephole_w2.log 2>&1
cat gpurun_out/r2za_sharded_peephole_w2.log
