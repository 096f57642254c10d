# round 2, GPU call s (1 GPU): compute-sanitizer over the kernels added at the end of the round -- the in-place generic gate kernel
# (K3i: shared-memory staging, warp shuffles) and the prefetching ring (cp.async into shared memory, named barriers, done[] flags)
set -x
F='grep -v "Host Frame\|^=========         in \|^=========                in "'
(timeout 400 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_parity.py -q -x -k "generic_kernel or wide_dense_gates_in_place and not 13-13 and not 14-12" 2>&1 | grep -v "Host Frame\|^=========         in \|^=========                in " | tail -12) > gpurun_out/r2s_sanitizer_memcheck_k3i.log 2>&1
(timeout 400 compute-sanitizer --tool racecheck --print-limit 8 python -m pytest tests/test_gpu_parity.py -q -x -k "generic_kernel or wide_dense_gates_in_place and not 13-13 and not 14-12 and not 20-5" 2>&1 | grep -v "Host Frame\|^=========         in \|^=========                in " | tail -12) > gpurun_out/r2s_sanitizer_racecheck_k3i.log 2>&1
(timeout 400 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_jit.py -q -x -k "66 and (12 or 14 or 15 or 16)" 2>&1 | grep -v "Host Frame\|^=========         in \|^=========                in " | tail -12) > gpurun_out/r2s_sanitizer_memcheck_ring.log 2>&1
(timeout 400 compute-sanitizer --tool racecheck --print-limit 8 python -m pytest tests/test_gpu_jit.py -q -x -k "66 and (12 or 14 or 15)" 2>&1 | grep -v "Host Frame\|^=========         in \|^=========                in " | tail -12) > gpurun_out/r2s_sanitizer_racecheck_ring.log 2>&1
tail -n 6 gpurun_out/r2s_sanitizer_*.log
