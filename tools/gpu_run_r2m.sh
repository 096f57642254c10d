# round 2, GPU call m (1 GPU): (1) the self-fed ring (variant 130) on the cases where the producer-warp shape traps: parity against
# the default shape + time per pass; (2) the in-place generic gate kernel (K3i) through the parity tests that use wide dense gates;
# (3) the producer-warp shape's trap under cuda-gdb (exception type + source line)
set -x
L=gpurun_out/r2m_ring.log
: > $L
for c in "24 layers6" "26 layers" "28 h2" "28 layers" "30 layers" "30 qft"; do
  set -- $c
  if timeout 150 python tools/pipe_bisect.py $1 $2 130 3 >> $L 2>&1; then echo "CASE $c: ok" >> $L; else echo "CASE $c: FAILED rc=$?" >> $L; fi
  tail -4 $L | cut -c1-200
done
grep "CASE\|diff\|variant" $L | cut -c1-200
(timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_suite.py tests/test_gpu_edge.py -q -x 2>&1 | tail -8) > gpurun_out/r2m_k3i_tests.log 2>&1
cat gpurun_out/r2m_k3i_tests.log
(timeout 240 cuda-gdb -batch -ex "set cuda break_on_launch none" -ex run -ex "info cuda kernels" -ex bt -ex "x/4i \$pc" --args python tools/pipe_bisect.py 26 layers 66 2 2>&1 | grep -v "^\[New Thread\|^\[Thread\|warning: " | tail -40) > gpurun_out/r2m_gdb_pipe66.log 2>&1
cut -c1-220 gpurun_out/r2m_gdb_pipe66.log
