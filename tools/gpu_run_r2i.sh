# round 2, GPU call i (1 GPU): full GPU suite on the current tree, then the bench line with the plan cache / second-tier planner
set -x
(timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2i_gpu_tests.log 2>&1
cat gpurun_out/r2i_gpu_tests.log
(timeout 420 python bench.py --no-diagnostics 2> gpurun_out/r2i_bench.err | tail -1) > gpurun_out/r2i_bench_n33.json
tail -n 3 gpurun_out/r2i_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2i_bench_n33.json'))
print({k:d[k] for k in ('value','ms_per_step','passes_per_step','rounds_per_step','warmup')}, d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['jit'], d['e2e']['value'], d.get('c3',{}).get('value'))
"
