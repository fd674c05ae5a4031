set -x
mkdir -p gpurun_out/r03j
timeout 900 python -m pytest tests -m gpu -q -k "vbx or linkage or end_to_end or bench_workload or edge_cases or exclude_overlap or clustering_class or cfg3" > gpurun_out/r03j/pytest.log 2>&1; tail -4 gpurun_out/r03j/pytest.log
B200_TIMING=2 timeout 300 python bench.py --no-cpu-baseline --no-eager-baseline --steps 3 --warmup 3 > gpurun_out/r03j/bench_timing.json 2> gpurun_out/r03j/bench_timing.err; grep "b200 " gpurun_out/r03j/bench_timing.err | tail -3; grep "^{" gpurun_out/r03j/bench_timing.json | cut -c1-300
