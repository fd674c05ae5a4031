set -x
mkdir -p gpurun_out/r03b
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r03b/pytest.log 2>&1; tail -4 gpurun_out/r03b/pytest.log
timeout 600 python bench.py > gpurun_out/r03b/bench.json 2> gpurun_out/r03b/bench.err; cat gpurun_out/r03b/bench.json; tail -3 gpurun_out/r03b/bench.err
B200_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r03b/bench_timing.json 2> gpurun_out/r03b/bench_timing.err; grep "b200 timing" gpurun_out/r03b/bench_timing.err | tail -3
