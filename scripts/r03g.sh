set -x
mkdir -p gpurun_out/r03g
nvidia-smi --query-gpu=name,serial,clocks.sm,temperature.gpu --format=csv,noheader > gpurun_out/r03g/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r03g/pytest.log 2>&1; tail -6 gpurun_out/r03g/pytest.log; grep "sincnet diagnosis" gpurun_out/r03g/pytest.log
timeout 600 python bench.py > gpurun_out/r03g/bench.json 2> gpurun_out/r03g/bench.err; cat gpurun_out/r03g/bench.json; tail -3 gpurun_out/r03g/bench.err
B200_TIMING=2 timeout 300 python bench.py --no-cpu-baseline --no-eager-baseline --steps 2 --warmup 3 > gpurun_out/r03g/bench_timing.json 2> gpurun_out/r03g/bench_timing.err; grep "b200 " gpurun_out/r03g/bench_timing.err | tail -3
