set -x
mkdir -p gpurun_out/r02x
timeout 300 python scripts/seg_perf.py 4736 tc > gpurun_out/r02x/seg_perf.log 2>&1; cat gpurun_out/r02x/seg_perf.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02x/pytest.log 2>&1; tail -8 gpurun_out/r02x/pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02x/bench.json 2> gpurun_out/r02x/bench.err; tail -c 700 gpurun_out/r02x/bench.json
