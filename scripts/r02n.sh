set -x
mkdir -p gpurun_out/r02n
scripts/micro/gate_bench > gpurun_out/r02n/gate_bench.log 2>&1; scripts/micro/gate_bench5 >> gpurun_out/r02n/gate_bench.log 2>&1; cat gpurun_out/r02n/gate_bench.log
python -m pytest tests -m gpu -q -x > gpurun_out/r02n/pytest.log 2>&1; tail -5 gpurun_out/r02n/pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02n/bench.json 2> gpurun_out/r02n/bench.err; tail -c 1500 gpurun_out/r02n/bench.json
python scripts/seg_perf.py 4736 > gpurun_out/r02n/seg_perf.log 2>&1; cat gpurun_out/r02n/seg_perf.log
