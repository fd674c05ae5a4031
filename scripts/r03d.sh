set -x
mkdir -p gpurun_out/r03d
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r03d/pytest.log 2>&1; tail -4 gpurun_out/r03d/pytest.log
B="timeout 300 python bench.py --no-cpu-baseline --no-eager-baseline"
$B > gpurun_out/r03d/bench.json 2> gpurun_out/r03d/bench.err; cat gpurun_out/r03d/bench.json; tail -3 gpurun_out/r03d/bench.err
B200_TIMING=2 $B --steps 2 --warmup 3 > gpurun_out/r03d/bench_timing.json 2> gpurun_out/r03d/bench_timing.err; grep "b200 " gpurun_out/r03d/bench_timing.err | tail -3
B200_TIMING=3 $B --steps 1 --warmup 3 > gpurun_out/r03d/bench_prof.json 2> gpurun_out/r03d/bench_prof.err; grep -A22 "Ordered by" gpurun_out/r03d/bench_prof.err | tail -24
timeout 400 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r03d/trunk_launches_296.csv python scripts/prof_emb.py emb 296 > gpurun_out/r03d/ncu_emb.log 2>&1; tail -2 gpurun_out/r03d/ncu_emb.log
python scripts/ncu_trunk_traffic.py gpurun_out/r03d/trunk_launches_296.csv gpurun_out/r03d/trunk_traffic.json 296 | head -8
