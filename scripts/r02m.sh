set -x
mkdir -p gpurun_out/r02m
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "embedding_parity or end_to_end" > gpurun_out/r02m/pytest.log 2>&1; tail -3 gpurun_out/r02m/pytest.log
NCU="ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv"
$NCU --log-file gpurun_out/r02m/traffic.csv python scripts/prof_emb.py emb 256 > gpurun_out/r02m/t1.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02m/bench.json 2> gpurun_out/r02m/bench.err; tail -c 500 gpurun_out/r02m/bench.json
