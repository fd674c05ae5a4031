set -x
mkdir -p gpurun_out/r02v
EMB_PERF_ENVS="B200_TC3_DBG=0;B200_TC3_DBG=1;B200_TC3_DBG=2;B200_TC3_DBG=0;B200_TC3_DBG=1;B200_TC3_DBG=2" timeout 300 python scripts/emb_perf.py 256 > gpurun_out/r02v/emb_perf_dbg.log 2>&1; tail -n 6 gpurun_out/r02v/emb_perf_dbg.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02v/bench.json 2> gpurun_out/r02v/bench.err; tail -c 1200 gpurun_out/r02v/bench.json
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02v/pytest.log 2>&1; tail -5 gpurun_out/r02v/pytest.log
