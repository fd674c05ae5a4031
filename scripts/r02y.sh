set -x
mkdir -p gpurun_out/r02y
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "embedding_parity or end_to_end" > gpurun_out/r02y/pytest.log 2>&1; tail -5 gpurun_out/r02y/pytest.log
EMB_PERF_ENVS="X=1" timeout 300 python scripts/emb_perf.py 256 > gpurun_out/r02y/emb_perf.log 2>&1; tail -n 2 gpurun_out/r02y/emb_perf.log
NCU="timeout 300 ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02y/trunk_traffic_256.csv python scripts/prof_emb.py emb 256 > gpurun_out/r02y/t1.log 2>&1
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02y/seg_launches_2368.csv python scripts/prof_emb.py seg 2368 > gpurun_out/r02y/t6.log 2>&1
