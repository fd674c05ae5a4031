set -x
mkdir -p gpurun_out/r02u
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "embedding_parity" > gpurun_out/r02u/pytest.log 2>&1; tail -5 gpurun_out/r02u/pytest.log
EMB_PERF_ENVS="B200_TC4_RES_SMEM=1;B200_TC4_RES_SMEM=0;B200_TC4_RES_SMEM=1;B200_TC4_RES_SMEM=0;B200_TC4_RES_SMEM=1;B200_TC4_RES_SMEM=0" timeout 300 python scripts/emb_perf.py 256 > gpurun_out/r02u/emb_perf.log 2>&1; tail -n 6 gpurun_out/r02u/emb_perf.log
NCU="timeout 300 ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02u/trunk_traffic_256.csv python scripts/prof_emb.py emb 256 > gpurun_out/r02u/t1.log 2>&1
$NCU --set full --import-source on -k regex:conv_tc4_kernel -s 8 -c 2 -o gpurun_out/r02u/conv_tc4 python scripts/prof_emb.py emb 64 > gpurun_out/r02u/t2.log 2>&1
