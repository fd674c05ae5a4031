set -x
mkdir -p gpurun_out/r02s
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "embedding_parity" > gpurun_out/r02s/pytest.log 2>&1; tail -5 gpurun_out/r02s/pytest.log
EMB_PERF_ENVS="B200_TC4_RES_SMEM=1;B200_TC4_RES_SMEM=0;B200_TC4_RES_SMEM=1;B200_TC4_RES_SMEM=0;B200_TC4_RES_SMEM=1" timeout 300 python scripts/emb_perf.py 256 > gpurun_out/r02s/emb_perf.log 2>&1; tail -n 6 gpurun_out/r02s/emb_perf.log
NCU="timeout 300 ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02s/trunk_traffic_256.csv python scripts/prof_emb.py emb 256 > gpurun_out/r02s/t1.log 2>&1
$NCU --set full --import-source on -k regex:conv_tc4_kernel -s 8 -c 2 -o gpurun_out/r02s/conv_tc4 python scripts/prof_emb.py emb 64 > gpurun_out/r02s/t2.log 2>&1
$NCU --set full --import-source on -k regex:lstm_rec_tc_kernel -s 1 -c 1 -o gpurun_out/r02s/lstm python scripts/prof_emb.py seg 2368 > gpurun_out/r02s/t4.log 2>&1
