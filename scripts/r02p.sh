set -x
mkdir -p gpurun_out/r02p
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "segmentation_parity or segmentation_edge or embedding_parity" > gpurun_out/r02p/pytest.log 2>&1; tail -5 gpurun_out/r02p/pytest.log
timeout 300 python scripts/seg_perf.py 4736 tc > gpurun_out/r02p/seg_perf.log 2>&1; cat gpurun_out/r02p/seg_perf.log
EMB_PERF_MC=1 timeout 300 python scripts/emb_perf.py 256 > gpurun_out/r02p/emb_perf.log 2>&1; cat gpurun_out/r02p/emb_perf.log
NCU="timeout 300 ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02p/trunk_traffic_256.csv python scripts/prof_emb.py emb 256 > gpurun_out/r02p/t1.log 2>&1
$NCU --set full --import-source on -k regex:lstm_rec_tc_kernel -s 1 -c 1 -o gpurun_out/r02p/lstm python scripts/prof_emb.py seg 2368 > gpurun_out/r02p/t4.log 2>&1
$NCU --set full --import-source on -k regex:conv_tc4_kernel -s 8 -c 2 -o gpurun_out/r02p/conv_tc4 python scripts/prof_emb.py emb 64 > gpurun_out/r02p/t2.log 2>&1
ls -la gpurun_out/r02p
