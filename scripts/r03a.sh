set -x
mkdir -p gpurun_out/r03a
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "segmentation_parity or segmentation_edge or cfg2" > gpurun_out/r03a/pytest.log 2>&1; tail -4 gpurun_out/r03a/pytest.log
SEG_PERF_ENVS="X=0;X=1" timeout 300 python scripts/seg_perf.py 4736 env > gpurun_out/r03a/seg_perf.log 2>&1; cat gpurun_out/r03a/seg_perf.log
NCU="timeout 300 ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r03a/seg_launches_2368.csv python scripts/prof_emb.py seg 2368 > gpurun_out/r03a/t6.log 2>&1
$NCU --set full --import-source on -k regex:sinc_tc_kernel -s 1 -c 1 -o gpurun_out/r03a/sinc_tc python scripts/prof_emb.py seg 2368 > gpurun_out/r03a/t4.log 2>&1
$NCU --set full --import-source on -k regex:conv5_tc_kernel -s 2 -c 2 -o gpurun_out/r03a/conv5_tc python scripts/prof_emb.py seg 2368 > gpurun_out/r03a/t5.log 2>&1
