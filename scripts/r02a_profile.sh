set -x
mkdir -p gpurun_out/r02a
cd $GRAFT_REPO_ROOT
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r02a/trunk_traffic_256.csv python scripts/prof_emb.py emb 256 > gpurun_out/r02a/t1.log 2>&1
$NCU --set full --import-source on -k regex:conv_tc4_kernel -s 8 -c 2 -o gpurun_out/r02a/conv_tc4 python scripts/prof_emb.py emb 64 > gpurun_out/r02a/t2.log 2>&1
$NCU --set full --import-source on -k regex:conv_block32_kernel -s 3 -c 1 -o gpurun_out/r02a/conv_block32 python scripts/prof_emb.py emb 64 > gpurun_out/r02a/t3.log 2>&1
$NCU --set full --import-source on -k regex:sinc_tc_kernel -s 1 -c 1 -o gpurun_out/r02a/sinc_tc python scripts/prof_emb.py seg 2368 > gpurun_out/r02a/t4.log 2>&1
$NCU --set full --import-source on -k regex:conv5_tc_kernel -s 2 -c 2 -o gpurun_out/r02a/conv5_tc python scripts/prof_emb.py seg 2368 > gpurun_out/r02a/t5.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02a/seg_launches_2368.csv python scripts/prof_emb.py seg 2368 > gpurun_out/r02a/t6.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02a/bench.json 2> gpurun_out/r02a/bench.err
tail -c 600 gpurun_out/r02a/bench.json
ls -la gpurun_out/r02a
