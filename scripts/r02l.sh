set -x
mkdir -p gpurun_out/r02l
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "embedding_parity or end_to_end" > gpurun_out/r02l/pytest.log 2>&1; tail -3 gpurun_out/r02l/pytest.log
NCU="ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv"
$NCU --log-file gpurun_out/r02l/traffic_g3.csv python scripts/prof_emb.py emb 256 > gpurun_out/r02l/t1.log 2>&1
B200_TC4_G=2 $NCU --log-file gpurun_out/r02l/traffic_g2.csv python scripts/prof_emb.py emb 256 > gpurun_out/r02l/t2.log 2>&1
B200_TC4_G=4 $NCU --log-file gpurun_out/r02l/traffic_g4.csv python scripts/prof_emb.py emb 256 > gpurun_out/r02l/t3.log 2>&1
B200_TC3_DBG=1 $NCU --log-file gpurun_out/r02l/traffic_dbg1.csv python scripts/prof_emb.py emb 256 > gpurun_out/r02l/t4.log 2>&1
B200_TC3_DBG=2 $NCU --log-file gpurun_out/r02l/traffic_dbg2.csv python scripts/prof_emb.py emb 256 > gpurun_out/r02l/t5.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02l/bench.json 2> gpurun_out/r02l/bench.err; tail -c 500 gpurun_out/r02l/bench.json
