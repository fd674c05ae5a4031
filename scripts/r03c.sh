set -x
mkdir -p gpurun_out/r03c
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "embedding_parity or end_to_end or bench_workload or edge_cases" > gpurun_out/r03c/pytest.log 2>&1; tail -4 gpurun_out/r03c/pytest.log
B="timeout 300 python bench.py --no-cpu-baseline --no-eager-baseline"
$B > gpurun_out/r03c/bench.json 2> gpurun_out/r03c/bench.err; cat gpurun_out/r03c/bench.json; tail -3 gpurun_out/r03c/bench.err
B200_OPTIONS=fbank_share=0 $B > gpurun_out/r03c/bench_noshare.json 2> gpurun_out/r03c/bench_noshare.err; cat gpurun_out/r03c/bench_noshare.json
B200_OPTIONS=emb_max_batch=512 $B > gpurun_out/r03c/bench_emb512.json 2> gpurun_out/r03c/bench_emb512.err; cat gpurun_out/r03c/bench_emb512.json
B200_OPTIONS=emb_max_batch=296 $B > gpurun_out/r03c/bench_emb296.json 2> gpurun_out/r03c/bench_emb296.err; cat gpurun_out/r03c/bench_emb296.json
B200_OPTIONS=emb_max_batch=592 $B > gpurun_out/r03c/bench_emb592.json 2> gpurun_out/r03c/bench_emb592.err; cat gpurun_out/r03c/bench_emb592.json
B200_TIMING=2 $B --steps 2 --warmup 3 > gpurun_out/r03c/bench_timing.json 2> gpurun_out/r03c/bench_timing.err; grep "b200 " gpurun_out/r03c/bench_timing.err | tail -6
