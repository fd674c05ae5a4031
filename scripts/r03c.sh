set -x
mkdir -p gpurun_out/r03c
B200_OPTIONS=emb_max_batch=296 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r03c/pytest.log 2>&1; tail -4 gpurun_out/r03c/pytest.log
B="timeout 300 python bench.py --no-cpu-baseline --no-eager-baseline"
$B > gpurun_out/r03c/bench.json 2> gpurun_out/r03c/bench.err; cat gpurun_out/r03c/bench.json; tail -3 gpurun_out/r03c/bench.err
B200_OPTIONS=fbank_share=0 $B > gpurun_out/r03c/bench_noshare.json 2> gpurun_out/r03c/bench_noshare.err; cat gpurun_out/r03c/bench_noshare.json
B200_OPTIONS=emb_max_batch=296 $B > gpurun_out/r03c/bench_emb296.json 2> gpurun_out/r03c/bench_emb296.err; cat gpurun_out/r03c/bench_emb296.json
B200_OPTIONS=emb_max_batch=592 $B > gpurun_out/r03c/bench_emb592.json 2> gpurun_out/r03c/bench_emb592.err; cat gpurun_out/r03c/bench_emb592.json
B200_TIMING=2 $B --steps 2 --warmup 3 > gpurun_out/r03c/bench_timing.json 2> gpurun_out/r03c/bench_timing.err; grep "b200 " gpurun_out/r03c/bench_timing.err | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r03c/smoke.log 2>&1; tail -2 gpurun_out/r03c/smoke.log
B200_TIMING=3 $B --steps 1 --warmup 3 > gpurun_out/r03c/bench_prof.json 2> gpurun_out/r03c/bench_prof.err; grep -A22 "Ordered by" gpurun_out/r03c/bench_prof.err | tail -24
