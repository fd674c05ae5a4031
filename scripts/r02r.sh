set -x
mkdir -p gpurun_out/r02r
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "embedding_parity" > gpurun_out/r02r/pytest.log 2>&1; tail -5 gpurun_out/r02r/pytest.log
EMB_PERF_ENVS="B200_TC4_RES_SMEM=0;B200_TC4_RES_SMEM=0,B200_TC4_WRAP=1;B200_TC4_RES_SMEM=0,B200_TC4_WRAP=0" timeout 300 python scripts/emb_perf.py 256 > gpurun_out/r02r/emb_perf0.log 2>&1; cat gpurun_out/r02r/emb_perf0.log
EMB_PERF_ENVS="B200_TC4_RES_SMEM=1;B200_TC4_RES_SMEM=0;B200_TC4_RES_SMEM=1" timeout 300 python scripts/emb_perf.py 256 > gpurun_out/r02r/emb_perf1.log 2>&1; tail -n 4 gpurun_out/r02r/emb_perf1.log
if grep -q Error gpurun_out/r02r/emb_perf1.log; then
  timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python scripts/prof_emb.py emb 256 > gpurun_out/r02r/sanitizer.log 2>&1; grep -v "^=========     Host Frame\|^=========         in " gpurun_out/r02r/sanitizer.log | head -60
fi
