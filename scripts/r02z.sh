set -x
mkdir -p gpurun_out/r02z
SEG_PERF_ENVS="B200_GEMM_DBG=0;B200_GEMM_DBG=1;B200_GEMM_DBG=2;B200_GEMM_DBG=0;B200_GEMM_DBG=1;B200_GEMM_DBG=2" timeout 300 python scripts/seg_perf.py 4736 env > gpurun_out/r02z/seg_perf_gemm.log 2>&1; cat gpurun_out/r02z/seg_perf_gemm.log
