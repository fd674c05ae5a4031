set -x
mkdir -p gpurun_out/r02t
EMB_PERF_ENVS="B200_TC4_RES_SMEM=0,B200_TC4_ASLOTS=4;B200_TC4_RES_SMEM=0,B200_TC4_ASLOTS=7;B200_TC4_RES_SMEM=0,B200_TC4_ASLOTS=4" timeout 300 python scripts/emb_perf.py 256 > gpurun_out/r02t/emb_perf_aslots.log 2>&1; tail -n 4 gpurun_out/r02t/emb_perf_aslots.log
cp pyannote_audio_b200/lib/libb200diar.so /tmp/lib_normal.so; cp pyannote_audio_b200/lib/dbg/libb200diar.so pyannote_audio_b200/lib/libb200diar.so
EMB_PERF_ENVS="B200_TC4_RES_SMEM=1" timeout 300 python scripts/emb_perf.py 256 > gpurun_out/r02t/emb_perf_dbg.log 2>&1; head -n 30 gpurun_out/r02t/emb_perf_dbg.log
cp /tmp/lib_normal.so pyannote_audio_b200/lib/libb200diar.so
