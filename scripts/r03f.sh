set -x
mkdir -p gpurun_out/r03f
L=gpurun_out/r03f/repro.log
: > $L
nvidia-smi --query-gpu=name,serial,uuid,vbios_version,clocks.sm,temperature.gpu --format=csv >> $L 2>&1
for i in 1 2 3 4 5 6; do timeout 200 python scripts/seg_repro.py >> $L 2>&1; done
grep -v Warning $L | tail -20
