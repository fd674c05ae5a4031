set -x
mkdir -p gpurun_out/r03e
L=gpurun_out/r03e/stress.log
: > $L
for i in 1 2; do
  for m in 1 2 3; do timeout 200 python scripts/seg_stress.py $m 150 37.3 >> $L 2>&1; done
done
B200_SINC_EARLY=1 timeout 200 python scripts/seg_stress.py 1 150 37.3 >> $L 2>&1
timeout 200 python scripts/seg_stress.py 1 60 320 >> $L 2>&1
grep "mode" $L
