set -x
mkdir -p gpurun_out/r03k
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521"
timeout 110 $T scripts/pool_check.py > gpurun_out/r03k/pool_check4.log 2>&1; grep "pool_check" gpurun_out/r03k/pool_check4.log | grep -v identical | tail -5; grep -c identical gpurun_out/r03k/pool_check4.log
