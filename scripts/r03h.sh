set -x
mkdir -p gpurun_out/r03h
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $T scripts/pool_check.py > gpurun_out/r03h/pool_check.log 2>&1; grep -v Warning gpurun_out/r03h/pool_check.log | tail -8
timeout 400 $T bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r03h/bench2.json 2> gpurun_out/r03h/bench2.err; cat gpurun_out/r03h/bench2.json; tail -3 gpurun_out/r03h/bench2.err
timeout 300 $T scripts/sharded_check.py > gpurun_out/r03h/sharded_check.log 2>&1; grep -v Warning gpurun_out/r03h/sharded_check.log | tail -4
