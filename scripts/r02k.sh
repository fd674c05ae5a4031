set -x
mkdir -p gpurun_out/r02k
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 scripts/pool_check.py > gpurun_out/r02k/pool_check.log 2>&1; tail -15 gpurun_out/r02k/pool_check.log
$TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02k/bench_n2_p2p.json 2> gpurun_out/r02k/bench_n2_p2p.err; tail -c 900 gpurun_out/r02k/bench_n2_p2p.json; tail -3 gpurun_out/r02k/bench_n2_p2p.err
$TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --collective nccl > gpurun_out/r02k/bench_n2_nccl.json 2> gpurun_out/r02k/bench_n2_nccl.err; tail -c 900 gpurun_out/r02k/bench_n2_nccl.json
$TR --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --parallelism files > gpurun_out/r02k/bench_n2_files.json 2> gpurun_out/r02k/bench_n2_files.err; tail -c 400 gpurun_out/r02k/bench_n2_files.json
$TR --master-port 29515 scripts/sharded_check.py > gpurun_out/r02k/sharded_check.log 2>&1; tail -3 gpurun_out/r02k/sharded_check.log
